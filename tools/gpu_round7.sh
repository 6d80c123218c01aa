#!/bin/bash
set -u
TAG=${1:-r01k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_orb_gpu.py tests/test_pipeline_gpu.py tests/test_stereo_gpu.py -q -m gpu > gpurun_out/test_orb_${TAG}.log 2>&1
echo "orb tests exit $?"; tail -4 gpurun_out/test_orb_${TAG}.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_orb_gpu.py -q -m gpu > gpurun_out/sanitizer_orb_${TAG}.log 2>&1
echo "sanitizer exit $?"; tail -3 gpurun_out/sanitizer_orb_${TAG}.log
timeout 300 python bench.py --no-ba --no-lines --no-stereo --no-mapping --no-cpu-baseline > gpurun_out/bench_fe_${TAG}.json 2>&1
echo "bench exit $?"; head -c 200 gpurun_out/bench_fe_${TAG}.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fe_'+__import__('sys').argv[1]+'.json').read().strip().splitlines()[-1]) if False else None
PY
