#!/bin/bash
# tools/gpu_round6.sh TAG -- run ON THE GPU BOX: ORB tests first (FAST v2), whole suite, default bench, A/B against FAST v1
set -u
TAG=${1:-r01j}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_orb_gpu.py tests/test_pipeline_gpu.py tests/test_stereo_gpu.py -q -m gpu > gpurun_out/test_orb_${TAG}.log 2>&1
echo "orb tests exit $?"; tail -6 gpurun_out/test_orb_${TAG}.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_orb_gpu.py --deselect tests/test_pipeline_gpu.py --deselect tests/test_stereo_gpu.py > gpurun_out/test_rest_${TAG}.log 2>&1
echo "rest tests exit $?"; tail -4 gpurun_out/test_rest_${TAG}.log
timeout 700 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; head -c 700 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
PLP_FAST_V1=1 timeout 300 python bench.py --no-ba --no-lines --no-stereo --no-mapping --no-cpu-baseline > gpurun_out/bench_fastv1_${TAG}.json 2>&1
echo "bench v1 exit $?"; head -c 300 gpurun_out/bench_fastv1_${TAG}.json
