#!/bin/bash
set -u
TAG=${1:-r01l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_match_gpu.py tests/test_pose_opt_gpu.py tests/test_pipeline_gpu.py tests/test_reloc_match_gpu.py -q -m gpu > gpurun_out/test_track_${TAG}.log 2>&1
echo "tracking tests exit $?"; tail -4 gpurun_out/test_track_${TAG}.log
timeout 300 python bench.py --no-ba --no-lines --no-stereo --no-mapping --no-cpu-baseline > gpurun_out/bench_fe_${TAG}.json 2>&1
echo "bench exit $?"; head -c 200 gpurun_out/bench_fe_${TAG}.json
