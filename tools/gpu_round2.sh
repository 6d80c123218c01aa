#!/bin/bash
# tools/gpu_round2.sh TAG -- run ON THE GPU BOX: new-feature tests, default bench (with the mapping leg), BA kernel breakdown
set -u
TAG=${1:-r01f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fuse_gpu.py tests/test_bow_gpu.py ${EXTRA_TESTS:-} -q -m gpu > gpurun_out/test_new_${TAG}.log 2>&1
echo "new tests exit $?"; tail -4 gpurun_out/test_new_${TAG}.log
timeout 700 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --batch 512 --no-ba --no-lines --no-stereo --no-mapping --no-cpu-baseline > gpurun_out/bench_b512_${TAG}.json 2>&1
echo "bench512 exit $?"
timeout 300 python tools/ba_profile.py > gpurun_out/ba_profile_${TAG}.log 2>&1
echo "ba_profile exit $?"; tail -30 gpurun_out/ba_profile_${TAG}.log
