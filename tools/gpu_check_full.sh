#!/bin/bash
# tools/gpu_check_full.sh TAG -- run ON THE GPU BOX: whole GPU suite, BA breakdown, default bench
set -u
TAG=${1:-r01h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/test_all_${TAG}.log 2>&1
echo "all tests exit $?"; tail -6 gpurun_out/test_all_${TAG}.log
timeout 300 python tools/ba_profile.py > gpurun_out/ba_profile_${TAG}.log 2>&1
echo "ba_profile exit $?"; head -9 gpurun_out/ba_profile_${TAG}.log
timeout 700 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -c 600 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
