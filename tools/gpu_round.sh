#!/bin/bash
# tools/gpu_round.sh TAG -- run ON THE GPU BOX (under gpurun): new-feature tests first, then the whole GPU suite, the
# default bench, and the ncu passes (tools/profile.sh).  Everything lands in gpurun_out/.
set -u
TAG=${1:-r01e}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_${TAG}.txt 2>&1
timeout 600 python -m pytest tests/test_fuse_gpu.py tests/test_bow_gpu.py -q -m gpu > gpurun_out/test_new_${TAG}.log 2>&1
echo "new tests exit $?"; tail -5 gpurun_out/test_new_${TAG}.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_fuse_gpu.py --deselect tests/test_bow_gpu.py > gpurun_out/test_all_${TAG}.log 2>&1
echo "all tests exit $?"; tail -5 gpurun_out/test_all_${TAG}.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fuse_gpu.py tests/test_bow_gpu.py -q -m gpu -k "edge or seed0 or 10-3-1" > gpurun_out/sanitizer_${TAG}.log 2>&1
echo "sanitizer exit $?"; tail -3 gpurun_out/sanitizer_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -c 600 gpurun_out/bench_${TAG}.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1
echo "ref exit $?"
timeout 900 bash tools/profile.sh ${TAG} 64 > gpurun_out/profile_${TAG}.log 2>&1
echo "profile exit $?"
ls gpurun_out
