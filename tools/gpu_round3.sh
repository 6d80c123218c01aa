#!/bin/bash
# tools/gpu_round3.sh TAG -- run ON THE GPU BOX: whole GPU suite, default bench, BA breakdown + ncu capture of the BA kernels
set -u
TAG=${1:-r01g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/test_all_${TAG}.log 2>&1
echo "all tests exit $?"; tail -6 gpurun_out/test_all_${TAG}.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_essential_gpu.py tests/test_bow_gpu.py -q -m gpu -k "edge or seed0 or ransac-1 or pairs-0" > gpurun_out/sanitizer_${TAG}.log 2>&1
echo "sanitizer exit $?"; tail -3 gpurun_out/sanitizer_${TAG}.log
timeout 700 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -c 1200 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 300 python tools/ba_profile.py > gpurun_out/ba_profile_${TAG}.log 2>&1
echo "ba_profile exit $?"; head -9 gpurun_out/ba_profile_${TAG}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ba_linearize_kernel|ba_solve_kernel|ba_update_kernel|ba_reduce_kernel" -s 8 -c 4 -o gpurun_out/prof_ba_${TAG} -f python tools/ba_profile.py > gpurun_out/ba_ncu_${TAG}.log 2>&1
echo "ba ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fuse_points_kernel|bow_transform_kernel|bow_match_kernel" -c 6 -o gpurun_out/prof_map_${TAG} -f python bench.py --steps 1 --warmup 3 --batch 64 --no-ba --no-lines --no-stereo --no-cpu-baseline > gpurun_out/map_ncu_${TAG}.log 2>&1
echo "map ncu exit $?"
ls gpurun_out
