#!/bin/bash
# tools/gpu_r02.sh TAG [what] -- run ON THE GPU BOX (under gpurun).  what: any of "tests bench ref prof proflines profba"
# (default: tests bench ref prof).  Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ here.
set -u
TAG=${1:-r02a}
WHAT=${2:-"tests bench ref prof"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_${TAG}.txt 2>&1
if [[ " $WHAT " == *" tmadbg "* ]]; then
  T="tests/test_orb_gpu.py -x -q -m gpu"
  PLP_BLUR_NO_TMA=1 timeout 200 python -m pytest $T > gpurun_out/tma_none_${TAG}.log 2>&1; echo "no-tma exit $?"; tail -2 gpurun_out/tma_none_${TAG}.log
  PLP_TMA_MAPS=global timeout 200 python -m pytest $T > gpurun_out/tma_global_${TAG}.log 2>&1; echo "tma global exit $?"; tail -2 gpurun_out/tma_global_${TAG}.log
  timeout 200 python -m pytest $T > gpurun_out/tma_param_${TAG}.log 2>&1; echo "tma param exit $?"; tail -2 gpurun_out/tma_param_${TAG}.log
  timeout 300 compute-sanitizer --tool memcheck python -m pytest $T -k "1234" > gpurun_out/tma_san_param_${TAG}.log 2>&1; echo "sanitizer param exit $?"
  grep -m 12 -A12 "=========" gpurun_out/tma_san_param_${TAG}.log | head -60
  PLP_TMA_MAPS=global timeout 300 compute-sanitizer --tool memcheck python -m pytest $T -k "1234" > gpurun_out/tma_san_global_${TAG}.log 2>&1; echo "sanitizer global exit $?"
  grep -m 12 -A12 "=========" gpurun_out/tma_san_global_${TAG}.log | head -40
fi
if [[ " $WHAT " == *" tests "* ]]; then
  # the kernels changed most recently first, under a short timeout (a hung copy engine must not eat the box)
  timeout 300 python -m pytest tests/test_orb_gpu.py tests/test_golden.py tests/test_pose_opt_gpu.py tests/test_match_gpu.py tests/test_pipeline_gpu.py -q -m gpu -x > gpurun_out/test_first_${TAG}.log 2>&1
  echo "first tests exit $?"; tail -4 gpurun_out/test_first_${TAG}.log
  timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/test_all_${TAG}.log 2>&1
  echo "gpu tests exit $?"; tail -6 gpurun_out/test_all_${TAG}.log
fi
if [[ " $WHAT " == *" bench "* ]]; then
  timeout 900 python bench.py --detail gpurun_out/bench_detail_${TAG}.json > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  echo "bench exit $?"; tail -c 1500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err | cut -c1-600
fi
if [[ " $WHAT " == *" ref "* ]]; then
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref2_${TAG}.json 2>&1
  echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref_${TAG}.json; cut -c1-300 gpurun_out/bench_ref2_${TAG}.json
fi
FE="--steps 2 --warmup 3 --batch 64 --streams 1 --no-cpu-baseline --no-ba --no-lines --no-stereo --no-mapping --no-pose3 --no-latency"
if [[ " $WHAT " == *" prof "* ]]; then
  # (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py $FE > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
  echo "launch list exit $?"
  # (2) full capture of the hot kernels of one warm step (setup extraction = 11 matched launches, a step = 14)
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:"fast_cells_kernel_v2|blur_tiles|quadtree_kernel|describe_kernel|point_match_kernel|pose_opt_kernel|pyr_resize_kernel" \
      -s 53 -c 14 -o gpurun_out/prof_${TAG} -f python bench.py $FE > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
  echo "full capture exit $?"
fi
if [[ " $WHAT " == *" proflines "* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_lines_${TAG}.csv \
      python bench.py --only-lines --steps 2 --warmup 3 --line-batch 64 --no-cpu-baseline > gpurun_out/bench_lines_under_ncu_${TAG}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:"lsd_scale_kernel|lsd_sort_kernel|lsd_grow_kernel|keyline_kernel|lbd_gradient_kernel|lbd_kernel" \
      -s 18 -c 6 -o gpurun_out/prof_lines_${TAG} -f \
      python bench.py --only-lines --steps 2 --warmup 3 --line-batch 64 --no-cpu-baseline > gpurun_out/bench_lines_under_ncu_full_${TAG}.log 2>&1
  echo "line profile exit $?"
fi
if [[ " $WHAT " == *" ooodbg "* ]]; then
  PLP_TEST_OOO=1 timeout 60 python -m pytest tests/test_lines_gpu.py -q -m gpu -x -s -k "lines-2-shape1-3" > gpurun_out/test_lines_ooodbg_${TAG}.log 2>&1
  echo "ooodbg exit $?"; grep "ooo timeout\|\[ooo\]\|passed\|failed" gpurun_out/test_lines_ooodbg_${TAG}.log | head -20 | cut -c1-400
  WHAT="$WHAT ooo"
fi
if [[ " $WHAT " == *" ooo "* ]]; then
  # the out-of-order region growing kernel, under short timeouts of its own
  PLP_TEST_OOO=1 timeout 150 python -m pytest tests/test_lines_gpu.py -q -m gpu -x -s > gpurun_out/test_lines_ooo_${TAG}.log 2>&1
  echo "lines tests (ooo) exit $?"; grep "\[ooo\]" gpurun_out/test_lines_ooo_${TAG}.log | head -12; tail -6 gpurun_out/test_lines_ooo_${TAG}.log | cut -c1-300
  PLP_TEST_OOO=1 timeout 150 python tools/lsd_latency.py 8:0 > gpurun_out/lsd_latency_ooo_${TAG}.log 2>&1; echo "lsd latency (ooo) exit $?"; cat gpurun_out/lsd_latency_ooo_${TAG}.log | cut -c1-600
fi
if [[ " $WHAT " == *" lsd "* ]]; then
  timeout 900 python -m pytest tests/test_lines_gpu.py -q -m gpu -x -s > gpurun_out/test_lines_${TAG}.log 2>&1
  echo "lines tests exit $?"; grep "\[mw\]" gpurun_out/test_lines_${TAG}.log | head -20; tail -5 gpurun_out/test_lines_${TAG}.log
  timeout 600 python tools/lsd_latency.py 8:0 > gpurun_out/lsd_latency_${TAG}.log 2>&1; echo "lsd latency exit $?"; cat gpurun_out/lsd_latency_${TAG}.log
fi
if [[ " $WHAT " == *" ba "* ]]; then
  timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_plane_gpu.py -q -m gpu -x > gpurun_out/test_ba_${TAG}.log 2>&1
  echo "ba tests exit $?"; tail -5 gpurun_out/test_ba_${TAG}.log
  timeout 300 python tools/ba_profile.py > gpurun_out/ba_profile_${TAG}.log 2>&1; echo "ba_profile exit $?"; cat gpurun_out/ba_profile_${TAG}.log
  timeout 600 python bench.py --only-ba --steps 10 --warmup 3 > gpurun_out/bench_ba_${TAG}.json 2> gpurun_out/bench_ba_${TAG}.err
  echo "bench ba exit $?"; cut -c1-1500 gpurun_out/bench_ba_${TAG}.json
fi
if [[ " $WHAT " == *" profba "* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_ba_${TAG}.csv \
      python bench.py --only-ba --no-ba-large --steps 2 --warmup 1 > gpurun_out/bench_ba_under_ncu_${TAG}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ba_" -s 60 -c 12 -o gpurun_out/prof_ba_${TAG} -f \
      python bench.py --only-ba --no-ba-large --steps 2 --warmup 1 > gpurun_out/bench_ba_under_ncu_full_${TAG}.log 2>&1
  echo "ba profile exit $?"
fi
ls -la gpurun_out | tail -30
if [[ " $WHAT " == *" lineslat "* ]]; then
  timeout 200 python bench.py --only-lines --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench_lines_${TAG}.json 2> gpurun_out/bench_lines_${TAG}.err
  echo "bench lines exit $?"; cut -c1-1500 gpurun_out/bench_lines_${TAG}.json
fi
