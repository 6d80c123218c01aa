#!/bin/bash
# tools/profile.sh -- run ON THE GPU BOX (under gpurun): launch list + one full ncu capture of the hot kernels of
# one bench step.  Outputs land in gpurun_out/ and are summarised into profiles/ by tools/summarize_profiles.py.
# Launch bookkeeping (bench.py --no-ba): setup extraction = 11 launches (7 resize, FAST, blur, quadtree, describe),
# one step = 20 launches (11 extraction + prep, project, match, gate, project, match, gather, pose, finish).
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
BATCH=${2:-64}
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 71 -c 40 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --batch ${BATCH} --no-cpu-baseline --no-ba --no-lines --no-stereo > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# (2) full capture of the hot kernels of one step (one launch each)
ncu --set full --clock-control none --import-source on \
    -k regex:"fast_cells_kernel_v2|blur_tiles_kernel|quadtree_kernel|describe_kernel|point_match_kernel|pose_opt_kernel|pyr_resize_kernel" \
    -s 53 -c 14 -o gpurun_out/prof_${TAG} -f \
    python bench.py --steps 2 --warmup 3 --batch ${BATCH} --no-cpu-baseline --no-ba --no-lines --no-stereo > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
# (3) line front end (LSD + LBD): launch list and one full capture of one step at the same batch
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_lines_${TAG}.csv \
    python bench.py --only-lines --steps 2 --warmup 3 --line-batch ${BATCH} --no-cpu-baseline > gpurun_out/bench_lines_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"lsd_scale_kernel|lsd_sort_kernel|lsd_grow_kernel|keyline_kernel|lbd_gradient_kernel|lbd_kernel" \
    -s 18 -c 6 -o gpurun_out/prof_lines_${TAG} -f \
    python bench.py --only-lines --steps 2 --warmup 3 --line-batch ${BATCH} --no-cpu-baseline > gpurun_out/bench_lines_under_ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
