#!/bin/bash
# tools/gpu_multi.sh TAG -- run ON AN 8-GPU BOX (gpurun --gpus 8): landmark-sharded local BA at 2 / 4 / 8 ranks against the
# oracle (tests/run_ba_multigpu.py), the BA curve of bench.py at 1 / 2 / 4 / 8 GPUs, and the full bench at 8 GPUs.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi_${TAG}.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for W in 2 4 8; do
  timeout 300 $TR --nproc-per-node $W --master-port $((29510 + W)) tests/run_ba_multigpu.py > gpurun_out/ba_multigpu_w${W}_${TAG}.log 2>&1
  echo "run_ba_multigpu world $W exit $?"; grep "ba multi-gpu" gpurun_out/ba_multigpu_w${W}_${TAG}.log
done
timeout 300 python bench.py --gpus 1 --only-ba --steps 5 --warmup 3 > gpurun_out/bench_ba_n1_${TAG}.json 2> gpurun_out/bench_ba_n1_${TAG}.err
echo "ba n1 exit $?"; cut -c1-600 gpurun_out/bench_ba_n1_${TAG}.json
for W in 2 4 8; do
  timeout 400 $TR --nproc-per-node $W --master-port $((29530 + W)) bench.py --gpus $W --only-ba --steps 5 --warmup 3 > gpurun_out/bench_ba_n${W}_${TAG}.json 2> gpurun_out/bench_ba_n${W}_${TAG}.err
  echo "ba n$W exit $?"; tail -1 gpurun_out/bench_ba_n${W}_${TAG}.json | cut -c1-600
done
timeout 900 $TR --nproc-per-node 8 --master-port 29577 bench.py --gpus 8 --steps 10 --warmup 3 --detail gpurun_out/bench_detail_n8_${TAG}.json > gpurun_out/bench_n8_${TAG}.json 2> gpurun_out/bench_n8_${TAG}.err
echo "bench n8 exit $?"; tail -1 gpurun_out/bench_n8_${TAG}.json | cut -c1-1500
