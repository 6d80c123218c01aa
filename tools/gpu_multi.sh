#!/bin/bash
# tools/gpu_multi.sh TAG "W1 W2 ..." [full] -- run ON A MULTI-GPU BOX (gpurun --gpus N): landmark-sharded local BA at the given
# world sizes against the oracle (tests/run_ba_multigpu.py) and the BA curve of bench.py (--only-ba) at the same sizes;
# "full" adds the complete bench.py line at the largest world size (what the driver runs for SCALE).
set -u
TAG=${1:-r02}
WORLDS=${2:-"2 4 8"}
FULL=${3:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi_${TAG}.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
LAST=1
for W in $WORLDS; do
  LAST=$W
  timeout 300 $TR --nproc-per-node $W --master-port $((29510 + W)) tests/run_ba_multigpu.py > gpurun_out/ba_multigpu_w${W}_${TAG}.log 2>&1
  echo "run_ba_multigpu world $W exit $?"; grep "ba multi-gpu" gpurun_out/ba_multigpu_w${W}_${TAG}.log
  timeout 400 $TR --nproc-per-node $W --master-port $((29530 + W)) bench.py --gpus $W --only-ba --steps 5 --warmup 3 > gpurun_out/bench_ba_n${W}_${TAG}.json 2> gpurun_out/bench_ba_n${W}_${TAG}.err
  echo "ba n$W exit $?"; tail -1 gpurun_out/bench_ba_n${W}_${TAG}.json | cut -c1-700
  PLP_BA_PEER=0 timeout 400 $TR --nproc-per-node $W --master-port $((29550 + W)) bench.py --gpus $W --only-ba --steps 5 --warmup 3 > gpurun_out/bench_ba_nccl_n${W}_${TAG}.json 2> gpurun_out/bench_ba_nccl_n${W}_${TAG}.err
  echo "ba (nccl) n$W exit $?"; tail -1 gpurun_out/bench_ba_nccl_n${W}_${TAG}.json | cut -c1-500
done
if [[ "$FULL" == "full" ]]; then
  timeout 900 $TR --nproc-per-node $LAST --master-port 29577 bench.py --gpus $LAST --steps 10 --warmup 3 --detail gpurun_out/bench_detail_n${LAST}_${TAG}.json > gpurun_out/bench_n${LAST}_${TAG}.json 2> gpurun_out/bench_n${LAST}_${TAG}.err
  echo "bench n$LAST exit $?"; tail -1 gpurun_out/bench_n${LAST}_${TAG}.json | cut -c1-1800; tail -3 gpurun_out/bench_n${LAST}_${TAG}.err | cut -c1-400
  timeout 300 $TR --nproc-per-node $LAST --master-port 29578 bench.py --impl reference --gpus $LAST --steps 3 --warmup 1 > gpurun_out/bench_ref_n${LAST}_${TAG}.json 2>&1
  echo "ref n$LAST exit $?"; tail -1 gpurun_out/bench_ref_n${LAST}_${TAG}.json | cut -c1-400
fi
