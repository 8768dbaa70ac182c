#!/bin/bash
# usage (GPU box with N >= 2 GPUs, via `gpurun --gpus N`): bash scripts/gpu_multi.sh <tag> <N> [steps]
TAG=${1:-m}; N=${2:-2}; STEPS=${3:-20}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x --timeout=500 2>&1 | grep -vE "^\s+\[|tensor\(" | tail -30 > gpurun_out/multi_tests_${TAG}.log
tail -15 gpurun_out/multi_tests_${TAG}.log
for C in auto nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29871 bench.py \
      --gpus $N --steps $STEPS --warmup 5 --collective $C > gpurun_out/bench_${TAG}_n${N}_${C}.json 2> gpurun_out/bench_${TAG}_n${N}_${C}.err
  tail -2 gpurun_out/bench_${TAG}_n${N}_${C}.err; cat gpurun_out/bench_${TAG}_n${N}_${C}.json
done
