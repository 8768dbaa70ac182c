#!/bin/bash
# usage (GPU box with N GPUs, via `gpurun --gpus N`): bash scripts/gpu_scale.sh <tag> <N> "<workloads>" [steps]
TAG=${1:-s}; N=${2:-2}; WL=${3:-dtu}; STEPS=${4:-30}
mkdir -p gpurun_out
for W in $WL; do
  S=$STEPS; [ "$W" != "dtu" ] && S=5
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29871 bench.py \
      --gpus $N --steps $S --warmup 5 --workload $W > gpurun_out/bench_${TAG}_${W}_n${N}.json 2> gpurun_out/bench_${TAG}_${W}_n${N}.err
  tail -1 gpurun_out/bench_${TAG}_${W}_n${N}.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$W N=%d views/s=%.0f ms/step=%.3f e2e=%.0f breakdown=%s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v, 3) for k, v in d['breakdown_ms_per_step'].items()}))
except Exception as e:
    print('$W failed', e); print(open('gpurun_out/bench_${TAG}_${W}_n${N}.err').read()[-1500:])
"
done
