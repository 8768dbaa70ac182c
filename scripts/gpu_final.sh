#!/bin/bash
# usage (GPU box, 1 GPU, via gpurun): bash scripts/gpu_final.sh <tag>  -> ncu evidence + the bench lines of the three workloads + the reference arm
TAG=${1:-r2}
bash scripts/gpu_profile.sh ${TAG} > /dev/null 2>&1
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${TAG}_dtu_n1.json 2> gpurun_out/bench_${TAG}_dtu_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
for w in bmvs stress; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_${w}_n1.json 2> gpurun_out/bench_${TAG}_${w}_n1.err
done
ls -la gpurun_out | grep ${TAG}
