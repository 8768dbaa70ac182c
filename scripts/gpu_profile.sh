#!/bin/bash
# usage (GPU box, via gpurun): bash scripts/gpu_profile.sh <tag>
# 1) launch list of the bench command (gpu__time_duration per launch, cold-cache, serialised)
# 2) one `--set full` capture of the four raster kernels of one step
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_ -s 16 -c 4 -o gpurun_out/prof_${TAG} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out | tail -6
