#!/bin/bash
# usage (GPU box, via gpurun): bash scripts/gpu_r2.sh <tag> [pytest-args]   -> GPU tests + smoke + a short bench line
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --timeout=1500 ${@:2} 2>&1 | grep -vE "^\s+\[|tensor\(" > gpurun_out/tests_${TAG}.log
grep -E "ambiguous|leaf gradients|ORACLE|masked" gpurun_out/tests_${TAG}.log | cut -c1-400
tail -25 gpurun_out/tests_${TAG}.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
