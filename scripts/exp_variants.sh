#!/bin/bash
# A/B of kernel build variants in ONE gpu call: bench the default library, then every differentiable-blocksworld_b200/_exp/lib_*.so
# (built with -D experiment macros; loaded through DBW_RENDER_LIB).  Prints views/s, ms/step and the per-kernel times.
run() { timeout 150 python bench.py --no-cpu-baseline --steps ${STEPS:-100} 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms_per_step']
print('$1', round(d['value']), round(d['ms_per_step'],4), ' '.join(f'{v:.3f}' for v in k.values()))"; }
run default
shopt -s nullglob
for so in differentiable-blocksworld_b200/_exp/lib_*.so; do DBW_RENDER_LIB=$PWD/$so run $(basename $so .so); done
run default
