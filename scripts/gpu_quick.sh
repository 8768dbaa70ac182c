#!/bin/bash
# usage (on the GPU box, via gpurun): bash scripts/gpu_quick.sh  -> parity tests + smoke
python -m pytest tests -m gpu -q --timeout=900 2>&1 | grep -vE "^\s+\[|^\s+\.\.\.|tensor\(" | tail -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
