python -m pytest tests -m gpu -q --timeout=900 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head
for m in 0; do
  DBW_DEBUG_SKIP=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:round(v,3) for k,v in d['roofline']['kernels_ms_per_step'].items()})"
done
