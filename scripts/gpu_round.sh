#!/bin/bash
for z in 0 8 16 24 48; do
echo "== FDB_ZERO_CTAS=$z"
FDB_ZERO_CTAS=$z python bench.py --n 256 --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['gpu_launches'])"
done
