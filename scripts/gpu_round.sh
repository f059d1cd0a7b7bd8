#!/bin/bash
for k in 32 64 128; do
echo "== FDB_PIPELINE_CHUNKS=$k"
FDB_PIPELINE_CHUNKS=$k python bench.py --n 256 --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'])"
done
