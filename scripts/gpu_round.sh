#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --n 256 --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
FDB_MINB=2 python bench.py --n 256 --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('minb2', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
ncu --set full --clock-control none --import-source on -k regex:helmholtz -s 3 -c 1 -o gpurun_out/prof_action10 python bench.py --n 128 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log
