#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in 8 16 64 128; do
echo "== FDB_CHUNK=$c"
FDB_CHUNK=$c python bench.py --n 256 --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
done
python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python -c "
import json
r=json.load(open('gpurun_out/bench_ref.json')); o=json.load(open('gpurun_out/bench_ours.json'))
print('ref', r['value'], 'ours value', o['value'], 'ms', o['ms_per_step'], 'e2e', o['e2e']['value'], o['e2e']['ms_per_step'], 'cpu', o['cpu_baseline']['value'], 'ratio e2e/ref', o['e2e']['value']/r['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/launches_r01.csv | cut -c1-330
