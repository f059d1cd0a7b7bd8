#!/bin/bash
# final driver-like validation
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python -c "
import json
r=json.load(open('gpurun_out/bench_ref.json')); o=json.load(open('gpurun_out/bench_ours.json'))
print('ref', r['value'], 'ours value', o['value'], 'ms', o['ms_per_step'], 'e2e', o['e2e']['value'], o['e2e']['ms_per_step'], 'ratio e2e/ref', o['e2e']['value']/r['value'], o['clocks'], o['gpu_launches'])"
python benchmarks/run_configs.py > gpurun_out/configs_final.jsonl 2>/dev/null; cut -c1-200 gpurun_out/configs_final.jsonl
