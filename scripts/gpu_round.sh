#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --n 256 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r01b.json
python -c "import json; d=json.load(open('gpurun_out/bench_r01b.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['e2e']['ms_per_step'], d['roofline']['fp64'])"
python benchmarks/run_configs.py > gpurun_out/configs_r01b.jsonl 2> gpurun_out/configs.err
cut -c1-260 gpurun_out/configs_r01b.jsonl
ncu --set full --clock-control none --import-source on -k regex:helmholtz -s 3 -c 1 -o gpurun_out/prof_action9 python bench.py --n 256 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
