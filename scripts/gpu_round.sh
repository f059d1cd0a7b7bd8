#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python benchmarks/run_configs.py > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs.err
cat gpurun_out/configs_r01.jsonl | cut -c1-420
tail -3 gpurun_out/configs.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -12 gpurun_out/launches_r01.csv | cut -c1-300
