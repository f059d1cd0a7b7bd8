#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python benchmarks/run_configs.py > gpurun_out/configs_r01c.jsonl 2> gpurun_out/configs.err
grep -E "DG|matrix|CG5" gpurun_out/configs_r01c.jsonl | cut -c1-330
tail -2 gpurun_out/configs.err
