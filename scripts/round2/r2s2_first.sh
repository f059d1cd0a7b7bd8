#!/bin/bash
# session 2, call 1: issue-slot microbenchmark, full GPU suite at HEAD, default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 120 tools/_microbench_issue > gpurun_out/s2_microbench_issue.txt 2>&1; cat gpurun_out/s2_microbench_issue.txt
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s2_gpu_suite.txt
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; tail -c 3000 gpurun_out/s2_bench.json; tail -4 gpurun_out/s2_bench.err
