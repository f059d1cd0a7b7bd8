#!/bin/bash
# session 2, last call: full GPU suite + default bench line at HEAD, NOROT experiment build, CG1 / CG2 actions
mkdir -p gpurun_out
timeout 400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s2_final_suite.txt
( time timeout 300 python bench.py --steps 10 --warmup 3 ) > gpurun_out/s2_final_bench.json 2> gpurun_out/s2_final_bench.err; tail -c 2600 gpurun_out/s2_final_bench.json; tail -4 gpurun_out/s2_final_bench.err
FDB200_LIB=firedrake_b200/lib/variants/libfdb200_norot.so timeout 120 python tools/time_action.py --check 2>&1 | tail -1 | tee gpurun_out/s2_final_norot.json
timeout 200 ncu --set full --import-source on --clock-control none -k regex:helmholtz_action_kernel -s 3 -c 1 -f -o gpurun_out/r02s2_action_cg3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > /dev/null 2>&1
ls -la gpurun_out/r02s2_action_cg3.ncu-rep
timeout 120 python benchmarks/run_configs.py --only "Poisson CG1 action" 2>&1 | tail -1 | tee gpurun_out/s2_final_cg12.jsonl
timeout 120 python benchmarks/run_configs.py --only "Poisson CG2 action" 2>&1 | tail -1 | tee -a gpurun_out/s2_final_cg12.jsonl
