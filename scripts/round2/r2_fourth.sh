#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider -x 2>&1 | tee gpurun_out/r2_fourth_suite.log | tail -6
timeout 300 python benchmarks/run_configs.py --only "Poisson CG1 action" 2>&1 | cut -c1-260
FDB_Q1_THREAD=0 timeout 300 python benchmarks/run_configs.py --only "Poisson CG1 action" 2>&1 | cut -c1-260
timeout 300 python benchmarks/run_configs.py --only "dg_case" 2>&1 | cut -c1-300
FDB_DG_GENERIC=1 timeout 300 python benchmarks/run_configs.py --only "fused=True" 2>&1 | cut -c1-300
timeout 300 python benchmarks/run_configs.py --only "blocked_matrix_case" 2>&1 | cut -c1-300
timeout 600 python benchmarks/cg_multi.py --size 128 --degree 5 --host-baseline 2>/dev/null | tail -1 > gpurun_out/r2_cg_1.json; cut -c1-400 gpurun_out/r2_cg_1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_cg_1.json')); print({k:v for k,v in d['host_cg'].items() if k!='residual_history'}, d['speedup_vs_host_cg'])"
