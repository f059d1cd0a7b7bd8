#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider -x 2>&1 | tee gpurun_out/r2_second_suite.log | tail -15
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 1500 gpurun_out/r2_bench1.json; tail -3 gpurun_out/r2_bench1.err
timeout 300 python benchmarks/run_configs.py --only matrix > gpurun_out/r2_matrix_dmma.jsonl 2>&1; cat gpurun_out/r2_matrix_dmma.jsonl
FDB_MATRIX_DMMA=0 timeout 300 python benchmarks/run_configs.py --only matrix > gpurun_out/r2_matrix_old.jsonl 2>&1; cat gpurun_out/r2_matrix_old.jsonl
timeout 600 ncu --set full --import-source on -k regex:bdb_matrix -c 1 -f -o gpurun_out/r02_bdb_cg4 python benchmarks/run_configs.py --only "blocked_matrix_case" --quick > gpurun_out/r2_ncu_bdb.log 2>&1; tail -3 gpurun_out/r2_ncu_bdb.log
