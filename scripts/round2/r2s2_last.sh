#!/bin/bash
# session 2, leftover GPU seconds: launch list of the final bench command; stash stride 30 (no A1 bank conflicts)
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s2_launches_bench_n256.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r02s2_launches_bench.out 2>&1
tail -c 300 gpurun_out/r02s2_launches_bench.out; wc -l gpurun_out/r02s2_launches_bench_n256.csv
FDB200_LIB=firedrake_b200/lib/variants/libfdb200_gs30.so timeout 60 python tools/time_action.py --check 2>&1 | tail -1 | tee gpurun_out/s2_last_gs30.json
