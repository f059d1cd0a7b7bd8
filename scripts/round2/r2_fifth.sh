#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_matrix_gpu.py tests/test_jit_gpu.py -q -m gpu -p no:cacheprovider -x -k "matrix" 2>&1 | tail -3
python - <<'PY'
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
from firedrake_b200 import _lib, op2
from benchmarks.run_configs import matrix_case
L = _lib.init(0)
for mode in (0, 1):
    _lib.check(L.fdb_set_option(b"matrix_kernel", mode))
    for n, p in ((64, 3), (24, 4), (48, 2)):
        try:
            r = matrix_case(f"Poisson CG{p} matrix kernel={mode}", n, p)
            print(json.dumps({k: r[k] for k in ("case", "n", "assemble_ms", "nnz")}), flush=True)
        except Exception as e:
            print("failed", mode, n, p, repr(e)[:200])
        L.fdb_mirror_drop_all()
PY
# ncu captures for profiles/
timeout 900 ncu --set full --import-source on -k regex:bdb_matrix_sym_kernel -c 1 -f -o gpurun_out/r02_bdb_sym_cg4 python benchmarks/run_configs.py --only "blocked_matrix_case" --quick > /dev/null 2>&1
timeout 900 ncu --set full --import-source on -k regex:q1_action -c 1 -f -o gpurun_out/r02_q1_action python benchmarks/run_configs.py --only "Poisson CG1 action" > /dev/null 2>&1
timeout 900 ncu --set full --import-source on -k regex:dg_fused_nq -c 1 -f -o gpurun_out/r02_dg_fused python benchmarks/run_configs.py --only "fused=True" > /dev/null 2>&1
timeout 900 ncu --set full --import-source on -k regex:helmholtz_action_kernel -s 3 -c 1 -f -o gpurun_out/r02_action_cg3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n256.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r02_launches_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
