#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider 2>&1 | tee gpurun_out/r2_seventh_suite.log | tail -6
timeout 300 python benchmarks/run_configs.py --only "blocked_matrix_case" 2>&1 | cut -c1-300
python - <<'PY'
import sys, json
sys.path.insert(0, ".")
from firedrake_b200 import _lib
from benchmarks.run_configs import blocked_matrix_case
_lib.init(0)
print(json.dumps(blocked_matrix_case("config4 vector Helmholtz CG4 explicit, 16^3 (ncu target)", 16, 4, 3)))
PY
cat > /tmp/ncu16.py <<'PY'
import sys
sys.path.insert(0, ".")
from firedrake_b200 import _lib
from benchmarks.run_configs import blocked_matrix_case
_lib.init(0)
blocked_matrix_case("ncu", 16, 4, 3, steps=1)
PY
timeout 900 ncu --set full --import-source on -k regex:bdb_matrix_sym_kernel -c 1 -f -o gpurun_out/r02_bdb_sym_cg4_n16 python /tmp/ncu16.py > /dev/null 2>&1
ls -la gpurun_out/r02_bdb_sym_cg4_n16.ncu-rep
