#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import sys, json
sys.path.insert(0, '.')
sys.argv=['x']
import benchmarks.run_configs as rc
from firedrake_b200 import _lib
_lib.init(0)
for n,p in [(128,1),(48,2),(32,3),(64,3)]:
    try:
        print(json.dumps(rc.matrix_case(f"Poisson CG{p} matrix", n, p)), flush=True)
    except Exception as e:
        print("ERR", repr(e)[:300])
PY
