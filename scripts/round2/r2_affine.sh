#!/bin/bash
timeout 600 python -m pytest tests/test_action_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
python - <<'PY'
import sys, json
sys.path.insert(0, ".")
from firedrake_b200 import _lib
from benchmarks.run_configs import action_case
_lib.init(0)
for p in (1, 2):
    r = action_case(f"Poisson CG{p} action, affine mesh (warp 0)", 256, p, warp=0.0)
    print(json.dumps({k: r[k] for k in ("case", "ms", "dofs_per_s")}), flush=True)
    _lib.lib().fdb_mirror_drop_all()
PY
