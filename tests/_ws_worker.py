"""Worker of tests/test_action_gpu.py::test_warp_specialised_kernel: run with FDB_WS set (the engine reads it
once per process).  Compares the warp-specialised degree-3 action kernel with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from firedrake_b200 import op2                              # noqa: E402
from firedrake_b200.utility_meshes import ExtrudedHexMesh   # noqa: E402
from oracle import oracle                                   # noqa: E402
import test_action_gpu as T                                 # noqa: E402

assert int(os.environ.get("FDB_WS", "0")) > 0
oracle.build()
p = 3
for (nx, ny, nz, alpha, beta) in [(6, 6, 19, 1.0, 0.0), (3, 2, 9, 1.0, 2.0), (13, 11, 40, 0.5, 1.0)]:
    mesh = ExtrudedHexMesh(nx, ny, nz, warp=0.05, permute_seed=1)
    V, cells, m0, m1, x, y, X = T.build(mesh, p)
    k = op2.Kernel("helmholtz", degree=p, alpha=alpha, beta=beta)
    op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    yo = T.oracle_action(oracle, mesh, V, p, x.data_ro, alpha=alpha, beta=beta)
    err = T.relerr(y.data_ro, yo)
    assert err < T.TOL, (nx, ny, nz, err)
# a subset of columns (collist path of the movers)
mesh = ExtrudedHexMesh(5, 5, 12, warp=0.02)
V, cells, m0, m1, x, y, X = T.build(mesh, p)
idx = np.array([0, 3, 4, 11, 17, 24], dtype=np.int32)
k = op2.Kernel("helmholtz", degree=p)
op2.par_loop(k, op2.Subset(cells, idx), y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
yo = np.zeros(V.node_count)
for c in idx:
    yo += T.oracle_action(oracle, mesh, V, p, x.data_ro, start=int(c), end=int(c) + 1)
assert T.relerr(y.data_ro, yo) < T.TOL
print("WS_OK")
