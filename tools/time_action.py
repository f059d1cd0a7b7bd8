#!/usr/bin/env python
"""A/B timing of the action kernel of one libfdb200 build (select it with FDB200_LIB=...):

    FDB200_LIB=firedrake_b200/lib/variants/libfdb200_X.so python tools/time_action.py [--n 256] [--p 3] [--check]

Prints one JSON line: kernel-only milliseconds (no memset of y: the launch accumulates into y), and with
--check the relative error against the oracle on a small warped mesh (same library)."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from firedrake_b200 import _lib, op2                                   # noqa: E402
from firedrake_b200.fiat_lite import interval_element                  # noqa: E402
from firedrake_b200.utility_meshes import ExtrudedHexMesh              # noqa: E402


def setup(n, p, nz=None, seed=0):
    mesh = ExtrudedHexMesh(n, n, nz or n, warp=0.05)
    V = mesh.function_space(p)
    cells = op2.ExtrudedSet(op2.Set(mesh.num_base_cells), mesh.layers)
    nodes = op2.Set(V.node_count)
    vnodes = op2.Set(mesh.coord_space.node_count)
    m0 = op2.Map(cells, nodes, V.arity, V.cell_node_map, offset=V.offset)
    m1 = op2.Map(cells, vnodes, 8, mesh.coord_map, offset=mesh.coord_offset)
    x = op2.Dat(nodes, np.random.default_rng(seed).standard_normal(V.node_count))
    y = op2.Dat(nodes)
    X = op2.Dat(op2.DataSet(vnodes, 3), mesh.coordinates)
    k = op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=0.0)
    return mesh, V, cells, m0, m1, x, y, X, k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--p", type=int, default=3)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--check-only", action="store_true")
    a = ap.parse_args()
    out = {"lib": os.path.basename(_lib.LIB_PATH), "n": a.n, "p": a.p,
           "env": {k: v for k, v in os.environ.items() if k.startswith("FDB_")}}
    if a.check or a.check_only:
        from oracle import oracle
        mesh, V, cells, m0, m1, x, y, X, k = setup(6, a.p, nz=19, seed=3)
        op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
        yo = np.zeros(V.node_count)
        oracle.action_extruded(interval_element(a.p), 0, mesh.num_base_cells, [0, mesh.layers], yo, mesh.coordinates,
                               x.data_ro.copy(), V.cell_node_map, V.offset, mesh.coord_map, mesh.coord_offset)
        out["rel_err"] = float(np.abs(y.data_ro - yo).max() / np.abs(yo).max())
    if a.check_only:
        print(json.dumps(out), flush=True)
        return
    mesh, V, cells, m0, m1, x, y, X, k = setup(a.n, a.p)
    L = _lib.lib()

    def step():
        op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    for _ in range(3):
        step()
    t = C.c_void_p()
    L.fdb_timer_create(C.byref(t))
    ms = C.c_float()
    best = 1e30
    for rep in range(3):
        _lib.check(L.fdb_synchronize())
        L.fdb_timer_start(t)
        for _ in range(a.steps):
            step()
        L.fdb_timer_stop(t, C.byref(ms))
        best = min(best, ms.value / a.steps)
    out["kernel_ms"] = best
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
