#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (not the driver's
bench line): one JSON object per line on stdout.

    python benchmarks/run_configs.py [--quick]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from firedrake_b200 import _lib, op2                                   # noqa: E402
from firedrake_b200.assemble import (DirichletBC, FunctionSpace, assemble, cg, helmholtz,  # noqa: E402
                                     poisson, OneFormAssembler, Form)
from firedrake_b200.utility_meshes import ExtrudedHexMesh              # noqa: E402


def timed(fn, steps=5, warm=2):
    L = _lib.lib()
    for _ in range(warm):
        fn()
    t = C.c_void_p()
    L.fdb_timer_create(C.byref(t))
    ms = C.c_float()
    _lib.check(L.fdb_synchronize())
    L.fdb_timer_start(t)
    for _ in range(steps):
        fn()
    L.fdb_timer_stop(t, C.byref(ms))
    L.fdb_timer_destroy(t)
    return ms.value / steps


def action_case(name, n, p, cdim=1, alpha=1.0, beta=0.0, warp=0.05, permute=None, steps=5):
    mesh = ExtrudedHexMesh(n, n, n, warp=warp, permute_seed=permute)
    V = FunctionSpace(mesh, p, cdim=cdim)
    shape = (V.node_count,) if cdim == 1 else (V.node_count, cdim)
    u = V.dat(np.random.default_rng(0).standard_normal(shape))
    y = V.dat()
    asm = OneFormAssembler(Form(V, alpha, beta), u)
    ms = timed(lambda: asm.assemble(y), steps)
    ndof = V.node_count * cdim
    return {"case": name, "n": n, "degree": p, "cdim": cdim, "cells": mesh.num_cells, "dofs": ndof,
            "ms": ms, "dofs_per_s": ndof / ms * 1e3, "permute": permute, "warp": warp}


def matrix_case(name, n, p, steps=3):
    mesh = ExtrudedHexMesh(n, n, n, warp=0.05)
    V = FunctionSpace(mesh, p)
    t0 = time.perf_counter()
    mat = op2.Mat(op2.Sparsity((V.node_set, V.node_set), [(V.cell_node_map, V.cell_node_map, None)]))
    _lib.check(_lib.lib().fdb_synchronize())
    t_sparsity = time.perf_counter() - t0
    a = poisson(V)
    ms = timed(lambda: assemble(a, tensor=mat), steps, warm=1)
    return {"case": name, "n": n, "degree": p, "dofs": V.node_count, "nnz": mat.nnz,
            "sparsity_s": t_sparsity, "assemble_ms": ms, "dofs_per_s": V.node_count / ms * 1e3,
            "csr_GB": mat.nnz * 12 / 1e9}


def blocked_matrix_case(name, n, p, cdim, steps=2):
    """config 4, explicit: vector-valued space -> blocked CSR."""
    mesh = ExtrudedHexMesh(n, n, n, warp=0.05)
    V = FunctionSpace(mesh, p, cdim=cdim)
    t0 = time.perf_counter()
    mat = op2.Mat(op2.Sparsity((V.dof_dset, V.dof_dset), [(V.cell_node_map, V.cell_node_map, None)]))
    _lib.check(_lib.lib().fdb_synchronize())
    t_sparsity = time.perf_counter() - t0
    a = helmholtz(V)
    ms = timed(lambda: assemble(a, tensor=mat), steps, warm=1)
    return {"case": name, "n": n, "degree": p, "cdim": cdim, "dofs": V.node_count * cdim, "block_nnz": mat.nnz,
            "sparsity_s": t_sparsity, "assemble_ms": ms, "dofs_per_s": V.node_count * cdim / ms * 1e3,
            "baij_GB": mat.nnz * (8 * cdim * cdim + 4) / 1e9}


def generic_vs_fast_case(name, n, steps=5):
    """The generic NVRTC wrapper around a plain-C Q1 Poisson kernel against the hand-written
    kernel on the same problem."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_codegen as tc
    mesh = ExtrudedHexMesh(n, n, n, warp=0.05)
    V = FunctionSpace(mesh, 1)
    u = V.dat(np.random.default_rng(0).standard_normal(V.node_count))
    yg, yf = V.dat(), V.dat()
    k = op2.Kernel(tc.Q1_POISSON, "q1_poisson")

    def generic():
        yg.zero()
        yg.device_ptr
        op2.par_loop(k, V.cell_set, yg(op2.INC, V.cell_node_map), V.coordinates(op2.READ, V.coord_map),
                     u(op2.READ, V.cell_node_map))
    asm = OneFormAssembler(poisson(V), u)
    ms_g = timed(generic, steps)
    ms_f = timed(lambda: asm.assemble(yf), steps)
    err = float(np.abs(yg.data_ro - yf.data_ro).max() / np.abs(yf.data_ro).max())
    return {"case": name, "n": n, "dofs": V.node_count, "generic_ms": ms_g, "fast_ms": ms_f,
            "ratio": ms_g / ms_f, "rel_diff": err}


def cg_case(name, n, p, iters=20):
    mesh = ExtrudedHexMesh(n, n, n, warp=0.05)
    V = FunctionSpace(mesh, p)
    bcs = [DirichletBC(V, 0.0, ["bottom", "top"])]
    A = assemble(poisson(V), bcs=bcs, mat_type="matfree")
    b = V.dat(np.random.default_rng(1).standard_normal(V.node_count))
    bcs[0].zero(b)
    x = V.dat()
    cg(A, b, x, rtol=0.0, maxit=2)
    _lib.check(_lib.lib().fdb_synchronize())
    x.zero()
    t0 = time.perf_counter()
    its, hist = cg(A, b, x, rtol=0.0, maxit=iters)
    _lib.check(_lib.lib().fdb_synchronize())
    t = time.perf_counter() - t0
    return {"case": name, "n": n, "degree": p, "dofs": V.node_count, "iterations": its,
            "s_per_iteration": t / its, "dof_iterations_per_s": V.node_count * its / t,
            "residual_reduction": hist[-1] / hist[0]}


def dg_case(name, n, steps=10, fused=False):
    from firedrake_b200.assemble import DGAdvection
    from firedrake_b200.utility_meshes import QuadMesh
    m = QuadMesh(n, n)
    X = m.coordinates
    u = np.stack([0.5 - X[:, 1], X[:, 0] - 0.5], axis=1)
    prob = DGAdvection(m, dt=2 * np.pi / 600 * 40 / n, q_in=1.0, fused=fused)
    q = prob.function(1.0 + np.random.default_rng(0).random(m.num_cells * 4))
    uu = prob.velocity(u)
    out = prob.function()
    ms = timed(lambda: prob.assemble(q, uu, out), steps)
    ndof = m.num_cells * 4
    # algorithmic bytes: read q + write out (16 B/dof), coordinates + velocity (32 B/vertex),
    # cell maps (32 B/cell) and facet maps (2*(32+32)+8 B per interior facet)
    nbytes = 16 * ndof + 32 * m.node_count + 32 * m.num_cells + 136 * len(m.int_facet_cells)
    return {"case": name, "n": n, "cells": m.num_cells, "dofs": ndof, "ms": ms,
            "dofs_per_s": ndof / ms * 1e3, "algorithmic_GBps": nbytes / ms / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None, help="run only the cases whose name contains this substring")
    args = ap.parse_args()
    _lib.init(0)
    q = args.quick
    jobs = [
        lambda: action_case("config2 Poisson CG3 action, lexicographic", 64 if q else 256, 3),
        lambda: action_case("config2 Poisson CG3 action, random base-cell order", 64 if q else 256, 3, permute=0),
        lambda: action_case("config2 Poisson CG3 action, affine mesh (warp 0)", 64 if q else 256, 3, warp=0.0),
        lambda: action_case("Poisson CG1 action", 64 if q else 256, 1),
        lambda: action_case("Poisson CG2 action", 64 if q else 256, 2),
        lambda: action_case("config4 vector Helmholtz CG4 action (cdim 3)", 16 if q else 64, 4, cdim=3, beta=1.0),
        lambda: action_case("config5 Poisson CG5 action", 32 if q else 128, 5),
        lambda: cg_case("config5 Poisson CG5 matrix-free CG", 32 if q else 128, 5, iters=10 if q else 20),
        lambda: dg_case("config3 DG advection DQ1 RHS (cell + ext + int facet kernels)", 256 if q else 2048),
        lambda: dg_case("config3 DG advection DQ1 RHS (fused owner-computes kernel)", 256 if q else 2048, fused=True),
        lambda: matrix_case("Poisson CG1 matrix", 32 if q else 128, 1),
        lambda: matrix_case("Poisson CG2 matrix", 16 if q else 48, 2),
        lambda: matrix_case("Poisson CG3 matrix", 8 if q else 32, 3),
    ]
    jobs += [
        lambda: matrix_case("Poisson CG4 matrix (new instantiation)", 8 if q else 24, 4),
        lambda: blocked_matrix_case("config4 vector Helmholtz CG4 explicit (blocked CSR)", 8 if q else 32, 4, 3),
        lambda: generic_vs_fast_case("generic NVRTC wrapper vs hand-written kernel, Poisson CG1", 32 if q else 128),
    ]
    if args.only:
        import inspect
        jobs = [j for j in jobs if args.only in inspect.getsource(j)]
    for j in jobs:
        try:
            print(json.dumps(j()), flush=True)
        except Exception as e:                       # keep going: report the failure
            print(json.dumps({"error": repr(e)[:300]}), flush=True)
        _lib.lib().fdb_mirror_drop_all()


if __name__ == "__main__":
    main()
