#!/usr/bin/env python
"""Config 5: matrix-free Poisson CG_p on n^3 hexes, unpreconditioned CG across
N GPUs (slab partition, NCCL halos in every operator application, NCCL
all-reduce for the dot products).  Launch with torchrun; rank 0 prints JSON.

    python -m torch.distributed.run --nproc-per-node 4 benchmarks/cg_multi.py --size 128 --degree 5

The right-hand side is an analytic function of position (identical for every partition), so the
residual history is the parity signal: runs at 1/2/4/8 GPUs must produce the same history, and
``--host-baseline`` (1 process) adds the same fixed-iteration CG on the host cores with the
oracle's operator (bench.cpu_cg_baseline) -- the time to beat and an independent history.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from firedrake_b200 import _lib, op2                                                # noqa: E402
from firedrake_b200.assemble import DirichletBC, FunctionSpace, assemble, cg, interpolate, poisson  # noqa: E402
from firedrake_b200.halo import comm_init_from_env                                   # noqa: E402
from firedrake_b200.partition import SlabPartition                                   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", dest="n", type=int, default=128)
ap.add_argument("--degree", type=int, default=5)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--halo", default="exec", choices=["exec", "sum"])
ap.add_argument("--host-baseline", action="store_true")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
rank, world, dist = comm_init_from_env()
L = _lib.lib()
n, p = args.n, args.degree
part = SlabPartition(n, n, n, p, rank, world, warp=0.05, exec_halo=args.halo == "exec")
V = FunctionSpace(part.mesh, p, partition=part)
bcs = [DirichletBC(V, 0.0, ["bottom", "top"])]
A = assemble(poisson(V), bcs=bcs, mat_type="matfree")
# L = f*v*dx with an analytic f: the assembled load vector, as in demos/matrix_free/poisson.py.rst
from firedrake_b200.assemble import Form, OneFormAssembler                          # noqa: E402
f = interpolate(V, "sin(3.0 * x[0]) * cos(2.0 * x[1]) + x[2] * x[2] - 0.3 * x[0] * x[1]")
f.halo_valid = False
b = OneFormAssembler(Form(V, 0.0, 1.0), f).assemble()
bcs[0].zero(b)
x = V.dat()
scratch = op2.DeviceArray(8)
buf = np.zeros(1)


def allreduce(v):
    buf[0] = v
    _lib.check(L.fdb_memcpy_h2d(scratch.ptr, buf.ctypes.data, 8))
    _lib.check(L.fdb_allreduce(scratch.ptr, 1, 0))
    scratch.to_host(buf)
    return float(buf[0])


red = allreduce if world > 1 else None
cg(A, b, x, rtol=0.0, maxit=2, allreduce=red)
_lib.check(L.fdb_synchronize())
if dist is not None:
    dist.barrier()
ts = []
for rep in range(args.reps):             # identical solves: the spread is host / launch jitter
    x.zero()
    x.device_ptr
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    its, hist = cg(A, b, x, rtol=0.0, maxit=args.iters, allreduce=red)
    _lib.check(L.fdb_synchronize())
    if dist is not None:
        dist.barrier()
    ts.append(time.perf_counter() - t0)
t = min(ts)
no = V.V.owned_node_count
xx = float(np.dot(x.data_ro[:no].ravel(), x.data_ro[:no].ravel()))
xnorm = float(np.sqrt(red(xx) if red else xx))
if rank == 0:
    ndof = (n * p + 1) ** 3
    out = {"case": f"config5 Poisson CG{p} matrix-free CG, {n}^3, {world} GPU(s)", "halo": args.halo,
           "dofs": ndof, "iterations": its, "s_per_iteration": t / its, "timing": f"best of {args.reps} solves",
           "s_per_iteration_all": [v / its for v in ts],
           "dof_iterations_per_s": ndof * its / t,
           "residual_reduction": hist[-1] / hist[0], "residual_history": [float(h) for h in hist],
           "solution_norm": xnorm}
    if args.host_baseline and world == 1:
        import bench
        bh = b.data_ro_with_halos.copy()
        th, hh = bench.cpu_cg_baseline(part.mesh, V.V, p, bh, bcs[0].nodes, args.iters)
        out["host_cg"] = {"seconds": th, "s_per_iteration": th / args.iters, "cores": len(os.sched_getaffinity(0)),
                          "residual_history": hh,
                          "max_rel_history_diff": float(max(abs(a - c) / c for a, c in zip(hist, hh)))}
        out["speedup_vs_host_cg"] = th / t
    print(json.dumps(out))
