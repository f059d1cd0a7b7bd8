#!/usr/bin/env python
"""Config 5: matrix-free Poisson CG_p on n^3 hexes, unpreconditioned CG across
N GPUs (slab partition, NCCL halos in every operator application, NCCL
all-reduce for the dot products).  Launch with torchrun; rank 0 prints JSON.

    python -m torch.distributed.run --nproc-per-node 4 benchmarks/cg_multi.py --size 128 --degree 5
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
from firedrake_b200.assemble import DirichletBC, FunctionSpace, assemble, cg, poisson  # noqa: E402
from firedrake_b200.halo import comm_init_from_env                                   # noqa: E402
from firedrake_b200.partition import SlabPartition                                   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", dest="n", type=int, default=128)
ap.add_argument("--degree", type=int, default=5)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
rank, world, dist = comm_init_from_env()
L = _lib.lib()
n, p = args.n, args.degree
part = SlabPartition(n, n, n, p, rank, world, warp=0.05)
V = FunctionSpace(part.mesh, p, partition=part)
bcs = [DirichletBC(V, 0.0, ["bottom", "top"])]
A = assemble(poisson(V), bcs=bcs, mat_type="matfree")
b = V.dat(np.random.default_rng(rank).standard_normal(V.node_count))
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
x.zero()
t0 = time.perf_counter()
its, hist = cg(A, b, x, rtol=0.0, maxit=args.iters, allreduce=red)
_lib.check(L.fdb_synchronize())
if dist is not None:
    dist.barrier()
t = time.perf_counter() - t0
if rank == 0:
    ndof = (n * p + 1) ** 3
    print(json.dumps({"case": f"config5 Poisson CG{p} matrix-free CG, {n}^3, {world} GPU(s)",
                      "dofs": ndof, "iterations": its, "s_per_iteration": t / its,
                      "dof_iterations_per_s": ndof * its / t,
                      "residual_reduction": hist[-1] / hist[0]}))
