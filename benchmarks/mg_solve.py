#!/usr/bin/env python
"""Time-to-solution of the Poisson problem of config 2 / 5 with and without geometric multigrid:
matrix-free CG_p on a (coarse * 2^levels)^3 warped hex mesh, Dirichlet on bottom and top,
PCG preconditioned by one V-cycle (firedrake_b200/mg.py) against plain CG.  Single GPU or
torchrun (slab partition; the coarse x-resolution must be divisible by the number of ranks).

NOT YET RUN ON A GPU (written after round 1's GPU budget was spent; the V-cycle is verified on the
mock engine only).

    python benchmarks/mg_solve.py --coarse 16 --levels 3 --degree 3          # 128^3
    python -m torch.distributed.run --nproc-per-node 4 benchmarks/mg_solve.py --coarse 16 --levels 4
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from firedrake_b200 import _lib, mg, op2                                  # noqa: E402
from firedrake_b200.assemble import cg, poisson                           # noqa: E402
from firedrake_b200.halo import comm_init_from_env                        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--coarse", type=int, default=16)
ap.add_argument("--levels", type=int, default=3)
ap.add_argument("--degree", type=int, default=3)
ap.add_argument("--rtol", type=float, default=1e-8)
ap.add_argument("--plain-maxit", type=int, default=3000)
ap.add_argument("--no-plain", action="store_true")
args = ap.parse_args()
rank, world, dist = comm_init_from_env()
L = _lib.lib()
scratch = op2.DeviceArray(8)
buf = np.zeros(1)


def allreduce(v):
    buf[0] = v
    _lib.check(L.fdb_memcpy_h2d(scratch.ptr, buf.ctypes.data, 8))
    _lib.check(L.fdb_allreduce(scratch.ptr, 1, 0))
    scratch.to_host(buf)
    return float(buf[0])


red = allreduce if world > 1 else None
t0 = time.perf_counter()
h = mg.MeshHierarchy(args.coarse, args.coarse, args.coarse, args.levels, rank=rank, nranks=world, warp=0.05)
vc = mg.VCycle(h, args.degree, poisson, bc_domains=("bottom", "top"), allreduce=red)
top = len(h) - 1
V, A = vc.spaces[top], vc.ops[top]
b = V.dat(np.random.default_rng(rank).standard_normal(V.node_count))
for bc in vc.bcs[top]:
    bc.zero(b)
setup_s = time.perf_counter() - t0


def solve(fn):
    x = V.dat()
    x.device_ptr
    _lib.check(L.fdb_synchronize())
    if dist is not None:
        dist.barrier()
    t = time.perf_counter()
    its, hist = fn(x)
    _lib.check(L.fdb_synchronize())
    return its, hist[-1] / hist[0], time.perf_counter() - t


solve(lambda x: mg.pcg(A, b, x, lambda r, z: vc.apply(top, r, z), rtol=1e-2, allreduce=red))      # warm-up / JIT
n_mg, res_mg, t_mg = solve(lambda x: mg.pcg(A, b, x, lambda r, z: vc.apply(top, r, z), rtol=args.rtol,
                                            allreduce=red))
out = {"workload": f"Poisson CG{args.degree} on {args.coarse << args.levels}^3 hexes, {args.levels + 1} levels",
       "n_gpus": world, "setup_s": setup_s, "mg_pcg": {"iterations": n_mg, "rel_residual": res_mg, "seconds": t_mg}}
if not args.no_plain:
    n_cg, res_cg, t_cg = solve(lambda x: cg(A, b, x, rtol=args.rtol, maxit=args.plain_maxit, allreduce=red))
    out["plain_cg"] = {"iterations": n_cg, "rel_residual": res_cg, "seconds": t_cg}
if rank == 0:
    print(json.dumps(out))
