"""torchrun worker: distributed matrix-free CG (config 5 style) -- halo
exchanges inside every operator application, NCCL all-reduce for the dot
products, Dirichlet rows -- against a serial SciPy solve of the oracle matrix."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from firedrake_b200 import _lib, op2                                     # noqa: E402
from firedrake_b200.assemble import DirichletBC, FunctionSpace, assemble, cg, helmholtz  # noqa: E402
from firedrake_b200.fiat_lite import interval_element                    # noqa: E402
from firedrake_b200.halo import comm_init_from_env                       # noqa: E402
from firedrake_b200.partition import SlabPartition                       # noqa: E402
from firedrake_b200.utility_meshes import ExtrudedHexMesh                # noqa: E402
from oracle import oracle                                                # noqa: E402

rank, world, dist = comm_init_from_env()
L = _lib.lib()
p, (nx, ny, nz) = 2, (6, 4, 5)
part = SlabPartition(nx, ny, nz, p, rank, world, warp=0.05)
V = FunctionSpace(part.mesh, p, partition=part)
bcs = [DirichletBC(V, 0.0, ["bottom", "top"])]
A = assemble(helmholtz(V), bcs=bcs, mat_type="matfree")
lat = V.V.dof_lattice()
f = lambda Lq: np.cos(0.3 * Lq[:, 0]) * (1 + 0.1 * Lq[:, 1]) - 0.02 * Lq[:, 2] ** 2
b = V.dat(f(lat))
bcs[0].zero(b)
x = V.dat()
scratch = op2.DeviceArray(8)


def allreduce(v):
    _lib.check(L.fdb_memcpy_h2d(scratch.ptr, np.array([v]).ctypes.data, 8))
    _lib.check(L.fdb_allreduce(scratch.ptr, 1, 0))
    out = np.zeros(1)
    scratch.to_host(out)
    return float(out[0])


its, hist = cg(A, b, x, rtol=1e-12, maxit=400, allreduce=allreduce if world > 1 else None)
# serial reference: oracle matrix with masked rows/cols + unit diagonal, SciPy direct solve
gm = ExtrudedHexMesh(nx, ny, nz, warp=0.05)
gV = gm.function_space(p)
glat = gV.dof_lattice()
bn = np.union1d(gV.boundary_nodes("bottom"), gV.boundary_nodes("top"))
lg = np.arange(gV.node_count, dtype=np.int32)
lg[bn] = -1
ro, co = oracle.build_sparsity(gV.node_count, gV.cell_node_map, gV.offset, gm.nz)
va = np.zeros(len(co))
oracle.matrix_extruded(interval_element(p), 0, gm.num_base_cells, [0, gm.layers], ro, co, va,
                       gm.coordinates, gV.cell_node_map, gV.offset, gm.coord_map, gm.coord_offset,
                       lg, lg, 1.0, 1.0)
import scipy.sparse as sp
import scipy.sparse.linalg as spla
Ag = sp.csr_matrix((va, co, ro), shape=(gV.node_count,) * 2).tolil()
for r_ in bn:
    Ag[r_, r_] = 1.0
bg = f(glat)
bg[bn] = 0.0
xs = spla.spsolve(Ag.tocsc(), bg)
key = lambda Lq: (Lq[:, 0] * 1000 + Lq[:, 1]) * 1000 + Lq[:, 2]
lookup = dict(zip(key(glat).tolist(), xs.tolist()))
no = V.V.owned_node_count
ref = np.array([lookup[k] for k in key(lat[:no]).tolist()])
err = np.abs(x.data_ro[:no] - ref).max() / np.abs(xs).max()
print(f"rank {rank}/{world}: CG iterations {its}, residual {hist[-1] / hist[0]:.1e}, err vs serial {err:.2e}")
assert err < 1e-9 and hist[-1] < 1e-11 * hist[0]
if dist is not None:
    dist.barrier()
print("CG_OK")
