"""torchrun worker: DG advection RHS on a slab-partitioned quad mesh with a
ghost-cell halo (config 3: "1 -> 8 GPU halo exchange"), fused owner-computes
kernel, against the serial oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from firedrake_b200.assemble import DGAdvection, dg_slab          # noqa: E402
from firedrake_b200.halo import Halo, comm_init_from_env           # noqa: E402
from firedrake_b200.utility_meshes import QuadMesh                  # noqa: E402
from oracle import oracle                                           # noqa: E402

rank, world, dist = comm_init_from_env()
nx, ny = 13, 7
mesh, neigh = dg_slab(nx, ny, rank, world)
halo = Halo(neigh)
X = mesh.coordinates
vel = lambda P: np.stack([0.5 - P[:, 1], P[:, 0] - 0.5], axis=1)
qfun = lambda gcol, iy, k: 1.0 + 0.3 * np.sin(0.7 * gcol + 0.4 * iy) + 0.05 * k
gcol = mesh.cell_global_column
iy = np.concatenate([np.arange(ny)] * (mesh.num_cells // ny))
# local cell -> (global column, iy): rebuild iy from the coordinate map
y0 = X[mesh.coord_map[:, 0], 1]
iy = np.rint(y0 * ny).astype(int)
q = np.stack([qfun(gcol, iy, k) for k in range(4)], axis=1).ravel()
q[4 * mesh.num_owned_cells:] = 1e30                     # stale ghosts
prob = DGAdvection(mesh, dt=0.02, q_in=1.0, fused=True, halo=halo)
qd = prob.function(q)
qd.halo_valid = False
out = prob.assemble(qd, prob.velocity(vel(X)))
# serial reference
gm = QuadMesh(nx, ny)
gi = np.arange(gm.num_cells) // ny
gj = np.arange(gm.num_cells) % ny
gq = np.stack([qfun(gi, gj, k) for k in range(4)], axis=1).ravel()
ref = oracle.dg_rhs(gm, gq, vel(gm.coordinates), dt=0.02, q_in=1.0).reshape(-1, 4)
mine = out.data_ro[:4 * mesh.num_owned_cells].reshape(-1, 4)
own = slice(0, mesh.num_owned_cells)
want = ref[gcol[own] * ny + iy[own]]
err = np.abs(mine - want).max() / np.abs(ref).max()
print(f"rank {rank}/{world}: DG RHS err vs serial {err:.2e}")
assert err < 1e-12
if dist is not None:
    dist.barrier()
print("DG_OK")
