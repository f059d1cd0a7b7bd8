"""torchrun worker: distributed device action (NCCL halos) vs the serial oracle.
Run by tests/test_halo_gpu.py with one process per GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from firedrake_b200 import op2                              # noqa: E402
from firedrake_b200.fiat_lite import interval_element       # noqa: E402
from firedrake_b200.halo import Halo, comm_init_from_env    # noqa: E402
from firedrake_b200.partition import SlabPartition          # noqa: E402
from firedrake_b200.utility_meshes import ExtrudedHexMesh   # noqa: E402
from oracle import oracle                                   # noqa: E402

rank, world, dist = comm_init_from_env()
worst = 0.0
# (degree, mesh, exec-halo partition?): the reference protocol (ghost-sum reduce) and the
# owner-computes one (redundant exec-halo column, no reduce; SURVEY.md section 8e option (ii))
for p, (nx, ny, nz), exec_halo in [(1, (6, 4, 5), False), (3, (7, 5, 9), False), (2, (9, 3, 4), False),
                                   (3, (7, 5, 9), True), (1, (6, 4, 5), True), (4, (5, 3, 4), True)]:
    part = SlabPartition(nx, ny, nz, p, rank, world, warp=0.05, exec_halo=exec_halo)
    mesh, V = part.mesh, part.V
    halo = Halo(part.neighbours)
    cells = op2.ExtrudedSet(op2.Set(part.cell_sizes), mesh.layers)
    cells.owner_computes = part.exec_halo
    nodes = op2.Set(part.node_sizes)
    vnodes = op2.Set(mesh.coord_space.node_count)
    m0 = op2.Map(cells, nodes, V.arity, V.cell_node_map, offset=V.offset)
    m1 = op2.Map(cells, vnodes, 8, mesh.coord_map, offset=mesh.coord_offset)
    dn = op2.DataSet(nodes, 1, halo=halo)
    lat = V.dof_lattice()
    f = lambda Lq: np.sin(0.37 * Lq[:, 0]) + 0.11 * Lq[:, 1] * Lq[:, 2] - 0.05 * Lq[:, 0] * Lq[:, 2]
    xv = f(lat)
    xv[V.owned_node_count:] = 1e30            # stale ghosts: the exchange must repair them
    x = op2.Dat(dn, xv)
    x.halo_valid = False
    y = op2.Dat(dn)
    X = op2.Dat(op2.DataSet(vnodes, 3), mesh.coordinates)
    k = op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=0.5)
    op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    gm = ExtrudedHexMesh(nx, ny, nz, warp=0.05)
    gV = gm.function_space(p)
    glat = gV.dof_lattice()
    gy = np.zeros(gV.node_count)
    oracle.action_extruded(interval_element(p), 0, gm.num_base_cells, [0, gm.layers], gy,
                           gm.coordinates, f(glat), gV.cell_node_map, gV.offset, gm.coord_map,
                           gm.coord_offset, alpha=1.0, beta=0.5)
    key = lambda Lq: (Lq[:, 0] * 1000 + Lq[:, 1]) * 1000 + Lq[:, 2]
    lookup = dict(zip(key(glat).tolist(), gy.tolist()))
    no = V.owned_node_count
    ref = np.array([lookup[kk] for kk in key(lat[:no]).tolist()])
    err = np.abs(y.data_ro[:no] - ref).max() / np.abs(gy).max()
    worst = max(worst, err)
    if exec_halo:
        # the same loop with HOST-resident Dats (pinned buffers in, host buffers out):
        # Parloop._call_host_partitioned -- pipeline for the core cells, NCCL for the ghost rows
        xh = op2.Dat(dn, xv.copy(), pinned=True)
        xh.halo_valid = False
        yh = op2.Dat(dn, pinned=True)
        yh.zero()
        gk = op2.GlobalKernel(k, [m0, m1], extruded=True)
        op2.Parloop(gk, cells, [yh(op2.INC, m0), X(op2.READ, m1), xh(op2.READ, m0)], location="host")()
        errh = np.abs(yh._data[:no] - ref).max() / np.abs(gy).max()
        worst = max(worst, errh)
    # global reduction (C3): x.x summed over owned dofs of all ranks
    xo = op2.Dat(op2.Set(no), f(lat)[:no])
    import ctypes as C
    from firedrake_b200 import _lib
    L = _lib.lib()
    loc = xo.inner(xo)
    d = op2.DeviceArray.from_host(np.array([loc]))
    _lib.check(L.fdb_allreduce(d.ptr, 1, 0))
    tot = d.to_host(np.zeros(1))[0]
    assert abs(tot - (f(glat) ** 2).sum()) < 1e-9 * abs(tot), (tot, (f(glat) ** 2).sum())
if world >= 3:
    # Advisor (round 1): an owned dof that is a ghost on TWO neighbours appears in two neighbour
    # lists of the SUM unpack (partition corners in the reference's DMPlex partitions; the slab
    # partition never produces one).  Ring: my dof 0 is ghost row 4 on my right neighbour and ghost
    # row 5 on my left neighbour; local->global must add both contributions.
    left, right = (rank - 1) % world, (rank + 1) % world
    ring = Halo([(left, np.array([0], dtype=np.int32), np.array([4], dtype=np.int32)),
                 (right, np.array([0], dtype=np.int32), np.array([5], dtype=np.int32))])
    rs = op2.Set((4, 4, 6))
    d = op2.Dat(op2.DataSet(rs, 1, halo=ring), np.array([1.0 + rank, 0, 0, 0, 10.0 * (rank + 1), 100.0 * (rank + 1)]))
    d.device_ptr
    ring.local_to_global_begin(d)
    ring.local_to_global_end(d)
    # my ghost row 4 is LEFT's dof 0, row 5 is RIGHT's dof 0: my dof 0 receives right's row 4 and left's row 5
    expect = 1.0 + rank + 10.0 * (right + 1) + 100.0 * (left + 1)
    assert abs(d.data_ro[0] - expect) < 1e-12, (rank, d.data_ro[0], expect)
    ring.global_to_local_begin(d)
    ring.global_to_local_end(d)
    got = d.data_ro_with_halos
    exp_l = 1.0 + left + 10.0 * (rank + 1) + 100.0 * ((left - 1) % world + 1)
    assert abs(got[4] - exp_l) < 1e-12, (rank, got[4], exp_l)
print(f"rank {rank}/{world}: worst rel err {worst:.2e}")
assert worst < 1e-12
if dist is not None:
    dist.barrier()
print("HALO_OK")
