"""World-size-2/3 CPU tests (gloo) of the N>1 host logic: slab partition,
core/owned/ghost ordering, halo send/recv lists.  The distributed action
(oracle compute + gloo exchanges, the protocol of pyop2/parloop.py:243-260)
must equal the serial action dof for dof."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, p, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _gloo_halo import exchange
    from firedrake_b200.fiat_lite import interval_element
    from firedrake_b200.partition import SlabPartition
    from firedrake_b200.utility_meshes import ExtrudedHexMesh
    from oracle import oracle
    nx, ny, nz = 5, 3, 4
    part = SlabPartition(nx, ny, nz, p, rank, world, warp=0.05)
    mesh, V = part.mesh, part.V
    # a global field defined through the lattice position of each dof
    lat = V.dof_lattice()
    f = lambda L: np.sin(0.37 * L[:, 0]) + 0.11 * L[:, 1] * L[:, 2] - 0.05 * L[:, 0] * L[:, 2]
    x = f(lat)
    # ghosts start stale: the global->local exchange must repair them
    x[V.owned_node_count:] = np.nan
    exchange(part.neighbours, x, reverse=False)
    assert np.allclose(x, f(lat))
    core, owned, _ = part.cell_sizes
    y = np.zeros(V.node_count)
    el = interval_element(p)
    for start, end in ((0, core), (core, owned)):        # core part, then owned part
        oracle.action_extruded(el, start, end, [0, mesh.layers], y, mesh.coordinates, x,
                               V.cell_node_map, V.offset, mesh.coord_map, mesh.coord_offset,
                               alpha=1.0, beta=0.5)
    exchange(part.neighbours, y, reverse=True)
    # serial reference on the whole mesh
    gm = ExtrudedHexMesh(nx, ny, nz, warp=0.05)
    gV = gm.function_space(p)
    glat = gV.dof_lattice()
    gy = np.zeros(gV.node_count)
    oracle.action_extruded(el, 0, gm.num_base_cells, [0, gm.layers], gy, gm.coordinates, f(glat),
                           gV.cell_node_map, gV.offset, gm.coord_map, gm.coord_offset,
                           alpha=1.0, beta=0.5)
    key = lambda L: (L[:, 0] * 1000 + L[:, 1]) * 1000 + L[:, 2]
    lookup = dict(zip(key(glat).tolist(), gy.tolist()))
    no = V.owned_node_count
    ref = np.array([lookup[k] for k in key(lat[:no]).tolist()])
    err = np.abs(y[:no] - ref).max() / np.abs(gy).max()
    # every global dof is owned exactly once
    import torch
    cnt = torch.tensor([no], dtype=torch.int64)
    dist.all_reduce(cnt)
    q.put((rank, err, int(cnt.item()), gV.node_count, mesh.coordinates[:, 0].min()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,p", [(2, 1), (2, 3), (3, 2)])
def test_distributed_action_matches_serial(world, p):
    from oracle import oracle
    oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, p, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, err, total_owned, nglobal, _ in res:
        assert err < 1e-12, (rank, err)
        assert total_owned == nglobal


def test_slab_sets_are_core_owned_ghost():
    from firedrake_b200.partition import SlabPartition, slab_bounds
    assert [slab_bounds(10, 3, r) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    part = SlabPartition(6, 4, 3, 2, rank=1, nranks=3)
    mesh, V = part.mesh, part.V
    core, owned, total = part.cell_sizes
    assert (core, owned, total) == (4, 8, 8)           # ix == 0 cells are not core
    ghost = set(range(V.owned_node_count, V.node_count))
    touched = set(V.full_cell_node_list()[: core * mesh.nz].ravel().tolist())
    assert not (touched & ghost)                       # core cells never read a ghost dof
    (r0, s0, q0), (r1, s1, q1) = part.neighbours
    assert (r0, r1) == (0, 2) and len(s0) == 0 and len(q1) == 0
    assert set(q0.tolist()) == ghost                   # recv list == the ghost tail, in order
    assert np.array_equal(q0, np.arange(V.owned_node_count, V.node_count))
    assert len(s1) == len(q0) and s1.max() < V.owned_node_count
