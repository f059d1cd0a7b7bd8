"""World-size-2/3 CPU tests (gloo + the mock engine) of the DISTRIBUTED host logic:
Parloop's halo protocol for the fast path (passes on 2/4 real GPUs: calibration) and,
new, the same protocol for generic parloops -- ghost refresh of read Dats, local->global
sum of INC Dats, all-reduce of Globals (pyop2/parloop.py:243-260, 354-455) -- through
assemble_functional and a C-string Poisson kernel on a slab-partitioned mesh."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from test_partition_gloo import ROOT, _free_port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _mock_engine as me
    import test_codegen as tc
    from firedrake_b200 import op2
    from firedrake_b200.assemble import (FunctionSpace, OneFormAssembler, assemble_functional, helmholtz,
                                         interpolate)
    from firedrake_b200.partition import SlabPartition
    from firedrake_b200.utility_meshes import ExtrudedHexMesh
    from oracle import oracle
    out = {}
    with me.install(oracle) as eng:
        eng.dist = dist
        nx, ny, nz = 5, 3, 4
        # ---- serial references on the whole mesh (every rank computes them redundantly, no comm)
        eng.dist = None
        gm = ExtrudedHexMesh(nx, ny, nz, warp=0.05)
        G = FunctionSpace(gm, 1)
        expr = "sin(2.0 * x[0]) + x[1] * x[2]"
        gu = interpolate(G, expr)
        gy = OneFormAssembler(helmholtz(G), gu).assemble()
        gyg = G.dat()
        op2.par_loop(op2.Kernel(tc.Q1_POISSON, "q1_poisson"), G.cell_set, gyg(op2.INC, G.cell_node_map),
                     G.coordinates(op2.READ, G.coord_map), gu(op2.READ, G.cell_node_map))
        ref = dict(dx=assemble_functional(G, gu, "dx"), ds=assemble_functional(G, gu, "ds"))
        glat = G.V.dof_lattice()
        key = lambda L: (L[:, 0] * 1000 + L[:, 1]) * 1000 + L[:, 2]
        look_fast = dict(zip(key(glat).tolist(), gy.data_ro.tolist()))
        look_gen = dict(zip(key(glat).tolist(), gyg.data_ro.tolist()))
        # ---- the same on this rank's slab, with halos
        eng.dist = dist
        part = SlabPartition(nx, ny, nz, 1, rank, world, warp=0.05)
        V = FunctionSpace(part.mesh, 1, partition=part)
        u = interpolate(V, expr)
        lat = V.V.dof_lattice()
        no = V.V.owned_node_count
        y = OneFormAssembler(helmholtz(V), u).assemble()                       # fast path + halos
        out["fast"] = float(np.abs(y.data_ro[:no] - np.array([look_fast[k] for k in key(lat[:no]).tolist()])).max())
        u2 = interpolate(V, expr)                                               # ghosts stale again
        yg = V.dat()
        yg.device_ptr
        op2.par_loop(op2.Kernel(tc.Q1_POISSON, "q1_poisson"), V.cell_set, yg(op2.INC, V.cell_node_map),
                     V.coordinates(op2.READ, V.coord_map), u2(op2.READ, V.cell_node_map))
        out["generic"] = float(np.abs(yg.data_ro[:no] - np.array([look_gen[k] for k in key(lat[:no]).tolist()])).max())
        # explicit distributed matrix, kept unassembled (mat_type="is"): SpMV == serial assembled SpMV,
        # with Dirichlet conditions on the top face (constrained nodes on the slab interfaces included)
        from firedrake_b200.assemble import DirichletBC, assemble
        eng.dist = None
        gA = assemble(helmholtz(G), bcs=[DirichletBC(G, 0.0, "top")])
        gz = G.dat()
        gA.mult(gu, gz)
        look_mat = dict(zip(key(glat).tolist(), gz.data_ro.tolist()))
        eng.dist = dist
        A = assemble(helmholtz(V), bcs=[DirichletBC(V, 0.0, "top")], mat_type="is")
        u3 = interpolate(V, expr)
        z = V.dat()
        z.device_ptr
        A.mult(u3, z)
        out["ismat"] = float(np.abs(z.data_ro[:no] - np.array([look_mat[k] for k in key(lat[:no]).tolist()])).max())
        # exec-halo (owner-computes) partition: redundant execution of the right neighbour's first
        # cell column replaces the local->global reduce (SURVEY.md section 8e option (ii))
        for pdeg in (1, 3):
            eng.dist = None
            Gp = FunctionSpace(gm, pdeg)
            gup = interpolate(Gp, expr)
            gyp = OneFormAssembler(helmholtz(Gp), gup).assemble()
            lookp = dict(zip(key(Gp.V.dof_lattice()).tolist(), gyp.data_ro.tolist()))
            eng.dist = dist
            parte = SlabPartition(nx, ny, nz, pdeg, rank, world, warp=0.05, exec_halo=True)
            Ve = FunctionSpace(parte.mesh, pdeg, partition=parte)
            ue = interpolate(Ve, expr)
            ue.halo_valid = False
            ye = OneFormAssembler(helmholtz(Ve), ue).assemble()
            noe = Ve.V.owned_node_count
            late = Ve.V.dof_lattice()
            out["exec%d" % pdeg] = float(np.abs(ye.data_ro[:noe] - np.array([lookp[k] for k in key(late[:noe]).tolist()])).max())
        # advisor (round 1): a SECOND INC loop into the same Dat without zero() in between must add
        # this loop's contributions once (ghost rows restart from the INC identity)
        u4 = interpolate(V, expr)
        op2.par_loop(op2.Kernel(tc.Q1_POISSON, "q1_poisson"), V.cell_set, yg(op2.INC, V.cell_node_map),
                     V.coordinates(op2.READ, V.coord_map), u4(op2.READ, V.cell_node_map))
        out["twice"] = float(np.abs(yg.data_ro[:no] - 2 * np.array([look_gen[k] for k in key(lat[:no]).tolist()])).max())
        # advisor (round 1): CG on the unassembled (mat_type="is") operator, pc none: the raw vector
        # updates leave stale ghost rows that every mult must refresh
        from firedrake_b200.assemble import cg
        import torch

        def allreduce(v):
            t = torch.tensor([v], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())
        eng.dist = None
        gb = G.dat(np.cos(glat[:, 0] * 0.7) + 0.1 * glat[:, 1] - 0.05 * glat[:, 2] ** 2)
        for bc in [DirichletBC(G, 0.0, "top")]:
            bc.zero(gb)
        gx = G.dat()
        gx.device_ptr
        n_ser, _ = cg(gA, gb, gx, rtol=1e-11, maxit=500)
        look_sol = dict(zip(key(glat).tolist(), gx.data_ro.tolist()))
        eng.dist = dist
        bl = V.dat(np.cos(lat[:, 0] * 0.7) + 0.1 * lat[:, 1] - 0.05 * lat[:, 2] ** 2)
        for bc in [DirichletBC(V, 0.0, "top")]:
            bc.zero(bl)
        xl = V.dat()
        xl.device_ptr
        n_par, _ = cg(A, bl, xl, rtol=1e-11, maxit=500, allreduce=allreduce)
        sol = np.array([look_sol[k] for k in key(lat[:no]).tolist()])
        out["cg_is"] = float(np.abs(xl.data_ro[:no] - sol).max() / np.abs(sol).max())
        out["cg_its"] = (n_ser, n_par)
        out["dx"] = abs(assemble_functional(V, u2, "dx") - ref["dx"])
        out["ds"] = abs(assemble_functional(V, u2, "ds") - ref["ds"])
        out["scale"] = float(np.abs(gy.data_ro).max())
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_generic_parloops(world):
    from oracle import oracle
    oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, out in res:
        assert out["fast"] < 1e-12 * out["scale"], (rank, out)
        assert out["generic"] < 1e-12 * out["scale"], (rank, out)
        assert out["ismat"] < 1e-12 * out["scale"], (rank, out)
        assert out["dx"] < 1e-12 and out["ds"] < 1e-12, (rank, out)
        assert out["twice"] < 1e-12 * out["scale"], (rank, out)
        assert out["exec1"] < 1e-12 * out["scale"] and out["exec3"] < 1e-11 * out["scale"], (rank, out)
        assert out["cg_is"] < 1e-8 and abs(out["cg_its"][0] - out["cg_its"][1]) <= 2, (rank, out)


def _mg_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _mock_engine as me
    from firedrake_b200 import mg
    from firedrake_b200.assemble import helmholtz
    from oracle import oracle

    def allreduce(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    nx = 2 * world
    rhs = lambda P: np.sin(3.0 * P[:, 0]) * (1 + P[:, 1]) + P[:, 2] ** 2
    with me.install(oracle) as eng:
        # serial reference (no communication)
        eng.dist = None
        h = mg.MeshHierarchy(nx, 4, 4, 2, warp=0.03)
        vc = mg.VCycle(h, 1, helmholtz, bc_domains=("bottom",), coarse_rtol=1e-10)
        V, A = vc.spaces[2], vc.ops[2]
        b = V.dat(rhs(V.V.dof_coordinates()))
        for bc in vc.bcs[2]:
            bc.zero(b)
        x = V.dat()
        x.device_ptr
        n_serial, _ = mg.pcg(A, b, x, lambda r, z: vc.apply(2, r, z), rtol=1e-9)
        key = lambda L: (L[:, 0] * 1000 + L[:, 1]) * 1000 + L[:, 2]
        look = dict(zip(key(V.V.dof_lattice()).tolist(), x.data_ro.tolist()))
        # the same hierarchy cut into slabs
        eng.dist = dist
        hp = mg.MeshHierarchy(nx, 4, 4, 2, rank=rank, nranks=world, warp=0.03)
        vp = mg.VCycle(hp, 1, helmholtz, bc_domains=("bottom",), coarse_rtol=1e-10, allreduce=allreduce)
        W, B = vp.spaces[2], vp.ops[2]
        bp = W.dat(rhs(W.V.dof_coordinates()))
        for bc in vp.bcs[2]:
            bc.zero(bp)
        xp = W.dat()
        xp.device_ptr
        n_par, hist = mg.pcg(B, bp, xp, lambda r, z: vp.apply(2, r, z), rtol=1e-9, allreduce=allreduce)
        no = W.V.owned_node_count
        ref = np.array([look[k] for k in key(W.V.dof_lattice()[:no]).tolist()])
        err = float(np.abs(xp.data_ro[:no] - ref).max() / np.abs(ref).max())
    q.put((rank, n_serial, n_par, err))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_multigrid(world):
    """The V-cycle on a slab-partitioned hierarchy (transfers through ghost rows, restricted
    residuals summed into their owners, all-reduced inner products) converges like the serial one
    and to the same solution."""
    from oracle import oracle
    oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mg_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, n_serial, n_par, err in res:
        assert abs(n_par - n_serial) <= 1, res
        assert err < 1e-7, res
