"""CPU checks of the multigrid transfer kernels (firedrake_b200/mg.py; reference
firedrake/mg/interface.py:37-280): the coarse-cell -> fine-node map, and the generated
prolong / restrict / inject parloops executed through the host build of the generated
wrapper (tests/_jit_host.py) -- polynomial exactness, restrict == prolong^T,
inject o prolong == id, vector-valued spaces -- plus NVRTC compilation for sm_100a."""
import numpy as np
import pytest

from firedrake_b200 import mg, op2
from firedrake_b200.codegen import CStringKernel, WrapperSpec
from firedrake_b200.utility_meshes import ExtrudedHexMesh

import _jit_host as jh


def _levels(p, warp=0.0, seed=None):
    mc = ExtrudedHexMesh(2, 3, 2, warp=warp, permute_seed=seed)
    mf = ExtrudedHexMesh(4, 6, 4, warp=warp, permute_seed=None if seed is None else seed + 1)
    return mc, mf, mc.function_space(p), mf.function_space(p)


def _objects(mc, mf, Vc, Vf, cdim=1):
    cells = op2.ExtrudedSet(op2.Set(mc.num_base_cells), mc.layers)
    cn, fn = op2.Set(Vc.node_count), op2.Set(Vf.node_count)
    mcn = op2.Map(cells, cn, Vc.arity, Vc.cell_node_map, offset=Vc.offset)
    vals, off = mg.coarse_to_fine_node_map(Vc, Vf)
    c2f = op2.Map(cells, fn, vals.shape[1], vals, offset=off)
    return cells, op2.DataSet(cn, cdim), op2.DataSet(fn, cdim), mcn, c2f


def _run(kernel, mc, args, dats, maps):
    spec = WrapperSpec(kernel, args, extruded=True)
    jh.run(spec, 0, mc.num_base_cells, [d._data for d in dats], [m.values_with_halo for m in maps],
           layers=[0, mc.layers])
    return spec


@pytest.mark.parametrize("p", [1, 2, 3])
def test_coarse_to_fine_map_addresses_the_right_nodes(p):
    """Every lattice point of every coarse cell (all layers, through the doubled offsets)
    must be the fine node sitting at the trilinear image of its lattice position."""
    mc, mf, Vc, Vf = _levels(p, seed=3)
    vals, off = mg.coarse_to_fine_node_map(Vc, Vf)
    Pf = Vf.dof_coordinates()
    pos = mg.fine_lattice_positions(p)
    m = 2 * p + 1
    hx, hy, hz = mc.Lx / mc.nx, mc.Ly / mc.ny, mc.Lz / mc.nz
    for c in range(mc.num_base_cells):
        for l in range(mc.nz):
            nodes = vals[c] + off * l
            x = (mc.cell_ix[c] + pos)[:, None, None] * hx + 0 * pos[None, :, None] + 0 * pos[None, None, :]
            y = (mc.cell_iy[c] + pos)[None, :, None] * hy + 0 * x
            z = (l + pos)[None, None, :] * hz + 0 * x
            ref = np.stack([x, y, z], axis=-1).reshape(m ** 3, 3)
            assert np.abs(Pf[nodes] - ref).max() < 1e-13
    # every fine node is reached
    allnodes = (vals[:, None, :] + off[None, None, :] * np.arange(mc.nz)[None, :, None]).ravel()
    assert np.array_equal(np.unique(allnodes), np.arange(Vf.node_count))


@pytest.mark.parametrize("p,cdim", [(1, 1), (2, 1), (3, 1), (2, 3)])
def test_prolong_restrict_inject(p, cdim):
    mc, mf, Vc, Vf = _levels(p, seed=5)
    cells, dc, df, mcn, c2f = _objects(mc, mf, Vc, Vf, cdim)
    Pc, Pf = Vc.dof_coordinates(), Vf.dof_coordinates()
    # a polynomial of degree p in each direction is reproduced exactly by prolongation
    poly = lambda P: np.stack([(1 + P[:, 0]) ** p * (2 - P[:, 1]) ** p * (0.5 + P[:, 2]) ** p * (c + 1)
                               for c in range(cdim)], axis=1)
    uc = op2.Dat(dc, poly(Pc))
    uf = op2.Dat(df)
    _run(mg.prolong_kernel(p, cdim), mc, [uf(op2.WRITE, c2f), uc(op2.READ, mcn)], [uf, uc], [c2f, mcn])
    assert np.abs(uf._data.reshape(-1, cdim) - poly(Pf)).max() < 1e-12 * np.abs(poly(Pf)).max()
    # injection recovers the coarse function
    back = op2.Dat(dc)
    _run(mg.inject_kernel(p, cdim), mc, [back(op2.WRITE, mcn), uf(op2.READ, c2f)], [back, uf], [mcn, c2f])
    assert np.abs(back._data - uc._data).max() < 1e-12 * np.abs(uc._data).max()
    # multiplicity weights, then restrict == prolong^T
    w = op2.Dat(op2.DataSet(df.set, 1))
    m3 = (2 * p + 1) ** 3
    count = CStringKernel(f"static void count(double *w) {{ for (int i = 0; i < {m3}; ++i) w[i] += 1.0; }}", "count")
    _run(count, mc, [w(op2.INC, c2f)], [w], [c2f])
    assert w._data.min() >= 1 and w._data.max() <= 8
    w._data[:] = 1.0 / w._data
    rng = np.random.default_rng(11)
    vc = op2.Dat(dc, rng.standard_normal((Vc.node_count, cdim)))
    rf = op2.Dat(df, rng.standard_normal((Vf.node_count, cdim)))
    pv = op2.Dat(df)
    _run(mg.prolong_kernel(p, cdim), mc, [pv(op2.WRITE, c2f), vc(op2.READ, mcn)], [pv, vc], [c2f, mcn])
    rc = op2.Dat(dc)
    _run(mg.restrict_kernel(p, cdim), mc, [rc(op2.INC, mcn), rf(op2.READ, c2f), w(op2.READ, c2f)],
         [rc, rf, w], [mcn, c2f])
    lhs, rhs = float((pv._data * rf._data).sum()), float((vc._data * rc._data).sum())
    assert abs(lhs - rhs) < 1e-11 * max(abs(lhs), 1.0)


def test_transfer_kernels_compile_for_sm100a():
    mc, mf, Vc, Vf = _levels(3)
    for cdim in (1, 3):
        cells, dc, df, mcn, c2f = _objects(mc, mf, Vc, Vf, cdim)
        uc, uf, w = op2.Dat(dc), op2.Dat(df), op2.Dat(op2.DataSet(df.set, 1))
        for k, args in ((mg.prolong_kernel(3, cdim), [uf(op2.WRITE, c2f), uc(op2.READ, mcn)]),
                        (mg.restrict_kernel(3, cdim), [uc(op2.INC, mcn), uf(op2.READ, c2f), w(op2.READ, c2f)]),
                        (mg.inject_kernel(3, cdim), [uc(op2.WRITE, mcn), uf(op2.READ, c2f)])):
            assert WrapperSpec(k, args, extruded=True).compile()[:4] == b"\x7fELF"


def test_transfer_matrices():
    for p in (1, 2, 3, 4):
        P, J = mg.prolongation_matrix(p), mg.injection_matrix(p)
        assert np.allclose(P.sum(axis=1), 1.0)                 # partition of unity
        assert np.allclose(J @ P, np.eye(p + 1), atol=1e-13)   # inject o prolong = id in 1-D
