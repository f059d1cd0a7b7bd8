"""BASELINE.json's FULL sizes: entry-by-entry against the oracle (its banded multi-threaded
wrapper finishes a 256^3 pass in seconds on the GPU box's host cores), and through
size-independent properties:

* config 2  Poisson CG3, 256^3 hexes: constants are in the null space of the
  stiffness operator; 1^T M 1 = |Omega|; linearity; symmetry x.(K y) = y.(K x)
* config 5  Poisson CG5, 128^3
* config 3  DG advection DQ1, 2048^2 quads: q == q_in with div u = 0 is stationary
"""
import numpy as np
import pytest

from firedrake_b200 import op2
from firedrake_b200.assemble import DGAdvection, FunctionSpace, Form, OneFormAssembler
from firedrake_b200.utility_meshes import ExtrudedHexMesh, QuadMesh

pytestmark = pytest.mark.gpu


def _dot(a, b):
    return a.inner(b)


@pytest.mark.parametrize("n,p,cdim,beta", [(256, 3, 1, 0.0), (128, 5, 1, 0.0), (64, 4, 3, 1.0)])
def test_full_size_matches_oracle(engine, oracle, n, p, cdim, beta):
    """Configs 2, 5 and 4 (action) at BASELINE size: every DoF of the device result against the
    oracle on the same seeded input; 1e-12 relative in the max norm.
    Reference anchor: tests/firedrake/regression/test_matrix_free.py:98-127."""
    from firedrake_b200.fiat_lite import interval_element
    mesh = ExtrudedHexMesh(n, n, n, warp=0.05)
    V = mesh.function_space(p)
    cells = op2.ExtrudedSet(op2.Set(mesh.num_base_cells), mesh.layers)
    nodes = op2.Set(V.node_count)
    vnodes = op2.Set(mesh.coord_space.node_count)
    m0 = op2.Map(cells, nodes, V.arity, V.cell_node_map, offset=V.offset)
    m1 = op2.Map(cells, vnodes, 8, mesh.coord_map, offset=mesh.coord_offset)
    rng = np.random.default_rng(1234)
    xh = np.empty(V.node_count * cdim)
    for i in range(0, xh.size, 1 << 24):
        xh[i:i + (1 << 24)] = rng.standard_normal(min(1 << 24, xh.size - i))
    shape = (V.node_count,) if cdim == 1 else (V.node_count, cdim)
    x = op2.Dat(op2.DataSet(nodes, cdim), xh.reshape(shape))
    y = op2.Dat(op2.DataSet(nodes, cdim))
    X = op2.Dat(op2.DataSet(vnodes, 3), mesh.coordinates)
    k = op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=beta, cdim=cdim)
    op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    yo = np.zeros(V.node_count * cdim)
    oracle.action_extruded_parallel(interval_element(p), mesh, yo, np.ascontiguousarray(mesh.coordinates), xh,
                                    V.cell_node_map, V.offset, mesh.coord_map, mesh.coord_offset, cdim=cdim,
                                    alpha=1.0, beta=beta, native=True)
    err = np.abs(y.data_ro.reshape(-1) - yo).max() / np.abs(yo).max()
    assert err < 1e-12, err


@pytest.mark.parametrize("n,p", [(256, 3), (128, 5)])
def test_operator_properties_at_full_size(engine, n, p):
    mesh = ExtrudedHexMesh(n, n, n, warp=0.05)
    V = FunctionSpace(mesh, p)
    N = V.node_count
    assert N == (n * p + 1) ** 3
    rng = np.random.default_rng(0)
    x, y = V.dat(), V.dat()
    xa = x.data_with_halos
    ya = y.data_with_halos
    for i in range(0, N, 1 << 24):
        xa[i:i + (1 << 24)] = rng.standard_normal(min(1 << 24, N - i))
        ya[i:i + (1 << 24)] = rng.standard_normal(min(1 << 24, N - i))
    ones = V.dat()
    ones.data_with_halos[:] = 1.0
    K = lambda u: OneFormAssembler(Form(V, 1.0, 0.0), u)
    M = lambda u: OneFormAssembler(Form(V, 0.0, 1.0), u)
    out = V.dat()
    # constants in the null space: |K 1| ~ rounding of entries of size ~ p^2/h
    K(ones).assemble(out)
    assert np.sqrt(_dot(out, out) / N) < 1e-11
    # 1^T M 1 = volume of the (boundary-preserving) warped unit cube
    M(ones).assemble(out)
    assert abs(_dot(out, ones) - 1.0) < 1e-10
    # symmetry: x.(K y) == y.(K x)
    Ky, Kx = V.dat(), V.dat()
    K(y).assemble(Ky)
    K(x).assemble(Kx)
    a, b = _dot(x, Ky), _dot(y, Kx)
    assert abs(a - b) < 1e-9 * max(abs(a), abs(b), np.sqrt(_dot(Kx, Kx) * _dot(y, y)))
    # linearity: K(2x + y) == 2 Kx + Ky
    z = V.dat()
    z.axpy(2.0, x)
    z.axpy(1.0, y)
    Kz = V.dat()
    K(z).assemble(Kz)
    Kz.axpy(-2.0, Kx)
    Kz.axpy(-1.0, Ky)
    assert np.sqrt(_dot(Kz, Kz)) < 1e-11 * np.sqrt(_dot(Kx, Kx))


def test_dg_constant_state_at_full_size(engine):
    n = 2048
    m = QuadMesh(n, n)
    X = m.coordinates
    u = np.stack([0.5 - X[:, 1], X[:, 0] - 0.5], axis=1)
    for fused in (False, True):
        prob = DGAdvection(m, dt=2 * np.pi / 600 * 40 / n, q_in=1.0, fused=fused)
        out = prob.assemble(prob.function(np.ones(m.num_cells * 4)), prob.velocity(u.copy()))
        assert np.abs(out.data_ro).max() < 1e-16
