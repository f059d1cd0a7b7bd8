"""DG advection (config 3): cell + exterior-facet +
interior-facet kernels on the device against the oracle, plus the invariants
of reference tests/firedrake/regression/test_dg_advection.py:62-75."""
import numpy as np
import pytest

from firedrake_b200.assemble import DGAdvection
from firedrake_b200.utility_meshes import QuadMesh

pytestmark = pytest.mark.gpu


def make(nx, ny, distort=0.02, seed=0):
    m = QuadMesh(nx, ny)
    X = m.coordinates
    rng = np.random.default_rng(seed)
    inner = (X[:, 0] > 1e-9) & (X[:, 0] < 1 - 1e-9) & (X[:, 1] > 1e-9) & (X[:, 1] < 1 - 1e-9)
    X[inner] += distort / max(nx, ny) * 8 * rng.standard_normal((inner.sum(), 2))
    u = np.stack([0.5 - X[:, 1], X[:, 0] - 0.5], axis=1)          # solid-body rotation
    return m, u, rng


@pytest.mark.parametrize("nq", [2, 3])
def test_rhs_matches_oracle(engine, oracle, nq):
    m, u, rng = make(9, 7)
    q = 1.0 + rng.random(m.num_cells * 4)
    prob = DGAdvection(m, dt=0.01, q_in=1.0, nq=nq)
    out = prob.assemble(prob.function(q.copy()), prob.velocity(u.copy()))
    ref = oracle.dg_rhs(m, q, u, dt=0.01, q_in=1.0, nq=nq)
    assert np.abs(out.data_ro - ref).max() < 1e-12 * np.abs(ref).max()


def test_constant_state_is_stationary(engine):
    """q == q_in == const and div u = 0  =>  L1 == 0 (divergence theorem): the
    cell, exterior and interior facet kernels cancel exactly."""
    m, u, _ = make(40, 40)
    prob = DGAdvection(m, dt=2 * np.pi / 600, q_in=1.0)
    out = prob.assemble(prob.function(np.ones(m.num_cells * 4)), prob.velocity(u.copy()))
    assert np.abs(out.data_ro).max() < 1e-14


def test_interior_fluxes_conserve_mass(engine, oracle):
    """Sum over all test functions: interior-facet terms cancel pairwise, so
    sum(L1) only sees the boundary flux (mass conservation,
    test_dg_advection.py:62-75)."""
    m, u, rng = make(16, 12)
    q = 1.0 + rng.random(m.num_cells * 4)
    prob = DGAdvection(m, dt=0.05, q_in=1.0)
    out = prob.assemble(prob.function(q.copy()), prob.velocity(u.copy()))
    # DQ1 nodal basis sums to one: sum_i L1_i = dt*( int q div(u) - boundary flux ) = -dt * boundary flux
    only_boundary = oracle.dg_rhs(m, q, u, dt=0.05, q_in=1.0)
    assert abs(out.data_ro.sum() - only_boundary.sum()) < 1e-13
    # and with an interior-only perturbation of q the sum does not change
    q2 = q.copy()
    interior_cells = [c for c in range(m.num_cells) if c not in set(m.ext_facet_cells.tolist())]
    q2[np.array(interior_cells)[:, None] * 4 + np.arange(4)] += 0.3
    out2 = prob.assemble(prob.function(q2), prob.velocity(u.copy()))
    assert abs(out2.data_ro.sum() - out.data_ro.sum()) < 1e-12


@pytest.mark.parametrize("nq", [2, 3])
def test_fused_owner_computes_kernel(engine, oracle, nq):
    """One pass over the cells (cell + its four facets, no atomics) equals the
    cell / exterior / interior parloops and the oracle, and is bit-reproducible."""
    m, u, rng = make(11, 9)
    q = 1.0 + rng.random(m.num_cells * 4)
    fused = DGAdvection(m, dt=0.01, q_in=1.0, nq=nq, fused=True)
    outs = [fused.assemble(fused.function(q.copy()), fused.velocity(u.copy())).data_ro.copy() for _ in range(2)]
    assert np.array_equal(outs[0], outs[1])
    ref = oracle.dg_rhs(m, q, u, dt=0.01, q_in=1.0, nq=nq)
    assert np.abs(outs[0] - ref).max() < 1e-12 * np.abs(ref).max()
    loops = DGAdvection(m, dt=0.01, q_in=1.0, nq=nq)
    out3 = loops.assemble(loops.function(q.copy()), loops.velocity(u.copy()))
    assert np.abs(outs[0] - out3.data_ro).max() < 1e-12 * np.abs(ref).max()
