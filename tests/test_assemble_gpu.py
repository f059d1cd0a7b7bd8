"""The assemble()/DirichletBC/ImplicitMatrixContext surface on the device
 and the device CG of config 5."""
import numpy as np
import pytest

from firedrake_b200 import op2
from firedrake_b200.assemble import (DirichletBC, FunctionSpace, ImplicitMatrixContext, assemble, cg,
                                     helmholtz, poisson)
from firedrake_b200.utility_meshes import ExtrudedHexMesh

pytestmark = pytest.mark.gpu


def test_matfree_equals_assembled_with_bcs(engine):
    """reference tests/firedrake/regression/test_matrix_free.py:98-127."""
    mesh = ExtrudedHexMesh(4, 4, 5, warp=0.05, permute_seed=0)
    V = FunctionSpace(mesh, 2)
    bcs = [DirichletBC(V, 0.0, ["bottom", "top"])]
    a = helmholtz(V)
    A = assemble(a, bcs=bcs)
    Amf = assemble(a, bcs=bcs, mat_type="matfree")
    assert isinstance(Amf, ImplicitMatrixContext)
    x = V.dat(np.random.default_rng(0).standard_normal(V.node_count))
    y1, y2 = V.dat(), V.dat()
    A.mult(x, y1)
    Amf.mult(x, y2)
    assert np.abs(y1.data_ro - y2.data_ro).max() < 1e-12 * np.abs(y1.data_ro).max()


def test_poisson_solve_strong_bcs_extrusion(engine):
    """reference tests/firedrake/extrusion/test_poisson_strong_bcs_extrusion.py:
    -div grad u = 0, u = 0 on the bottom, u = 42 on the top  =>  u = 42 z.
    Solved matrix-free with device CG (lifting of the boundary values)."""
    for p in (1, 3):
        mesh = ExtrudedHexMesh(3, 3, 6, warp=0.0)
        V = FunctionSpace(mesh, p)
        z = V.V.dof_coordinates()[:, 2]
        bcs = [DirichletBC(V, 0.0, ["bottom", "top"])]
        a = poisson(V)
        # lift: u = u0 + g with g = 42 on the top nodes;  A u0 = -K g on free rows
        g = np.zeros(V.node_count)
        g[V.boundary_nodes("top")] = 42.0
        Kg = assemble(a, u=V.dat(g.copy()))
        b = V.dat(-Kg.data_ro)
        bcs[0].zero(b)
        A = assemble(a, bcs=bcs, mat_type="matfree")
        u0 = V.dat()
        its, hist = cg(A, b, u0, rtol=1e-12, maxit=500)
        u = u0.data_ro + g
        assert np.abs(u - 42.0 * z).max() < 1e-6, (p, its, hist[-1])


def test_cg_matches_scipy(engine):
    mesh = ExtrudedHexMesh(4, 3, 4, warp=0.05)
    V = FunctionSpace(mesh, 2)
    a = helmholtz(V)
    A = assemble(a)
    Amf = assemble(a, mat_type="matfree")
    rng = np.random.default_rng(5)
    bv = rng.standard_normal(V.node_count)
    x = V.dat()
    its, hist = cg(Amf, V.dat(bv.copy()), x, rtol=1e-11, maxit=2000)
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    ro, co, va = A.csr()
    xs = spla.spsolve(sp.csr_matrix((va, co, ro), shape=(V.node_count,) * 2).tocsc(), bv)
    assert np.abs(x.data_ro - xs).max() < 1e-8 * np.abs(xs).max()
    assert hist[-1] < 1e-10 * hist[0]


def test_get_diagonal(engine):
    """reference tests/firedrake/regression/test_assemble.py:188-206: the
    assembled diagonal equals the diagonal of the assembled matrix."""
    mesh = ExtrudedHexMesh(4, 3, 5, warp=0.05)
    V = FunctionSpace(mesh, 3)
    bcs = [DirichletBC(V, 0.0, "bottom")]
    a = helmholtz(V)
    A = assemble(a, bcs=bcs)
    Amf = assemble(a, bcs=bcs, mat_type="matfree")
    d = Amf.getDiagonal(V.dat())
    ro, co, va = A.csr()
    diag = np.array([va[ro[r]:ro[r + 1]][co[ro[r]:ro[r + 1]] == r][0] for r in range(V.node_count)])
    assert np.abs(d.data_ro - diag).max() < 1e-12 * np.abs(diag).max()


@pytest.mark.parametrize("p", [1, 3, 4])
def test_interpolate_coordinates(engine, p):
    """interpolate(SpatialCoordinate) into CG_p^3 ==
    the trilinear image of the reference node positions (NumPy restatement
    ``ExtrudedFunctionSpace.dof_coordinates``), and a scalar Q1 field too."""
    from firedrake_b200.assemble import interpolate_q1
    mesh = ExtrudedHexMesh(4, 3, 5, warp=0.05, permute_seed=2)
    V = FunctionSpace(mesh, p)
    Xp = interpolate_q1(V, V.coordinates)
    ref = V.V.dof_coordinates()
    assert np.abs(Xp.data_ro - ref).max() < 1e-14
    w = op2.Dat(op2.DataSet(V.vertex_set, 1), 2.0 * mesh.coordinates[:, 0] - mesh.coordinates[:, 2])
    wp = interpolate_q1(V, w)
    assert np.abs(wp.data_ro - (2.0 * ref[:, 0] - ref[:, 2])).max() < 1e-14


def test_submatrix_and_duplicate_of_matrix_free_context(engine):
    """createSubMatrix / duplicate of the matrix-free context (firedrake/matrix_free/operators.py:380-470):
    the virtual sub-matrix applied to a compact vector equals the sub-block of the assembled matrix."""
    import scipy.sparse as sp
    from firedrake_b200 import op2
    from firedrake_b200.assemble import DirichletBC, FunctionSpace, assemble, helmholtz
    from firedrake_b200.utility_meshes import ExtrudedHexMesh
    mesh = ExtrudedHexMesh(3, 4, 3, warp=0.05)
    V = FunctionSpace(mesh, 2)
    bcs = [DirichletBC(V, 0.0, "top")]
    ctx = assemble(helmholtz(V), bcs=bcs, mat_type="matfree")
    A = assemble(helmholtz(V), bcs=bcs)
    ro, co, va = A.csr()
    Ad = sp.csr_matrix((va, co, ro), shape=(V.node_count,) * 2).toarray()
    rng = np.random.default_rng(11)
    rows = np.sort(rng.choice(V.node_count, 40, replace=False)).astype(np.int32)
    cols = np.sort(rng.choice(V.node_count, 55, replace=False)).astype(np.int32)
    sub = ctx.createSubMatrix(rows, cols)
    xs = op2.Dat(sub.col_set, rng.standard_normal(len(cols)))
    ys = op2.Dat(sub.row_set)
    sub.mult(xs, ys)
    ref = Ad[np.ix_(rows, cols)] @ xs.data_ro
    assert np.abs(ys.data_ro - ref).max() < 1e-12 * max(np.abs(ref).max(), 1.0)
    whole = np.arange(V.node_count)
    dup = ctx.createSubMatrix(whole, whole)
    assert type(dup) is type(ctx) and dup is not ctx
    x = V.dat(rng.standard_normal(V.node_count))
    y1, y2 = V.dat(), V.dat()
    ctx.mult(x, y1)
    ctx.duplicate().mult(x, y2)
    assert np.array_equal(y1.data_ro, y2.data_ro) or np.abs(y1.data_ro - y2.data_ro).max() < 1e-13 * np.abs(y1.data_ro).max()
