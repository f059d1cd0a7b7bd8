"""Pins the CPU oracle (oracle/) against every value-level anchor the reference
holds for this path -- golden arrays at the PyOP2
boundary, and the analytic invariants of the Firedrake regression tests --
plus an independent dense NumPy evaluation of the element tensors.  CPU only.
"""
import json
import os

import numpy as np
import pytest

from firedrake_b200.fiat_lite import gauss_legendre, gll_points, interval_element
from firedrake_b200.utility_meshes import ExtrudedHexMesh, UnitSquareTriMesh

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "pyop2_test_matrices.json")))


def dense(rowptr, colidx, vals, n):
    A = np.zeros((n, n))
    for r in range(n):
        A[r, colidx[rowptr[r]:rowptr[r + 1]]] = vals[rowptr[r]:rowptr[r + 1]]
    return A


# ---- 1. PyOP2 golden arrays (reference tests/pyop2/test_matrices.py:637-658)
def test_pyop2_mass_matrix_and_rhs_golden(oracle, golden):
    cmap = np.array(golden["elem_node_map"], dtype=np.int32)
    coords = np.array(golden["coords"])
    tab = (6, np.ascontiguousarray(golden["CG1"], dtype=float),
           np.ascontiguousarray(golden["d_CG1"], dtype=float), np.array(golden["w"]))
    rowptr, colidx = oracle.build_sparsity(4, cmap)
    vals = np.zeros(len(colidx))
    oracle.tri_matrix("mass", 0, 2, rowptr, colidx, vals, coords, cmap, tab, use_abs=0)
    M = dense(rowptr, colidx, vals, 4)
    np.testing.assert_allclose(M, np.array(golden["expected_matrix"]), atol=golden["expected_matrix_eps"])
    b = np.zeros(4)
    oracle.tri_rhs(0, 2, b, coords, np.array(golden["f"]), cmap, tab, use_abs=0)
    np.testing.assert_allclose(b, np.array(golden["expected_rhs"]), rtol=golden["expected_rhs_eps"])
    # test_solve (:660-666): M x = b recovers f
    np.testing.assert_allclose(np.linalg.solve(M, b), golden["f"], rtol=1e-8)


# ---- 2. wrapper semantics: indirect INC, extruded layer loop
def test_indirect_inc_counts(oracle):
    """reference tests/pyop2/test_indirect_loop.py:150-156: every element
    increments one shared entry -> nelems."""
    nelems = 4096
    cmap = np.zeros((nelems, 1), dtype=np.int32)
    u = np.zeros(1)
    oracle.indirect_inc(0, nelems, cmap, np.ones(nelems), u)
    assert u[0] == nelems


def test_extruded_layer_loop_and_offsets(oracle):
    """reference tests/pyop2/test_extrusion.py:363-370 (a direct INC over an
    ExtrudedSet with 10 node layers touches each entry layers-1 = 9 times) and
    pyop2/codegen/builder.py:100-123 (index = map + offset*layer)."""
    mesh = ExtrudedHexMesh(3, 2, 9)
    V = mesh.function_space(1)
    el = interval_element(1)
    # mass action on ones: each cell adds its volume share; column sums tell
    # how many layers were visited
    y = np.zeros(V.node_count)
    oracle.action_extruded(el, 0, mesh.num_base_cells, [0, mesh.layers], y, mesh.coordinates,
                           np.ones(V.node_count), V.cell_node_map, V.offset, mesh.coord_map,
                           mesh.coord_offset, alpha=0.0, beta=1.0)
    assert abs(y.sum() - 1.0) < 1e-13
    full = V.full_cell_node_list()
    assert full.shape == (3 * 2 * 9, 8)
    # offsets: P1 columns advance by one node per layer
    assert np.all(V.offset == 1)
    np.testing.assert_array_equal(full[1] - full[0], V.offset)


def test_extruded_dof_layout_q3():
    """Column layout of SURVEY.md section 9.3: Q3 x P3 vertex / edge / face
    columns have layer strides 3 / 6 / 12 and hold n0*L + n1*(L-1) dofs."""
    nz = 5
    mesh = ExtrudedHexMesh(2, 2, nz)
    V = mesh.function_space(3)
    assert sorted(set(V.offset.tolist())) == [3, 6, 12]
    assert V.node_count == (2 * 3 + 1) ** 2 * (nz * 3 + 1)
    full = V.full_cell_node_list()
    # every node is reached, the top of layer l is the bottom of layer l+1
    assert np.array_equal(np.unique(full), np.arange(V.node_count))
    n = 4
    for b in range(n * n):
        top = full[0::nz][:, b * n + 1]       # layer 0, vertical dof 1 (top vertex)
        bot_next = full[1::nz][:, b * n + 0]  # layer 1, vertical dof 0
        assert np.array_equal(top, bot_next)


# ---- 3. independent dense evaluation of the element tensors
def dense_element_matrix(p, X, alpha, beta):
    """B^T D B with explicitly tabulated 3-D basis gradients (NumPy)."""
    el = interval_element(p)
    n, B, D, wq, xq = el.ndof, el.B, el.D, el.wq, el.xq
    nd = n ** 3
    A = np.zeros((nd, nd))
    for qx in range(n):
        for qy in range(n):
            for qz in range(n):
                xi = np.array([xq[qx], xq[qy], xq[qz]])
                J = np.zeros((3, 3))
                for bx in (0, 1):
                    for by in (0, 1):
                        for bz in (0, 1):
                            s = [(xi[d] if b else 1 - xi[d]) for d, b in enumerate((bx, by, bz))]
                            ds = [(1.0 if b else -1.0) for b in (bx, by, bz)]
                            g = np.array([ds[0] * s[1] * s[2], s[0] * ds[1] * s[2], s[0] * s[1] * ds[2]])
                            J += np.outer(X[(bx * 2 + by) * 2 + bz], g)
                K = np.linalg.inv(J)
                dw = abs(np.linalg.det(J)) * wq[qx] * wq[qy] * wq[qz]
                gref = np.zeros((nd, 3))
                val = np.zeros(nd)
                for ax in range(n):
                    for ay in range(n):
                        for az in range(n):
                            i = (ax * n + ay) * n + az
                            gref[i] = [D[qx, ax] * B[qy, ay] * B[qz, az],
                                       B[qx, ax] * D[qy, ay] * B[qz, az],
                                       B[qx, ax] * B[qy, ay] * D[qz, az]]
                            val[i] = B[qx, ax] * B[qy, ay] * B[qz, az]
                G = gref @ K                      # physical gradients
                A += dw * (alpha * G @ G.T + beta * np.outer(val, val))
    return A


@pytest.mark.parametrize("p", [1, 2, 3])
def test_cell_kernels_vs_dense_numpy(oracle, p):
    rng = np.random.default_rng(p)
    X = np.array([[bx, by, bz] for bx in (0, 1) for by in (0, 1) for bz in (0, 1)], dtype=float)
    X = X * [1.0, 0.7, 1.3] + 0.12 * rng.standard_normal((8, 3))      # non-affine hex
    el = interval_element(p)
    for alpha, beta in [(1.0, 0.0), (0.0, 1.0), (1.0, 1.0)]:
        Aref = dense_element_matrix(p, X, alpha, beta)
        A = oracle.cell_matrix(el, X.ravel(), alpha, beta)
        scale = np.abs(Aref).max()
        assert np.abs(A - Aref).max() < 1e-13 * scale
        w = rng.standard_normal(el.ndof ** 3)
        y = oracle.cell_action(el, X.ravel(), w, alpha, beta)
        assert np.abs(y - Aref @ w).max() < 1e-12 * scale * np.abs(w).max()
        assert np.abs(A - A.T).max() < 1e-13 * scale              # symmetry


def test_tabulation_tables():
    for p in range(1, 6):
        el = interval_element(p)
        assert np.allclose(el.B.sum(axis=1), 1.0, atol=1e-14)       # partition of unity
        assert np.allclose(el.D.sum(axis=1), 0.0, atol=1e-12)
        assert abs(el.wq.sum() - 1.0) < 1e-15
        # nodes are GLL points in entity order: endpoints first
        assert el.nodes[0] == 0.0 and el.nodes[1] == 1.0
        assert np.allclose(np.sort(el.nodes), gll_points(p))
        # the rule integrates degree 2p+1 exactly
        x, w = gauss_legendre(p + 1)
        assert abs((w * x ** (2 * p + 1)).sum() - 1.0 / (2 * p + 2)) < 1e-15


# ---- 4. Firedrake regression invariants
def assemble_matrix(oracle, mesh, V, p, alpha, beta, row_lgmap=None, col_lgmap=None):
    rowptr, colidx = oracle.build_sparsity(V.node_count, V.cell_node_map, V.offset, mesh.nz)
    vals = np.zeros(len(colidx))
    oracle.matrix_extruded(interval_element(p), 0, mesh.num_base_cells, [0, mesh.layers], rowptr,
                           colidx, vals, mesh.coordinates, V.cell_node_map, V.offset,
                           mesh.coord_map, mesh.coord_offset, row_lgmap, col_lgmap, alpha, beta)
    return rowptr, colidx, vals


def action(oracle, mesh, V, p, x, alpha, beta):
    y = np.zeros(V.node_count)
    oracle.action_extruded(interval_element(p), 0, mesh.num_base_cells, [0, mesh.layers], y,
                           mesh.coordinates, x, V.cell_node_map, V.offset, mesh.coord_map,
                           mesh.coord_offset, alpha=alpha, beta=beta)
    return y


@pytest.mark.parametrize("p", [1, 2, 3])
def test_mass_action_integrates_x(oracle, p):
    """reference tests/firedrake/regression/test_assemble.py:60-73:
    sum(assemble(action(M, f))) == 0.5 for f = x on the unit domain (1e-12)."""
    mesh = ExtrudedHexMesh(3, 4, 2, warp=0.04)
    V = mesh.function_space(p)
    f = V.dof_coordinates()[:, 0].copy()
    # integral of x over the unit cube, warp vanishes on the boundary but moves
    # interior mass: compare against the warped-mesh value from quadrature
    y = action(oracle, mesh, V, p, f, 0.0, 1.0)
    ref_mesh = ExtrudedHexMesh(3, 4, 2, warp=0.0)
    Vr = ref_mesh.function_space(p)
    yr = action(oracle, ref_mesh, Vr, p, Vr.dof_coordinates()[:, 0].copy(), 0.0, 1.0)
    assert abs(yr.sum() - 0.5) < 1e-12
    # divergence theorem: int x dV = int x^2/2 n_x dS, and the warp is zero on
    # the boundary, so the warped value is 0.5 as well
    assert abs(y.sum() - 0.5) < 1e-12


@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("bcs", [False, True])
def test_assembled_matvec_equals_matfree_action(oracle, p, bcs):
    """reference tests/firedrake/regression/test_matrix_free.py:98-127:
    A.mult(x) of the assembled matrix == the matrix-free action, with and
    without Dirichlet conditions.  BC semantics (SURVEY.md section 8a row
    A6/A10/A12): masked rows/cols (lgmap = -1) are dropped, then the diagonal
    is set to 1; matrix-free: zero the BC entries of x, apply, then copy x on
    the BC rows."""
    mesh = ExtrudedHexMesh(3, 3, 3, warp=0.05, permute_seed=2)
    V = mesh.function_space(p)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(V.node_count)
    lg = None
    if bcs:
        bnodes = np.union1d(V.boundary_nodes("bottom"), V.boundary_nodes("top"))
        lg = np.arange(V.node_count, dtype=np.int32)
        lg[bnodes] = -1
    rowptr, colidx, vals = assemble_matrix(oracle, mesh, V, p, 1.0, 1.0, lg, lg)
    A = dense(rowptr, colidx, vals, V.node_count)
    if bcs:
        A[bnodes, bnodes] = 1.0                       # set_local_diagonal_entries
        xin = x.copy()
        xin[bnodes] = 0.0
        y = action(oracle, mesh, V, p, xin, 1.0, 1.0)
        y[bnodes] = x[bnodes]
    else:
        y = action(oracle, mesh, V, p, x, 1.0, 1.0)
    assert np.abs(A @ x - y).max() < 1e-12 * np.abs(y).max()
    assert np.abs(A - A.T).max() < 1e-13 * np.abs(A).max()


def test_poisson_strong_bcs_extrusion_exact(oracle):
    """reference tests/firedrake/extrusion/test_poisson_strong_bcs_extrusion.py:
    21-52: u = 42 z is in the space, so (K u)_i = 0 on every row that is not on
    the bottom/top Dirichlet boundary (natural BCs on the sides)."""
    for p in (1, 2, 3):
        mesh = ExtrudedHexMesh(3, 2, 4, warp=0.0)
        V = mesh.function_space(p)
        u = 42.0 * V.dof_coordinates()[:, 2]
        r = action(oracle, mesh, V, p, u, 1.0, 0.0)
        interior = np.ones(V.node_count, bool)
        interior[np.union1d(V.boundary_nodes("bottom"), V.boundary_nodes("top"))] = False
        assert np.abs(r[interior]).max() < 1e-11


def test_poisson_nullspace_and_rowsums(oracle):
    mesh = ExtrudedHexMesh(3, 3, 2, warp=0.06)
    V = mesh.function_space(2)
    rowptr, colidx, vals = assemble_matrix(oracle, mesh, V, 2, 1.0, 0.0)
    A = dense(rowptr, colidx, vals, V.node_count)
    assert np.abs(A.sum(axis=1)).max() < 1e-12 * np.abs(A).max()
    M = dense(*assemble_matrix(oracle, mesh, V, 2, 0.0, 1.0), V.node_count)
    assert abs(M.sum() - 1.0) < 1e-13                     # 1^T M 1 = |Omega|
    # diagonal == Mat diagonal (reference test_assemble.py:188-206): action on e_i
    e = np.zeros(V.node_count)
    e[7] = 1.0
    assert abs(action(oracle, mesh, V, 2, e, 1.0, 0.0)[7] - A[7, 7]) < 1e-13 * abs(A[7, 7])


def test_sparsity_semantics(oracle):
    """reference pyop2/sparsity.pyx:106-160, 198-204, 347-373: union of
    rowmap x colmap over cells and layers, diagonal always allocated."""
    mesh = ExtrudedHexMesh(2, 1, 3)
    V = mesh.function_space(1)
    rowptr, colidx = oracle.build_sparsity(V.node_count, V.cell_node_map, V.offset, mesh.nz)
    full = V.full_cell_node_list()
    pat = set()
    for row in full:
        for i in row:
            for j in row:
                pat.add((int(i), int(j)))
    for r in range(V.node_count):
        pat.add((r, r))
    got = {(r, int(c)) for r in range(V.node_count) for c in colidx[rowptr[r]:rowptr[r + 1]]}
    assert got == pat
    # columns sorted within rows (binary search in MatSetValues restatement)
    for r in range(V.node_count):
        c = colidx[rowptr[r]:rowptr[r + 1]]
        assert np.all(np.diff(c) > 0)


def test_config1_p1_triangles(oracle):
    """BASELINE.json configs[0]: Poisson CG1 on UnitSquareMesh(64, 64),
    bilinear form -> CSR on the CPU.  Pins: 8192 cells / 4225 dofs, symmetry,
    constants in the null space, x^T K x = |grad x|^2 area for linear x, and
    matrix-vs-action agreement."""
    mesh = UnitSquareTriMesh(64, 64)
    assert mesh.num_cells == 8192 and mesh.node_count == 4225
    tab = oracle.tri_table_fiat()
    rowptr, colidx = oracle.build_sparsity(mesh.node_count, mesh.cell_node_map)
    vals = np.zeros(len(colidx))
    oracle.tri_matrix("laplace", 0, mesh.num_cells, rowptr, colidx, vals, mesh.coordinates,
                      mesh.cell_node_map, tab)
    import scipy.sparse as sp
    K = sp.csr_matrix((vals, colidx, rowptr), shape=(4225, 4225))
    assert abs(K - K.T).max() < 1e-14
    assert np.abs(K @ np.ones(4225)).max() < 1e-12
    u = 3 * mesh.coordinates[:, 0] - 2 * mesh.coordinates[:, 1]
    assert abs(u @ (K @ u) - 13.0) < 1e-11
    y = np.zeros(4225)
    x = np.random.default_rng(0).standard_normal(4225)
    oracle.tri_action(0, mesh.num_cells, y, mesh.coordinates, x, mesh.cell_node_map, tab)
    assert np.abs(y - K @ x).max() < 1e-12 * np.abs(y).max()
    # mass: 1^T M 1 = 1, and the P1 mass matrix of the reference golden test
    vals[:] = 0
    oracle.tri_matrix("mass", 0, mesh.num_cells, rowptr, colidx, vals, mesh.coordinates,
                      mesh.cell_node_map, tab)
    assert abs(vals.sum() - 1.0) < 1e-13


def test_dg_advection_oracle_invariants(oracle):
    """DG advection restatement (oracle/dg_advection.c): with q == q_in and a
    divergence-free velocity the cell, exterior-facet and interior-facet terms
    cancel exactly (divergence theorem) even on distorted quads; interior
    fluxes cancel in the sum over test functions (mass conservation,
    reference tests/firedrake/regression/test_dg_advection.py:62-75)."""
    from firedrake_b200.utility_meshes import QuadMesh
    m = QuadMesh(8, 6)
    rng = np.random.default_rng(0)
    X = m.coordinates
    inner = (X[:, 0] > 1e-9) & (X[:, 0] < 1 - 1e-9) & (X[:, 1] > 1e-9) & (X[:, 1] < 1 - 1e-9)
    X[inner] += 0.02 * rng.standard_normal((inner.sum(), 2))
    u = np.stack([0.5 - X[:, 1], X[:, 0] - 0.5], axis=1)
    r = oracle.dg_rhs(m, np.ones(m.num_cells * 4), u, dt=0.1, q_in=1.0)
    assert np.abs(r).max() < 1e-15
    # sum_i L1_i = -dt * (boundary flux): independent of interior cell values
    q = 1 + rng.random(m.num_cells * 4)
    s0 = oracle.dg_rhs(m, q, u, dt=0.1).sum()
    bcells = set(m.ext_facet_cells.tolist())
    q2 = q.copy()
    for c in range(m.num_cells):
        if c not in bcells:
            q2[4 * c:4 * c + 4] += 0.7
    assert abs(oracle.dg_rhs(m, q2, u, dt=0.1).sum() - s0) < 1e-13
    # facet bookkeeping of the synthetic mesh
    assert len(m.int_facet_cells) == 7 * 6 + 8 * 5 and len(m.ext_facet_cells) == 2 * (8 + 6)


def test_dg_advection_oracle_analytic_single_cell(oracle):
    """One unit-square cell, u = (1, 0), q = x, q_in = 0: the weak form of
    q_t = -div(u q) = -1 gives L1_i = -dt * int phi_i = -dt/4 for each DQ1 basis
    function (integration by parts of the cell term against the outflow term)."""
    from firedrake_b200.fiat_lite import interval_element
    from firedrake_b200.utility_meshes import QuadMesh
    m = QuadMesh(1, 1)
    el = interval_element(1, 2, "gl")                 # DQ1 nodes = the two Gauss points
    xn = el.nodes
    q = np.array([xn[ax] for ax in range(2) for ay in range(2)])   # q = x at node (ax, ay)
    u = np.tile(np.array([1.0, 0.0]), (4, 1))
    for nq in (2, 3, 4):
        r = oracle.dg_rhs(m, q, u, dt=0.5, q_in=0.0, nq=nq)
        np.testing.assert_allclose(r, -0.5 / 4 * np.ones(4), rtol=0, atol=1e-15)
    # and with the inflow value q_in = 2 the x = 0 face adds +dt * int phi_i(0, y) * 2 dy
    B0, _ = el.tabulate([0.0])
    r = oracle.dg_rhs(m, q, u, dt=0.5, q_in=2.0)
    expect = np.array([-0.125 + 0.5 * 2.0 * B0[0, ax] * 0.5 for ax in range(2) for ay in range(2)])
    np.testing.assert_allclose(r, expect, atol=1e-15)


def test_gll_points_known_values():
    """GLL nodes for p = 2, 3, 4 against their closed forms on [0, 1]."""
    np.testing.assert_allclose(gll_points(2), [0.0, 0.5, 1.0], atol=1e-16)
    np.testing.assert_allclose(gll_points(3), [0.0, 0.5 - 0.5 / np.sqrt(5), 0.5 + 0.5 / np.sqrt(5), 1.0], atol=1e-15)
    np.testing.assert_allclose(gll_points(4), [0.0, 0.5 - 0.5 * np.sqrt(3 / 7), 0.5, 0.5 + 0.5 * np.sqrt(3 / 7), 1.0],
                               atol=1e-15)
    # Lagrange property of the tabulation at its own nodes
    for p in (1, 2, 3, 4, 5):
        el = interval_element(p)
        B, _ = el.tabulate(el.nodes)
        np.testing.assert_allclose(B, np.eye(p + 1), atol=1e-13)


@pytest.mark.parametrize("p,levels,rate", [(1, (4, 6), 1.9), (2, (2, 4), 2.9), (3, (1, 3), 3.9)])
def test_helmholtz_convergence_rates(oracle, p, levels, rate):
    """reference tests/firedrake/extrusion/test_helmholtz_scalar.py:8-36 (quadrilateral=True):
    -lap u + u = f with natural conditions, u = cos(2 pi x) cos(2 pi y) cos(2 pi z), f and the
    exact solution interpolated into the space, L2 error rate > p + 0.9.  CG1 on the reference's
    16^3 / 32^3; CG2 and CG3 one level coarser than the reference (their rates are already
    asymptotic there) to keep the CPU suite short."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    errs = []
    for ii in range(*levels):
        n = 2 ** ii
        mesh = ExtrudedHexMesh(n, n, n, permute_seed=ii)
        V = mesh.function_space(p)
        P = V.dof_coordinates()
        exact = np.cos(2 * np.pi * P[:, 0]) * np.cos(2 * np.pi * P[:, 1]) * np.cos(2 * np.pi * P[:, 2])
        f = (1 + 12 * np.pi ** 2) * exact
        rowptr, colidx, vals = assemble_matrix(oracle, mesh, V, p, 1.0, 1.0, None, None)
        A = sp.csr_matrix((vals, colidx, rowptr), shape=(V.node_count, V.node_count))
        b = action(oracle, mesh, V, p, f, 0.0, 1.0)                       # inner(f, v)*dx
        u, info = spla.cg(A, b, rtol=1e-12, maxiter=5000)
        assert info == 0
        e = u - exact
        errs.append(np.sqrt(e @ action(oracle, mesh, V, p, e, 0.0, 1.0)))  # ||e||_L2 through the mass form
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(len(errs) - 1)]
    assert min(rates) > rate, (errs, rates)
