"""GPU parity of the explicit-matrix path:
device sparsity construction, rank-2 global kernel with MatSetValues-style
scatter, BC-masked lgmaps + diagonal, SpMV -- against the oracle's CSR
assembly on the same inputs.  Tolerance 1e-12 * max|A| entrywise."""
import numpy as np
import pytest

from firedrake_b200 import op2
from firedrake_b200.fiat_lite import interval_element
from firedrake_b200.utility_meshes import ExtrudedHexMesh

pytestmark = pytest.mark.gpu


def setup(mesh, p):
    V = mesh.function_space(p)
    cells = op2.ExtrudedSet(op2.Set(mesh.num_base_cells), mesh.layers)
    nodes = op2.Set(V.node_count)
    vnodes = op2.Set(mesh.coord_space.node_count)
    m0 = op2.Map(cells, nodes, V.arity, V.cell_node_map, offset=V.offset)
    m1 = op2.Map(cells, vnodes, 8, mesh.coord_map, offset=mesh.coord_offset)
    X = op2.Dat(op2.DataSet(vnodes, 3), mesh.coordinates)
    return V, cells, nodes, m0, m1, X


def oracle_matrix(oracle, mesh, V, p, alpha, beta, lg=None):
    rowptr, colidx = oracle.build_sparsity(V.node_count, V.cell_node_map, V.offset, mesh.nz)
    vals = np.zeros(len(colidx))
    oracle.matrix_extruded(interval_element(p), 0, mesh.num_base_cells, [0, mesh.layers], rowptr,
                           colidx, vals, mesh.coordinates, V.cell_node_map, V.offset,
                           mesh.coord_map, mesh.coord_offset, lg, lg, alpha, beta)
    return rowptr, colidx, vals


@pytest.mark.parametrize("p", [1, 2, 3])
def test_sparsity_matches_reference_semantics(engine, oracle, p):
    mesh = ExtrudedHexMesh(4, 3, 5, permute_seed=1)
    V, cells, nodes, m0, m1, X = setup(mesh, p)
    mat = op2.Mat(op2.Sparsity((nodes, nodes), [(m0, m0, None)]))
    rowptr, colidx, _ = mat.csr()
    ro, co = oracle.build_sparsity(V.node_count, V.cell_node_map, V.offset, mesh.nz)
    assert np.array_equal(rowptr, ro) and np.array_equal(colidx, co)


@pytest.mark.parametrize("p,alpha,beta", [(1, 1.0, 0.0), (2, 1.0, 1.0), (3, 1.0, 0.0), (3, 0.0, 1.0), (4, 1.0, 1.0)])
def test_matrix_matches_oracle(engine, oracle, matrix_kernel, p, alpha, beta):
    mesh = ExtrudedHexMesh(3, 3, 4, warp=0.05, permute_seed=2) if p < 4 else \
        ExtrudedHexMesh(2, 2, 3, warp=0.05, permute_seed=2)
    V, cells, nodes, m0, m1, X = setup(mesh, p)
    mat = op2.Mat(op2.Sparsity((nodes, nodes), [(m0, m0, None)]))
    k = op2.Kernel("helmholtz", degree=p, alpha=alpha, beta=beta, rank=2)
    mat.zero()
    op2.par_loop(k, cells, mat(op2.INC, (m0, m0)), X(op2.READ, m1))
    mat.assemble()
    _, _, vals = mat.csr()
    _, _, vo = oracle_matrix(oracle, mesh, V, p, alpha, beta)
    assert np.abs(vals - vo).max() < 1e-12 * np.abs(vo).max()


def test_bc_lgmaps_diagonal_and_matvec(engine, oracle, matrix_kernel):
    """reference tests/firedrake/regression/test_matrix_free.py:98-127 on the
    device: assembled (BC rows/cols dropped, unit diagonal) A.mult(x) equals the
    matrix-free action with the BC protocol of matrix_free/operators.py:211-242."""
    p = 2
    mesh = ExtrudedHexMesh(4, 3, 4, warp=0.04)
    V, cells, nodes, m0, m1, X = setup(mesh, p)
    bnodes = np.union1d(V.boundary_nodes("bottom"), V.boundary_nodes("top")).astype(np.int32)
    lg = np.arange(V.node_count, dtype=np.int32)
    lg[bnodes] = -1
    mat = op2.Mat(op2.Sparsity((nodes, nodes), [(m0, m0, None)]))
    k2 = op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=1.0, rank=2)
    op2.par_loop(k2, cells, mat(op2.INC, (m0, m0), lgmaps=(lg, lg)), X(op2.READ, m1))
    mat.set_local_diagonal_entries(bnodes, 1.0)
    mat.assemble()
    _, _, vals = mat.csr()
    ro, co, vo = oracle_matrix(oracle, mesh, V, p, 1.0, 1.0, lg)
    import scipy.sparse as sp
    Ao = sp.csr_matrix((vo, co, ro), shape=(V.node_count,) * 2).tolil()
    for b in bnodes:
        Ao[b, b] = 1.0
    Ao = Ao.tocsr()
    Ag = sp.csr_matrix((vals, co, ro), shape=(V.node_count,) * 2)
    assert abs(Ag - Ao).max() < 1e-12 * abs(Ao).max()
    # SpMV vs matrix-free action with BCs
    rng = np.random.default_rng(3)
    xv = rng.standard_normal(V.node_count)
    x = op2.Dat(nodes, xv.copy())
    y = op2.Dat(nodes)
    mat.mult(x, y)
    xin = op2.Dat(nodes, xv.copy())
    xin.zero(op2.Subset(nodes, bnodes))                    # bc.zero(x)
    ymf = op2.Dat(nodes)
    k1 = op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=1.0)
    op2.par_loop(k1, cells, ymf(op2.INC, m0), X(op2.READ, m1), xin(op2.READ, m0))
    ymf_h = ymf.data_ro.copy()
    ymf_h[bnodes] = xv[bnodes]                             # bc.set(y, x)
    assert np.abs(y.data_ro - ymf_h).max() < 1e-12 * np.abs(ymf_h).max()
    assert np.abs(y.data_ro - Ao @ xv).max() < 1e-12 * np.abs(ymf_h).max()


def test_matrix_properties_at_scale(engine):
    """Size-independent invariants on a larger mesh: symmetry, zero row sums
    (constants in the null space), 1^T M 1 = |Omega|."""
    p = 2
    mesh = ExtrudedHexMesh(10, 8, 12, warp=0.05)
    V, cells, nodes, m0, m1, X = setup(mesh, p)
    mat = op2.Mat(op2.Sparsity((nodes, nodes), [(m0, m0, None)]))
    op2.par_loop(op2.Kernel("helmholtz", degree=p, rank=2), cells, mat(op2.INC, (m0, m0)), X(op2.READ, m1))
    ro, co, va = mat.csr()
    import scipy.sparse as sp
    A = sp.csr_matrix((va, co, ro), shape=(V.node_count,) * 2)
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    assert np.abs(A @ np.ones(V.node_count)).max() < 1e-11 * abs(A).max()
    mat.zero()
    op2.par_loop(op2.Kernel("helmholtz", degree=p, alpha=0.0, beta=1.0, rank=2), cells,
                 mat(op2.INC, (m0, m0)), X(op2.READ, m1))
    _, _, vm = mat.csr()
    assert abs(vm.sum() - 1.0) < 1e-12
