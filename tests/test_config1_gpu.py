"""BASELINE.json configs[0] on the device: Poisson CG1 on UnitSquareMesh(64,64)
(8192 P1 triangles, 4225 dofs): bilinear form -> CSR and the action, against
the oracle (which is pinned to the PyOP2 golden mass matrix / RHS)."""
import numpy as np
import pytest

from firedrake_b200 import op2
from firedrake_b200.utility_meshes import UnitSquareTriMesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (0.0, 1.0), (1.0, 1.0)])
def test_p1_matrix_and_action(engine, oracle, alpha, beta):
    mesh = UnitSquareTriMesh(64, 64)
    cells = op2.Set(mesh.num_cells)
    nodes = op2.Set(mesh.node_count)
    m = op2.Map(cells, nodes, 3, mesh.cell_node_map)
    X = op2.Dat(op2.DataSet(nodes, 2), mesh.coordinates)
    mat = op2.Mat(op2.Sparsity((nodes, nodes), [(m, m, None)]))
    bn = mesh.boundary_nodes()
    lg = np.arange(mesh.node_count, dtype=np.int32)
    lg[bn] = -1
    k2 = op2.Kernel("helmholtz", degree=1, alpha=alpha, beta=beta, rank=2, cell="triangle")
    op2.par_loop(k2, cells, mat(op2.INC, (m, m), lgmaps=(lg, lg)), X(op2.READ, m))
    mat.set_local_diagonal_entries(bn, 1.0)
    ro, co, va = mat.csr()
    # oracle: same table (FIAT order, 3-point rule)
    tab = oracle.tri_table_fiat()
    rowptr, colidx = oracle.build_sparsity(mesh.node_count, mesh.cell_node_map)
    assert np.array_equal(ro, rowptr) and np.array_equal(co, colidx)
    vals = np.zeros(len(colidx))
    if alpha != 0.0:
        oracle.tri_matrix("laplace", 0, mesh.num_cells, rowptr, colidx, vals, mesh.coordinates,
                          mesh.cell_node_map, tab, lg, lg, beta=beta / alpha if alpha else 0.0)
        vals *= alpha
    else:
        oracle.tri_matrix("mass", 0, mesh.num_cells, rowptr, colidx, vals, mesh.coordinates,
                          mesh.cell_node_map, tab, lg, lg)
        vals *= beta
    import scipy.sparse as sp
    Ao = sp.csr_matrix((vals, colidx, rowptr), shape=(mesh.node_count,) * 2).tolil()
    for b in bn:
        Ao[b, b] = 1.0
    Ag = sp.csr_matrix((va, co, ro), shape=(mesh.node_count,) * 2)
    assert abs(Ag - Ao.tocsr()).max() < 1e-13 * abs(Ao).max()
    # action (no BCs) == unconstrained matrix times x
    x = op2.Dat(nodes, np.random.default_rng(1).standard_normal(mesh.node_count))
    y = op2.Dat(nodes)
    k1 = op2.Kernel("helmholtz", degree=1, alpha=alpha, beta=beta, cell="triangle")
    op2.par_loop(k1, cells, y(op2.INC, m), X(op2.READ, m), x(op2.READ, m))
    yo = np.zeros(mesh.node_count)
    if alpha != 0.0:
        oracle.tri_action(0, mesh.num_cells, yo, mesh.coordinates, x.data_ro.copy(), mesh.cell_node_map,
                          tab, alpha=alpha, beta=beta)
    else:
        oracle.tri_action(0, mesh.num_cells, yo, mesh.coordinates, x.data_ro.copy(), mesh.cell_node_map,
                          tab, alpha=0.0, beta=1.0)
        yo *= beta
    assert np.abs(y.data_ro - yo).max() < 1e-12 * np.abs(yo).max()


def test_pyop2_golden_arrays_on_the_device(engine):
    """The reference's own value-level pins for this path -- the 2-triangle mass
    matrix and RHS of tests/pyop2/test_matrices.py:462-503 (fixture
    tests/golden/pyop2_test_matrices.json) -- checked directly against the CUDA
    path.  The reference kernels hard-code an 8-digit quadrature table, so its
    RHS differs from the exact integral by ~5e-8 relative: tolerance 1e-6 for
    the RHS, the reference's own 1e-5 for the matrix."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pyop2_test_matrices.json")))
    cells, nodes = op2.Set(2), op2.Set(4)
    # the golden kernels order the P1 basis {x, y, 1-x-y}; FIAT order is {1-x-y, x, y}
    cmap = np.array(g["elem_node_map"], dtype=np.int32)[:, [2, 0, 1]]
    m = op2.Map(cells, nodes, 3, cmap)
    X = op2.Dat(op2.DataSet(nodes, 2), np.array(g["coords"]))
    mat = op2.Mat(op2.Sparsity((nodes, nodes), [(m, m, None)]))
    op2.par_loop(op2.Kernel("helmholtz", degree=1, alpha=0.0, beta=1.0, rank=2, cell="triangle"), cells,
                 mat(op2.INC, (m, m)), X(op2.READ, m))
    np.testing.assert_allclose(mat.values, np.array(g["expected_matrix"]), atol=g["expected_matrix_eps"])
    f = op2.Dat(nodes, np.array(g["f"]))
    b = op2.Dat(nodes)
    op2.par_loop(op2.Kernel("helmholtz", degree=1, alpha=0.0, beta=1.0, cell="triangle"), cells,
                 b(op2.INC, m), X(op2.READ, m), f(op2.READ, m))
    np.testing.assert_allclose(b.data_ro, np.array(g["expected_rhs"]), rtol=1e-6)
