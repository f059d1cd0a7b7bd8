"""GPU parity: the sm_100a Helmholtz-family action kernel, called through the
C ABI (firedrake_b200.op2 -> fdb_kernel_call), against the CPU oracle on the
same seeded inputs.  Tolerance: 1e-12 relative in the max norm scaled by
max|y| (different summation order, -ffast-math on the
CPU side)."""
import numpy as np
import pytest

from firedrake_b200 import op2
from firedrake_b200.fiat_lite import interval_element
from firedrake_b200.utility_meshes import ExtrudedHexMesh

pytestmark = pytest.mark.gpu

TOL = 1e-12


def build(mesh, p, cdim=1, seed=1234):
    V = mesh.function_space(p)
    cells = op2.ExtrudedSet(op2.Set(mesh.num_base_cells), mesh.layers)
    nodes = op2.Set(V.node_count)
    vnodes = op2.Set(mesh.coord_space.node_count)
    m0 = op2.Map(cells, nodes, V.arity, V.cell_node_map, offset=V.offset)
    m1 = op2.Map(cells, vnodes, 8, mesh.coord_map, offset=mesh.coord_offset)
    rng = np.random.default_rng(seed)
    shape = (V.node_count,) if cdim == 1 else (V.node_count, cdim)
    x = op2.Dat(op2.DataSet(nodes, cdim), rng.standard_normal(shape))
    y = op2.Dat(op2.DataSet(nodes, cdim))
    X = op2.Dat(op2.DataSet(vnodes, 3), mesh.coordinates)
    return V, cells, m0, m1, x, y, X


def oracle_action(oracle, mesh, V, p, xdata, cdim=1, alpha=1.0, beta=0.0, start=0, end=None):
    yo = np.zeros_like(xdata)
    oracle.action_extruded(interval_element(p), start, mesh.num_base_cells if end is None else end,
                           [0, mesh.layers], yo, mesh.coordinates, np.ascontiguousarray(xdata),
                           V.cell_node_map, V.offset, mesh.coord_map, mesh.coord_offset,
                           cdim=cdim, alpha=alpha, beta=beta)
    return yo


def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (1.0, 1.0), (0.0, 1.0)])
def test_action_matches_oracle(engine, oracle, p, alpha, beta):
    mesh = ExtrudedHexMesh(5, 4, 7, warp=0.05, permute_seed=0)
    V, cells, m0, m1, x, y, X = build(mesh, p)
    k = op2.Kernel("helmholtz", degree=p, alpha=alpha, beta=beta)
    op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    yo = oracle_action(oracle, mesh, V, p, x.data_ro, alpha=alpha, beta=beta)
    assert relerr(y.data_ro, yo) < TOL


@pytest.mark.parametrize("p", [1, 3])
def test_inc_semantics_and_ranges(engine, oracle, p):
    """The output is INCREMENTED (caller zeroes it), and [start, end) splits
    (core part / owned part, pyop2/parloop.py:250-253) compose."""
    mesh = ExtrudedHexMesh(6, 3, 5, warp=0.03)
    V, cells, m0, m1, x, y, X = build(mesh, p)
    y0 = np.random.default_rng(7).standard_normal(V.node_count)
    y.data[:] = y0
    k = op2.Kernel("helmholtz", degree=p)
    core = 7
    cells2 = op2.ExtrudedSet(op2.Set((core, mesh.num_base_cells, mesh.num_base_cells)), mesh.layers)
    m0b = op2.Map(cells2, m0.toset, V.arity, V.cell_node_map, offset=V.offset)
    m1b = op2.Map(cells2, m1.toset, 8, mesh.coord_map, offset=mesh.coord_offset)
    op2.par_loop(k, cells2, y(op2.INC, m0b), X(op2.READ, m1b), x(op2.READ, m0b))
    yo = oracle_action(oracle, mesh, V, p, x.data_ro)
    assert relerr(y.data_ro - y0, yo) < 1e-11


@pytest.mark.parametrize("p", [2, 3])
def test_coloured_scatter_is_deterministic(engine, oracle, p):
    mesh = ExtrudedHexMesh(7, 6, 9, warp=0.05, permute_seed=3)
    V, cells, m0, m1, x, y, X = build(mesh, p)
    k = op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=0.5)
    outs = []
    for _ in range(2):
        y.zero()
        op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0), scatter="coloured")
        outs.append(y.data_ro.copy())
    assert np.array_equal(outs[0], outs[1])          # bit-reproducible
    yo = oracle_action(oracle, mesh, V, p, x.data_ro, alpha=1.0, beta=0.5)
    assert relerr(outs[0], yo) < TOL
    # atomics vs colouring: the race oracle of SURVEY.md section 5
    y.zero()
    op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    assert relerr(y.data_ro, outs[0]) < TOL


def test_host_pointer_mode(engine, oracle):
    """Drop-in mode: host pointers + dat_version through the mirror cache."""
    p = 3
    mesh = ExtrudedHexMesh(4, 4, 6, warp=0.05)
    V, cells, m0, m1, x, y, X = build(mesh, p)
    k = op2.Kernel("helmholtz", degree=p)
    gk = op2.GlobalKernel(k, [m0, m1], extruded=True)
    loop = op2.Parloop(gk, cells, [y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0)], location="host")
    loop()
    yo = oracle_action(oracle, mesh, V, p, x.data_ro)
    assert relerr(y.data_ro, yo) < TOL
    # a host write bumps dat_version -> re-upload; a second call accumulates
    x.data[:] *= 2.0
    loop()
    assert relerr(y.data_ro, 3.0 * yo) < TOL


def test_map_generation_defeats_a_recycled_address(engine, oracle):
    """Advisor (round 1): the engine mirrors host map buffers by ADDRESS.  A new Map whose buffer
    lands on a freed Map's address must not hit the old mirror: maps carry a generation id
    (fdb_call_args.map_versions).  Emulated by rewriting the buffer in place under a fresh id."""
    p = 2
    mesh = ExtrudedHexMesh(4, 3, 5, warp=0.05)
    V, cells, m0, m1, x, y, X = build(mesh, p)
    k = op2.Kernel("helmholtz", degree=p)
    gk = op2.GlobalKernel(k, [m0, m1], extruded=True)
    op2.Parloop(gk, cells, [y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0)], location="host")()
    yo = oracle_action(oracle, mesh, V, p, x.data_ro)
    assert relerr(y.data_ro, yo) < TOL
    # "another mesh at the same address": the same cells in a different order, same buffers
    perm = np.random.default_rng(0).permutation(mesh.num_base_cells)
    m0.values_with_halo[:] = m0.values_with_halo[perm]
    m1.values_with_halo[:] = m1.values_with_halo[perm]
    m0._generation, m1._generation = next(op2._generations), next(op2._generations)
    y2 = op2.Dat(y.dataset)
    # only the first two columns: the result depends on WHICH cells they now are
    op2.Parloop(gk, cells, [y2(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0)], location="host")._compute((0, 2))
    yo2 = np.zeros_like(yo)
    from firedrake_b200.fiat_lite import interval_element
    oracle.action_extruded(interval_element(p), 0, 2, [0, mesh.layers], yo2, mesh.coordinates,
                           np.ascontiguousarray(x.data_ro), np.ascontiguousarray(m0.values_with_halo), V.offset,
                           np.ascontiguousarray(m1.values_with_halo), mesh.coord_offset)
    assert relerr(y2.data_ro, yo2) < TOL


def test_vector_space_aos(engine, oracle):
    """cdim = 3, node-major / component-fastest (vector Helmholtz, config 4)."""
    p = 2
    mesh = ExtrudedHexMesh(4, 3, 5, warp=0.05)
    V, cells, m0, m1, x, y, X = build(mesh, p, cdim=3)
    k = op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=1.0, cdim=3)
    op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    yo = oracle_action(oracle, mesh, V, p, x.data_ro, cdim=3, alpha=1.0, beta=1.0)
    assert relerr(y.data_ro, yo) < TOL


@pytest.mark.parametrize("cdim", [2, 3])
def test_vector_space_degree3(engine, oracle, cdim):
    """Degree 3 with several components per node: the geometry coefficients parked in shared memory
    at component 0 serve the later components of the same cells (action_hex.cu, STASH)."""
    p = 3
    mesh = ExtrudedHexMesh(3, 4, 11, warp=0.05, permute_seed=2)
    V, cells, m0, m1, x, y, X = build(mesh, p, cdim=cdim)
    k = op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=0.5, cdim=cdim)
    op2.par_loop(k, cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    yo = oracle_action(oracle, mesh, V, p, x.data_ro, cdim=cdim, alpha=1.0, beta=0.5)
    assert relerr(y.data_ro, yo) < TOL


@pytest.mark.parametrize("ws", [1, 3])
def test_warp_specialised_kernel(ws):
    """The opt-in warp-specialised degree-3 kernel (action_hex_ws.cuh; FDB_WS is read once per process,
    hence the worker): mass + stiffness, a ragged last unit, units straddling two columns, a subset."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FDB_WS=str(ws))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "_ws_worker.py")], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "WS_OK" in r.stdout


def test_subset_iteration(engine, oracle):
    p = 2
    mesh = ExtrudedHexMesh(5, 5, 4, warp=0.02)
    V, cells, m0, m1, x, y, X = build(mesh, p)
    idx = np.array([0, 3, 4, 11, 17, 24], dtype=np.int32)
    sub = op2.Subset(cells, idx)
    k = op2.Kernel("helmholtz", degree=p)
    op2.par_loop(k, sub, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    yo = np.zeros(V.node_count)
    for c in idx:
        yo += oracle_action(oracle, mesh, V, p, x.data_ro, start=int(c), end=int(c) + 1)
    assert relerr(y.data_ro, yo) < TOL


def test_patch_test_and_nullspace(engine):
    """Size-independent properties at a larger size: constants are in the null
    space of the Poisson operator; 1^T M 1 = |Omega|; linear fields give zero
    interior residual on a warped (non-affine) mesh."""
    p = 3
    mesh = ExtrudedHexMesh(12, 10, 16, warp=0.05, permute_seed=1)
    V, cells, m0, m1, x, y, X = build(mesh, p)
    x.data[:] = 1.0
    op2.par_loop(op2.Kernel("helmholtz", degree=p), cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    assert np.abs(y.data_ro).max() < 1e-11
    y.zero()
    op2.par_loop(op2.Kernel("helmholtz", degree=p, alpha=0.0, beta=1.0), cells,
                 y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    assert abs(y.data_ro.sum() - 1.0) < 1e-12


def test_errors_are_loud(engine):
    mesh = ExtrudedHexMesh(2, 2, 2)
    V, cells, m0, m1, x, y, X = build(mesh, 1)
    from firedrake_b200 import EngineError
    with pytest.raises(EngineError):
        op2.par_loop(op2.Kernel("helmholtz", degree=7), cells, y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    with pytest.raises(ValueError):
        op2.par_loop(op2.Kernel("helmholtz", degree=1), cells, y(op2.READ, m0), X(op2.READ, m1), x(op2.READ, m0))


@pytest.mark.parametrize("permute", [None, 5])
def test_pipelined_host_call(engine, oracle, permute):
    """Host-pointer call large enough to take the chunked H2D | kernel | D2H
    pipeline (>= 64*16 columns): same result as the oracle, for the
    cell-closure numbering (real overlap) and for a random base-cell order
    (schedule degenerates to upload-all / download-all)."""
    p = 1
    mesh = ExtrudedHexMesh(33, 32, 3, warp=0.05, permute_seed=permute)
    V, cells, m0, m1, x, y, X = build(mesh, p)
    gk = op2.GlobalKernel(op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=0.3), [m0, m1], extruded=True)
    loop = op2.Parloop(gk, cells, [y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0)], location="host")
    yo = oracle_action(oracle, mesh, V, p, x.data_ro, alpha=1.0, beta=0.3)
    for rep in range(2):
        y.zero()
        loop()
        assert relerr(y.data_ro, yo) < TOL
        x.data[:] *= 1.0            # version bump -> x is uploaded again (pipelined)


@pytest.mark.parametrize("p", [1, 3])
def test_native_hex_mesh_maps(engine, oracle, p):
    """Non-extruded hexes (``BoxMesh(..., hexahedral=True)``-style maps, reference
    firedrake/utility_meshes.py:1617-1636): one map row per cell, no offsets.
    Built here by expanding the extruded map; must equal the extruded result."""
    mesh = ExtrudedHexMesh(4, 3, 5, warp=0.05, permute_seed=4)
    V = mesh.function_space(p)
    full = V.full_cell_node_list()
    cfull = mesh.coord_space.full_cell_node_list()
    rng = np.random.default_rng(0)
    perm = rng.permutation(full.shape[0])             # cells in arbitrary order
    cells = op2.Set(full.shape[0])
    nodes = op2.Set(V.node_count)
    vnodes = op2.Set(mesh.coord_space.node_count)
    m0 = op2.Map(cells, nodes, V.arity, full[perm])
    m1 = op2.Map(cells, vnodes, 8, cfull[perm])
    x = op2.Dat(nodes, rng.standard_normal(V.node_count))
    y = op2.Dat(nodes)
    X = op2.Dat(op2.DataSet(vnodes, 3), mesh.coordinates)
    op2.par_loop(op2.Kernel("helmholtz", degree=p, alpha=1.0, beta=0.7), cells,
                 y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    yo = oracle_action(oracle, mesh, V, p, x.data_ro, alpha=1.0, beta=0.7)
    assert relerr(y.data_ro, yo) < TOL


def test_edge_cases(engine, oracle):
    """Empty ranges, a single layer, a single column, ranges not starting at 0."""
    p = 2
    mesh = ExtrudedHexMesh(3, 2, 1, warp=0.03)                     # one layer
    V, cells, m0, m1, x, y, X = build(mesh, p)
    k = op2.Kernel("helmholtz", degree=p)
    gk = op2.GlobalKernel(k, [m0, m1], extruded=True)
    loop = op2.Parloop(gk, cells, [y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0)])
    loop._compute((2, 2))                                          # empty: no launch, no error
    assert np.abs(y.data_ro).max() == 0.0
    loop._compute((1, 4))                                          # start > 0
    yo = oracle_action(oracle, mesh, V, p, x.data_ro, start=1, end=4)
    assert relerr(y.data_ro, yo) < TOL
    mesh1 = ExtrudedHexMesh(1, 1, 9, warp=0.0)                     # a single column
    V1, cells1, a0, a1, x1, y1, X1 = build(mesh1, 3)
    op2.par_loop(op2.Kernel("helmholtz", degree=3, beta=1.0), cells1, y1(op2.INC, a0), X1(op2.READ, a1),
                 x1(op2.READ, a0))
    assert relerr(y1.data_ro, oracle_action(oracle, mesh1, V1, 3, x1.data_ro, beta=1.0)) < TOL


@pytest.mark.parametrize("affine", [False, True])
@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (1.0, 1.0), (0.0, 2.0)])
@pytest.mark.parametrize("nz", [16, 37, 70])
def test_thread_per_cell_kernels(engine, oracle, p, alpha, beta, nz, affine):
    """Degrees 1 and 2 on columns of >= 16 layers run the one-thread-per-cell kernels (q1_action.cu,
    q2_action.cu): full and partial warps of layers, several warp-items per column, permuted base
    cells, [start, end) ranges and INC semantics."""
    # affine: a box mesh of parallelepipeds and the per-cell-metric variant (desc.affine_cells)
    mesh = (ExtrudedHexMesh(4, 3, nz, Lx=2.0, Ly=0.75, Lz=1.5, permute_seed=3) if affine
            else ExtrudedHexMesh(4, 3, nz, warp=0.05, permute_seed=3))
    V, cells, m0, m1, x, y, X = build(mesh, p)
    k = op2.Kernel("helmholtz", degree=p, alpha=alpha, beta=beta, affine=affine)
    gk = op2.GlobalKernel(k, [m0, m1], extruded=True)
    y.data[:] = 1.0
    for part in ((0, 5), (5, mesh.num_base_cells)):
        op2.Parloop(gk, cells, [y(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0)])._compute(part)
    yo = oracle_action(oracle, mesh, V, p, x.data_ro, alpha=alpha, beta=beta)
    assert relerr(y.data_ro - 1.0, yo) < TOL
    # a subset of the columns
    sub = cells(1, 4, 7, 10)
    y2 = op2.Dat(y.dataset)
    op2.par_loop(k, sub, y2(op2.INC, m0), X(op2.READ, m1), x(op2.READ, m0))
    yo2 = np.zeros_like(yo)
    from firedrake_b200.fiat_lite import interval_element
    for c in (1, 4, 7, 10):
        oracle.action_extruded(interval_element(p), c, c + 1, [0, mesh.layers], yo2, mesh.coordinates,
                               np.ascontiguousarray(x.data_ro), V.cell_node_map, V.offset, mesh.coord_map,
                               mesh.coord_offset, cdim=1, alpha=alpha, beta=beta)
    assert relerr(y2.data_ro, yo2) < TOL
