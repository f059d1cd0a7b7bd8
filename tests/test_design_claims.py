"""CPU checks of two structural claims the engine relies on:

1. the within-row CSR position of an element-tensor entry depends on the layer
   only through its class (bottom / interior / top) -- the per-column position
   table of csrc/mat.cu;
2. with the cell-closure dof numbering the rows touched by successive column
   chunks advance monotonically, so the pipelined host call can upload x and
   download y incrementally; with an arbitrary numbering the schedule
   degenerates to upload-all / download-all, never to a wrong one.
Both are restated in NumPy exactly as the C++ code computes them."""
import numpy as np
import pytest

from firedrake_b200.utility_meshes import ExtrudedHexMesh


@pytest.mark.parametrize("p,permute", [(1, None), (2, 3), (3, 0)])
def test_csr_position_is_layer_class_invariant(oracle, p, permute):
    mesh = ExtrudedHexMesh(3, 3, 6, permute_seed=permute)
    V = mesh.function_space(p)
    rowptr, colidx = oracle.build_sparsity(V.node_count, V.cell_node_map, V.offset, mesh.nz)
    nlay, ar = mesh.nz, V.arity
    cmap, off = V.cell_node_map.astype(np.int64), V.offset.astype(np.int64)

    def positions(c, layer):
        rows = cmap[c] + off * layer
        cols = cmap[c] + off * layer
        pos = np.empty((ar, ar), dtype=np.int64)
        for i, r in enumerate(rows):
            seg = colidx[rowptr[r]:rowptr[r + 1]]
            pos[i] = np.searchsorted(seg, cols)
            assert np.array_equal(seg[pos[i]], cols)
        return pos

    for c in range(mesh.num_base_cells):
        table = {0: positions(c, 0), 1: positions(c, 1), 2: positions(c, nlay - 1)}
        for layer in range(nlay):
            cls = 0 if layer == 0 else (2 if layer == nlay - 1 else 1)
            assert np.array_equal(positions(c, layer), table[cls]), (c, layer)


def chunk_plan(cmap, off, nlay, nrows, K):
    """NumPy restatement of the plan in csrc/global_kernel.cu (pipelined_host_action)."""
    ncols = cmap.shape[0]
    lo, hi = [], []
    for c in range(K):
        b0, b1 = ncols * c // K, ncols * (c + 1) // K
        rows = cmap[b0:b1].astype(np.int64)
        lo.append(rows.min())
        hi.append((rows + off[None, :].astype(np.int64) * (nlay - 1)).max() + 1)
    upto = np.maximum.accumulate(hi)
    final_below = np.empty(K, dtype=np.int64)
    mn = nrows
    for c in range(K - 1, -1, -1):
        final_below[c] = mn
        mn = min(mn, lo[c])
    final_below[K - 1] = nrows
    return np.array(lo), np.array(hi), upto, final_below


def test_pipeline_plan_is_safe_and_incremental():
    K = 16
    mesh = ExtrudedHexMesh(32, 32, 4)
    V = mesh.function_space(2)
    full = V.full_cell_node_list().reshape(mesh.num_base_cells, mesh.nz, V.arity)
    lo, hi, upto, final_below = chunk_plan(V.cell_node_map, V.offset, mesh.nz, V.node_count, K)
    ncols = mesh.num_base_cells
    for c in range(K):
        b0, b1 = ncols * c // K, ncols * (c + 1) // K
        touched = full[b0:b1]
        assert touched.max() < upto[c]                       # everything chunk c reads/writes is resident
        later = full[b1:]
        if later.size:
            assert later.min() >= final_below[c]             # nothing below is touched again: safe to download
    assert np.all(np.diff(upto) >= 0) and np.all(np.diff(final_below) >= 0)
    # real overlap for the cell-closure numbering: uploads/downloads advance with the chunks
    frac_up = upto / V.node_count
    frac_dn = final_below / V.node_count
    assert frac_up[K // 2] < 0.65 and frac_dn[K // 2] > 0.35
    # arbitrary numbering relative to the cell order (rows of the map shuffled):
    # degenerate schedule, but still safe
    perm = np.random.default_rng(1).permutation(ncols)
    rmap = np.ascontiguousarray(V.cell_node_map[perm])
    rfull = full[perm]
    rlo, rhi, rupto, rfinal = chunk_plan(rmap, V.offset, mesh.nz, V.node_count, K)
    for c in range(K):
        b0, b1 = ncols * c // K, ncols * (c + 1) // K
        assert rfull[b0:b1].max() < rupto[c]
        if c < K - 1:
            assert rfull[b1:].min() >= rfinal[c]
    assert rupto[0] > 0.9 * V.node_count                     # (almost) everything uploaded up front
    assert rfinal[K - 2] < 0.1 * V.node_count                # (almost) nothing downloadable early
