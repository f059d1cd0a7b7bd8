"""GPU test of the monolithic mixed matrices (op2.Mat over MixedDataSets, MatBlock views).

Kept in a module that sorts LAST: this code was written after the round's GPU budget was spent, so on
the device it is still unproven (non-strict xfail); on the CPU the same function runs against the mock
engine (tests/test_host_logic_mock.py::test_mixed_mat_host_logic)."""
import numpy as np
import pytest

from firedrake_b200 import op2

pytestmark = [pytest.mark.gpu]


def _mixed_block_kernel(i, j, nr, nc):
    """Element tensor of block (i, j): A[r][c] += (i + 1) * (r + 1) + 0.5 * (j + 1) * (c + 2)."""
    return op2.Kernel(f"static void blk{i}{j}(double *A) {{ for (int r = 0; r < {nr}; ++r) for (int c = 0; c < {nc}; ++c) "
                      f"A[r * {nc} + c] += {i + 1}.0 * (r + 1) + {0.5 * (j + 1)} * (c + 2); }}", f"blk{i}{j}")


@pytest.mark.xfail(strict=False, reason="monolithic mixed matrices were written after this round's GPU budget was spent: "
                                        "validated on the CPU mock (tests/test_host_logic_mock.py); the device paths "
                                        "they use (scalar Mat, generic wrapper with a Mat argument) are GPU-validated")
@pytest.mark.parametrize("extruded", [False, True])
def test_mixed_mat_monolithic(engine, extruded):
    """op2.Mat over MixedDataSets (pyop2/types/mat.py:607-700): monolithic scalar CSR, ``mat[i, j]``
    MatBlock arguments (vector-valued velocity block x scalar pressure block), block lgmaps, the whole
    mixed element tensor in one argument, diagonal entries of a block, mult on MixedDats."""
    rng = np.random.default_rng(11)
    ncol, nlay = 40, (3 if extruded else 1)
    # node numbering: column c holds its nodes contiguously per layer interface when extruded
    nv_col, np_col = 5, 3
    if extruded:
        cells = op2.ExtrudedSet(op2.Set(ncol), nlay + 1)
        nv, npr = ncol * nv_col * (nlay + 1), ncol * np_col * (nlay + 1)
        offv, offp = [nv_col] * 3, [np_col] * 2
    else:
        cells = op2.Set(ncol)
        nv, npr = ncol * nv_col, ncol * np_col
        offv = offp = None
    vset, pset = op2.Set(nv), op2.Set(npr)
    stride_v, stride_p = (nv_col * (nlay + 1), np_col * (nlay + 1)) if extruded else (nv_col, np_col)
    # each column couples its own nodes with some of the next column's (shared dofs across cells)
    mv_vals = np.array([[c * stride_v + 0, c * stride_v + 3, ((c + 1) % ncol) * stride_v + 1] for c in range(ncol)])
    mp_vals = np.array([[c * stride_p + 2, ((c + 7) % ncol) * stride_p + 0] for c in range(ncol)])
    mv = op2.Map(cells, vset, 3, mv_vals, offset=offv)
    mp = op2.Map(cells, pset, 2, mp_vals, offset=offp)
    W = op2.MixedDataSet([vset ** 2, pset])
    mm = op2.MixedMap([mv, mp])
    A = op2.Mat(op2.Sparsity((W, W), [(mm, mm, None)]))
    n0, n1 = nv * 2, npr
    assert A.sparsity.shape == (n0 + n1, n0 + n1) and A[0, 1] is A[0, 1]
    maps, sizes, offs = (mv, mp), (6, 2), (0, n0)
    cdims = (2, 1)

    def dofs(b, c, layer):
        m = maps[b]
        nodes = m.values[c] + (0 if m.offset is None else m.offset * layer)
        return (nodes[:, None] * cdims[b] + np.arange(cdims[b])[None, :]).ravel() + offs[b]

    def local(i, j):
        r, c = np.arange(sizes[i])[:, None], np.arange(sizes[j])[None, :]
        return (i + 1.0) * (r + 1) + 0.5 * (j + 1) * (c + 2)

    def expected(blocks, row_mask=None, col_mask=None):
        E = np.zeros((n0 + n1, n0 + n1))
        for c in range(ncol):
            for layer in range(nlay):
                for (i, j) in blocks:
                    rr, cc = dofs(i, c, layer), dofs(j, c, layer)
                    T = local(i, j).copy()
                    if row_mask is not None and i == row_mask[0]:
                        T[np.isin(rr, row_mask[1])] = 0
                    if col_mask is not None and j == col_mask[0]:
                        T[:, np.isin(cc, col_mask[1])] = 0
                    np.add.at(E, (rr[:, None], cc[None, :]), T)
        return E

    all_blocks = [(0, 0), (0, 1), (1, 0), (1, 1)]
    A.zero()
    for (i, j) in all_blocks:
        op2.par_loop(_mixed_block_kernel(i, j, sizes[i], sizes[j]), cells, A[i, j](op2.INC, (maps[i], maps[j])))
    A.assemble()
    E = expected(all_blocks)
    assert np.abs(A.values - E).max() < 1e-12
    assert np.abs(A[1, 0].values - E[n0:, :n0]).max() < 1e-12 and A[0, 1].values.shape == (n0, n1)
    # mult on MixedDats = dense product on the concatenated vector
    x = op2.MixedDat([op2.Dat(vset ** 2, rng.standard_normal((nv, 2))), op2.Dat(pset, rng.standard_normal(npr))])
    y = op2.MixedDat(W)
    A.mult(x, y)
    ye = E @ np.concatenate([x[0].data_ro.ravel(), x[1].data_ro.ravel()])
    assert np.abs(np.concatenate([y[0].data_ro.ravel(), y[1].data_ro.ravel()]) - ye).max() < 1e-11
    # block lgmaps (dof level, block-local): drop component 1 of velocity node 3 from block (0, 1)'s rows
    A.zero()
    lg_r = np.arange(n0, dtype=np.int32)
    lg_r[3 * 2 + 1] = -1
    lg_c = np.arange(n1, dtype=np.int32)
    lg_c[2] = -1
    op2.par_loop(_mixed_block_kernel(0, 1, 6, 2), cells, A[0, 1](op2.INC, (mv, mp), lgmaps=(lg_r, lg_c)))
    E01 = expected([(0, 1)], row_mask=(0, [3 * 2 + 1]), col_mask=(1, [n0 + 2]))
    assert np.abs(A.values - E01).max() < 1e-12
    # diagonal entries of a diagonal block: node rows, one component
    A[0, 0].set_local_diagonal_entries([0, 3], 7.0, idx=1)
    A[1, 1].set_local_diagonal_entries([2], -2.0)
    D = A.values
    assert D[0 * 2 + 1, 0 * 2 + 1] == 7.0 and D[3 * 2 + 1, 3 * 2 + 1] == 7.0 and D[n0 + 2, n0 + 2] == -2.0
    with pytest.raises(ValueError):
        A[0, 1].set_local_diagonal_entries([0])
    # the whole mixed element tensor (8 x 8: velocity dofs then pressure dofs) in one argument
    A.zero()
    kfull = op2.Kernel("static void full(double *A) { for (int r = 0; r < 8; ++r) for (int c = 0; c < 8; ++c) "
                       "A[r * 8 + c] += 10.0 * r + c; }", "full")
    op2.par_loop(kfull, cells, A(op2.INC, (mm, mm)))
    Ef = np.zeros_like(E)
    T = 10.0 * np.arange(8)[:, None] + np.arange(8)[None, :]
    for c in range(ncol):
        for layer in range(nlay):
            d = np.concatenate([dofs(0, c, layer), dofs(1, c, layer)])
            np.add.at(Ef, (d[:, None], d[None, :]), T)
    assert np.abs(A.values - Ef).max() < 1e-12
    with pytest.raises(op2.MapValueError):
        A(op2.INC, (mv, mv))
    with pytest.raises(NotImplementedError):
        op2.Sparsity((W, op2.MixedDataSet([pset, vset ** 2])), [(mm, mm, None)])
