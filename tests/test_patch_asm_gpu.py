"""Batched patch solves on the device against NumPy: the
additive Schwarz operator of TinyASM's BlockJacobi (tinyasm/tinyasm.cpp:27-120) on vertex-star
patches of assembled Helmholtz matrices, and its use as a smoother/preconditioner."""
import numpy as np
import pytest
import scipy.sparse as sp

from firedrake_b200.assemble import DirichletBC, FunctionSpace, assemble, helmholtz, poisson
from firedrake_b200.patch import PatchASM, vertex_star_patches
from firedrake_b200.utility_meshes import ExtrudedHexMesh

pytestmark = pytest.mark.gpu


def dense_asm(A, ptr, dofs, b):
    x = np.zeros_like(b)
    for k in range(len(ptr) - 1):
        d = dofs[ptr[k]:ptr[k + 1]]
        x[d] += np.linalg.solve(A[np.ix_(d, d)], b[d])
    return x


@pytest.mark.parametrize("p,shape", [(1, (4, 3, 5)), (2, (3, 3, 3)), (3, (2, 2, 3))])
def test_patch_inverses_and_additive_apply(engine, p, shape):
    mesh = ExtrudedHexMesh(*shape, warp=0.05, permute_seed=1)
    V = FunctionSpace(mesh, p)
    A = assemble(helmholtz(V))
    ro, co, va = A.csr()
    Ad = sp.csr_matrix((va, co, ro), shape=(V.node_count,) * 2).toarray()
    ptr, dofs = vertex_star_patches(V)
    sizes = np.diff(ptr)
    assert sizes.max() == (2 * p - 1) ** 3 and len(sizes) == np.prod(np.array(shape) + 1)
    pc = PatchASM(A, ptr, dofs)
    for k, inv in enumerate(pc.inverse_blocks()):
        d = dofs[ptr[k]:ptr[k + 1]]
        ref = np.linalg.inv(Ad[np.ix_(d, d)])
        assert np.abs(inv - ref).max() < 1e-10 * np.abs(ref).max()
    b = V.dat(np.random.default_rng(4).standard_normal(V.node_count))
    x = V.dat()
    pc.apply(b, x)
    ref = dense_asm(Ad, ptr, dofs, b.data_ro)
    assert np.abs(x.data_ro - ref).max() < 1e-10 * np.abs(ref).max()


def test_large_patches_take_the_global_memory_path(engine):
    """Patches too large for shared memory (here: one patch = the whole 343-dof space in a
    shuffled order) are inverted in place in global memory."""
    mesh = ExtrudedHexMesh(2, 2, 2, warp=0.03)
    V = FunctionSpace(mesh, 3)            # 343 dofs > 158 = largest block that fits 200 KB
    A = assemble(helmholtz(V))
    ro, co, va = A.csr()
    Ad = sp.csr_matrix((va, co, ro), shape=(V.node_count,) * 2).toarray()
    ptr = np.array([0, V.node_count], dtype=np.int64)
    dofs = np.random.default_rng(0).permutation(V.node_count).astype(np.int32)
    pc = PatchASM(A, ptr, dofs)
    inv = pc.inverse_blocks()[0]
    ref = np.linalg.inv(Ad[np.ix_(dofs, dofs)])
    assert np.abs(inv - ref).max() < 1e-9 * np.abs(ref).max()


def test_star_smoother_preconditions_cg(engine):
    """Vertex-star ASM as the preconditioner of CG on Poisson CG2 with Dirichlet rows: far fewer
    iterations than unpreconditioned CG and the same solution (the role ASMStarPC + TinyASM play
    in the reference: firedrake/preconditioners/asm.py)."""
    from firedrake_b200 import mg
    from firedrake_b200.assemble import cg
    mesh = ExtrudedHexMesh(6, 6, 6, warp=0.05)
    V = FunctionSpace(mesh, 2)
    bcs = [DirichletBC(V, 0.0, ["bottom", "top"])]
    A = assemble(poisson(V), bcs=bcs)
    ptr, dofs = vertex_star_patches(V, exclude=bcs[0].nodes)
    pc = PatchASM(A, ptr, dofs)
    b = V.dat(np.random.default_rng(2).standard_normal(V.node_count))
    bcs[0].zero(b)
    x0, x1 = V.dat(), V.dat()
    x0.device_ptr, x1.device_ptr
    n_plain, _ = cg(A, b, x0, rtol=1e-9, maxit=2000)
    n_pc, _ = mg.pcg(A, b, x1, lambda r, z: pc.apply(r, z), rtol=1e-9, maxit=2000)
    assert n_pc < 0.7 * n_plain, (n_pc, n_plain)     # unweighted additive Schwarz: ~2x fewer iterations here
    assert np.abs(x0.data_ro - x1.data_ro).max() < 1e-6 * np.abs(x0.data_ro).max()
