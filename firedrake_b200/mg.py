"""Geometric multigrid on extruded hex hierarchies:
``prolong`` / ``restrict`` / ``inject`` and a matrix-free V-cycle.

Reference: firedrake/mg/interface.py:37-113 (prolong), :116-190 (restrict),
:193-280 (inject), kernels firedrake/mg/kernels.py:157-380.  The reference loops
over FINE NODES; each one locates itself in a coarse cell by a Newton iteration
on the coarse coordinate field and evaluates the coarse basis there.  For the
nested hierarchies produced by uniform refinement that search always lands on
the parent cell at a reference position known a priori, so here the loops run
over COARSE CELLS and apply 1-D transfer matrices by sum factorisation:

    prolong   fine[(2p+1)^3 lattice of the 8 children] = (P (x) P (x) P) coarse      WRITE
    restrict  coarse += (P (x) P (x) P)^T (fine / multiplicity)                      INC
    inject    coarse[a] = fine function evaluated at coarse node a                   WRITE

``P[i][a]`` = coarse 1-D basis function a at fine lattice position i.  The transfer
kernels are generated C run through the engine's generic wrapper builder
(codegen.py / csrc/wrapper_jit.cu), with a coarse-cell -> fine-node map whose
layer offset is twice the fine space's (one coarse layer = two fine layers).

Status: kernels and maps are CPU-verified (tests/test_mg.py: polynomial
exactness, restrict == prolong^T, inject o prolong == id); the V-cycle logic is
CPU-verified against a mock engine (tests/test_host_logic_mock.py: level-independent
contraction ~0.12 for CG1, ~0.19 for CG2, PCG in 6-8 iterations on 4^3..32^3) and runs on
the GPU since round 2 (tests/test_jit_gpu.py::test_mg_*, benchmarks/mg_solve.py: CG3 128^3,
4 levels, 11 PCG iterations, profiles/r02_mg_solve_128.json).
"""
from __future__ import annotations

import numpy as np

from . import op2
from .codegen import CStringKernel
from .fiat_lite import _lagrange_tab, interval_element


# ------------------------------------------------------------------ 1-D tables
def fine_lattice_positions(p):
    """Positions on the COARSE reference interval of the 2p+1 fine nodes of its two
    children, ascending (child c holds GLL node m at (c + gll[m]) / 2)."""
    gll = np.sort(interval_element(p).nodes)
    return np.concatenate([gll / 2.0, (1.0 + gll[1:]) / 2.0])


def prolongation_matrix(p):
    """(2p+1, p+1): coarse basis a (dof numbering) at fine lattice position i."""
    B, _ = _lagrange_tab(interval_element(p).nodes, fine_lattice_positions(p))
    return B


def injection_matrix(p):
    """(p+1, 2p+1): value at coarse node a (dof numbering) of the fine piecewise
    polynomial with lattice coefficients; a node on the children's interface is
    evaluated in child 0."""
    el = interval_element(p)
    gll = np.sort(el.nodes)
    J = np.zeros((p + 1, 2 * p + 1))
    for a, xc in enumerate(el.nodes):
        c = 0 if xc <= 0.5 else 1
        xi = 2.0 * xc - c                                  # position inside child c
        B, _ = _lagrange_tab(gll, np.array([xi]))          # child basis, ascending positions
        J[a, c * p:c * p + p + 1] = B[0]
    J[np.abs(J) < 1e-15] = 0.0
    return J


def _table(a):
    a = np.asarray(a)
    if a.ndim == 1:
        return "{" + ", ".join(repr(float(v)) for v in a) + "}"
    return "{" + ", ".join(_table(r) for r in a) + "}"


_TENSOR = """
/* out[(i*NO+j)*NO+k][c] = sum_{a,b,d} T[i][a] T[j][b] T[k][d] in[(a*NI+b)*NI+d][c], by sum factorisation */
static inline void %(name)s_apply(const double T[%(NO)d][%(NI)d], const double *in, double *out)
{
    double t1[%(NO)d * %(NI)d * %(NI)d * %(CD)d], t2[%(NO)d * %(NO)d * %(NI)d * %(CD)d];
    for (int i = 0; i < %(NO)d; ++i)
        for (int r = 0; r < %(NI)d * %(NI)d * %(CD)d; ++r) {
            double s = 0.0;
            for (int a = 0; a < %(NI)d; ++a) s += T[i][a] * in[a * %(NI)d * %(NI)d * %(CD)d + r];
            t1[i * %(NI)d * %(NI)d * %(CD)d + r] = s;
        }
    for (int i = 0; i < %(NO)d; ++i)
        for (int j = 0; j < %(NO)d; ++j)
            for (int r = 0; r < %(NI)d * %(CD)d; ++r) {
                double s = 0.0;
                for (int b = 0; b < %(NI)d; ++b) s += T[j][b] * t1[(i * %(NI)d + b) * %(NI)d * %(CD)d + r];
                t2[(i * %(NO)d + j) * %(NI)d * %(CD)d + r] = s;
            }
    for (int ij = 0; ij < %(NO)d * %(NO)d; ++ij)
        for (int k = 0; k < %(NO)d; ++k)
            for (int c = 0; c < %(CD)d; ++c) {
                double s = 0.0;
                for (int d = 0; d < %(NI)d; ++d) s += T[k][d] * t2[(ij * %(NI)d + d) * %(CD)d + c];
                out[(ij * %(NO)d + k) * %(CD)d + c] = s;
            }
}
"""


def prolong_kernel(p, cdim=1):
    n, m = p + 1, 2 * p + 1
    code = (f"static const double PRO[{m}][{n}] = {_table(prolongation_matrix(p))};\n"
            + _TENSOR % dict(name="prolong", NO=m, NI=n, CD=cdim)
            + f"""
static void prolong(double *fine, const double *coarse)
{{
    prolong_apply(PRO, coarse, fine);
}}
""")
    return CStringKernel(code, "prolong")


def restrict_kernel(p, cdim=1):
    """coarse (INC) += P^T (fine * weight): ``weight`` = 1 / (number of coarse cells whose
    lattice contains the fine node), so that every fine node contributes exactly once --
    the reference attributes each fine node to one coarse cell instead
    (firedrake/mg/utils.py fine_node_to_coarse_node_map)."""
    n, m = p + 1, 2 * p + 1
    code = (f"static const double PROT[{n}][{m}] = {_table(prolongation_matrix(p).T)};\n"
            + _TENSOR % dict(name="restrict", NO=n, NI=m, CD=cdim)
            + f"""
static void restrict_(double *coarse, const double *fine, const double *weight)
{{
    double w[{m ** 3 * cdim}], r[{n ** 3 * cdim}];
    for (int i = 0; i < {m ** 3}; ++i)
        for (int c = 0; c < {cdim}; ++c) w[i * {cdim} + c] = fine[i * {cdim} + c] * weight[i];
    restrict_apply(PROT, w, r);
    for (int i = 0; i < {n ** 3 * cdim}; ++i) coarse[i] += r[i];
}}
""")
    return CStringKernel(code, "restrict_")


def inject_kernel(p, cdim=1):
    n, m = p + 1, 2 * p + 1
    code = (f"static const double INJ[{n}][{m}] = {_table(injection_matrix(p))};\n"
            + _TENSOR % dict(name="inject", NO=n, NI=m, CD=cdim)
            + f"""
static void inject(double *coarse, const double *fine)
{{
    inject_apply(INJ, fine, coarse);
}}
""")
    return CStringKernel(code, "inject")


# ------------------------------------------------------------------- hierarchy
def coarse_to_fine_node_map(Vc, Vf):
    """(coarse columns, (2p+1)^3) fine node of every lattice point of the BOTTOM coarse
    cell of each column, lattice order (i*(2p+1)+j)*(2p+1)+k, and the per-coarse-layer
    offsets.  ``Vc``/``Vf``: ExtrudedFunctionSpaces of the same degree on a mesh and
    its uniform refinement (2x in x, y and layers)."""
    mc, mf = Vc.mesh, Vf.mesh
    p, n = Vc.degree, Vc.degree + 1
    if (mf.nx, mf.ny, mf.nz) != (2 * mc.nx, 2 * mc.ny, 2 * mc.nz) or Vf.degree != p:
        raise ValueError("Vf must be the same space on the uniform refinement of Vc's mesh")
    fcol = np.full((mf.nx, mf.ny), -1, dtype=np.int64)
    fcol[mf.cell_ix, mf.cell_iy] = np.arange(mf.num_base_cells)
    pos2dof = np.empty(n, dtype=np.int64)
    pos2dof[np.array([0, n - 1] + list(range(1, n - 1)))] = np.arange(n)    # ascending position -> dof
    m = 2 * p + 1
    child = np.where(np.arange(m) <= p, 0, 1)                # lattice index -> child, local position
    lpos = np.arange(m) - child * p
    ldof = pos2dof[lpos]
    values = np.empty((mc.num_base_cells, m ** 3), dtype=np.int32)
    offset = np.empty(m ** 3, dtype=np.int32)
    fmap, foff = Vf.cell_node_map.astype(np.int64), np.asarray(Vf.offset, dtype=np.int64)
    for i in range(m):
        for j in range(m):
            cols = fcol[2 * mc.cell_ix + child[i], 2 * mc.cell_iy + child[j]]
            for k in range(m):
                loc = (ldof[i] * n + ldof[j]) * n + ldof[k]
                L = (i * m + j) * m + k
                values[:, L] = fmap[cols, loc] + foff[loc] * child[k]
                offset[L] = 2 * foff[loc]
    return values, offset


class TransferManager:
    """prolong / restrict / inject between two consecutive levels
    (firedrake/mg/embedded.py TransferManager, firedrake/mg/interface.py)."""

    def __init__(self, Vc, Vf):
        """``Vc``, ``Vf``: assemble.FunctionSpace on consecutive levels."""
        self.Vc, self.Vf = Vc, Vf
        p = Vc.degree
        vals, off = coarse_to_fine_node_map(Vc.V, Vf.V)
        self.c2f = op2.Map(Vc.cell_set, Vf.node_set, (2 * p + 1) ** 3, vals, offset=off, name="coarse_to_fine")
        self._k = (prolong_kernel(p, Vc.cdim), restrict_kernel(p, Vc.cdim), inject_kernel(p, Vc.cdim))
        self._weight = None

    @property
    def weight(self):
        """1 / multiplicity of every fine node in the coarse-cell lattices (scalar Dat)."""
        if self._weight is None:
            # scalar field on the fine nodes, sharing the fine space's halo: the counts of nodes on
            # a slab interface are summed into their owner and sent back to the ghost copies
            w = op2.Dat(op2.DataSet(self.Vf.node_set, 1, halo=self.Vf.dof_dset.halo))
            m3 = (2 * self.Vc.degree + 1) ** 3
            count = CStringKernel(f"static void count(double *w) {{ for (int i = 0; i < {m3}; ++i) w[i] += 1.0; }}",
                                  "count")
            op2.par_loop(count, self.Vc.cell_set, w(op2.INC, self.c2f))
            inv = CStringKernel("static void recip(double *w) { *w = 1.0 / *w; }", "recip")
            op2.par_loop(inv, self.Vf.node_set, w(op2.RW))
            self._weight = w
        return self._weight

    def prolong(self, coarse: op2.Dat, fine: op2.Dat):
        op2.par_loop(self._k[0], self.Vc.cell_set, fine(op2.WRITE, self.c2f),
                     coarse(op2.READ, self.Vc.cell_node_map))
        return fine

    def restrict(self, fine_dual: op2.Dat, coarse_dual: op2.Dat):
        coarse_dual.zero()
        coarse_dual.device_ptr                      # materialise the zero on the device
        op2.par_loop(self._k[1], self.Vc.cell_set, coarse_dual(op2.INC, self.Vc.cell_node_map),
                     fine_dual(op2.READ, self.c2f), self.weight(op2.READ, self.c2f))
        return coarse_dual

    def inject(self, fine: op2.Dat, coarse: op2.Dat):
        op2.par_loop(self._k[2], self.Vc.cell_set, coarse(op2.WRITE, self.Vc.cell_node_map),
                     fine(op2.READ, self.c2f))
        return coarse


def prolong(coarse, fine, manager):
    return manager.prolong(coarse, fine)


def restrict(fine_dual, coarse_dual, manager):
    return manager.restrict(fine_dual, coarse_dual)


def inject(fine, coarse, manager):
    return manager.inject(fine, coarse)


def _touched(*dats):
    """Vector algebra ran over the local entries: owned rows are right, ghost rows are not
    (pyop2/types/dat.py:622-678: any write invalidates the halo)."""
    for d in dats:
        d._device_written()
        d.halo_valid = False


# --------------------------------------------------------------------- V-cycle
class MeshHierarchy:
    """``ExtrudedMeshHierarchy`` of uniformly refined extruded hex meshes
    (firedrake/mg/mesh.py:190-260): level l has 2^l times the coarse resolution in
    every direction."""

    def __init__(self, nx, ny, nz, levels, rank=0, nranks=1, **mesh_kwargs):
        """``rank`` / ``nranks``: this process's slab of every level (firedrake_b200.partition); the
        coarse slab bounds must refine exactly (``nx`` divisible by ``nranks``) so that every
        coarse cell's children live on the same rank, as in a refined DMPlex distribution."""
        from .utility_meshes import ExtrudedHexMesh
        self.partitions = None
        if nranks > 1:
            if nx % nranks:
                raise ValueError("coarse nx must be divisible by the number of ranks")
            self.nranks, self.rank = nranks, rank
            self._sizes = [(nx << l, ny << l, nz << l) for l in range(levels + 1)]
            self._kwargs = mesh_kwargs
            self.partitions = {}
            self.meshes = [None] * (levels + 1)
            return
        self.meshes = [ExtrudedHexMesh(nx << l, ny << l, nz << l, **mesh_kwargs) for l in range(levels + 1)]

    def partition(self, level, degree):
        """SlabPartition of ``level`` for function spaces of ``degree`` (built on demand)."""
        from .partition import SlabPartition
        key = (level, degree)
        if key not in self.partitions:
            nx, ny, nz = self._sizes[level]
            self.partitions[key] = SlabPartition(nx, ny, nz, degree, self.rank, self.nranks, **self._kwargs)
            self.meshes[level] = self.partitions[key].mesh
        return self.partitions[key]

    def __len__(self):
        return len(self.meshes)

    def __getitem__(self, i):
        return self.meshes[i]


class VCycle:
    """Matrix-free geometric multigrid V-cycle for the Helmholtz family with rediscretised
    coarse operators (the ``pc_type mg`` + ``mat_type matfree`` setup of
    demos/multigrid/geometric_multigrid.py.rst): damped-Jacobi smoothing with the
    assembled diagonal (``ImplicitMatrixContext.getDiagonal``, operators.py:199-205),
    coarsest level solved by CG.  Everything stays on the device."""

    def __init__(self, hierarchy, degree, make_form, bc_domains=(), nu=2, omega=0.8,
                 coarse_rtol=1e-2, coarse_maxit=200, allreduce=None):
        """``allreduce``: callable summing a float over the ranks (partitioned hierarchies)."""
        from .assemble import DirichletBC, FunctionSpace, assemble
        self.allreduce = allreduce
        if hierarchy.partitions is None:
            self.spaces = [FunctionSpace(m, degree) for m in hierarchy.meshes]
        else:
            parts = [hierarchy.partition(l, degree) for l in range(len(hierarchy))]
            self.spaces = [FunctionSpace(pt.mesh, degree, partition=pt) for pt in parts]
        self.bcs = [[DirichletBC(V, 0.0, s) for s in bc_domains] for V in self.spaces]
        self.ops = [assemble(make_form(V), bcs=b, mat_type="matfree") for V, b in zip(self.spaces, self.bcs)]
        self.transfers = [TransferManager(self.spaces[l], self.spaces[l + 1]) for l in range(len(self.spaces) - 1)]
        self.nu, self.omega = nu, omega
        self.coarse_rtol, self.coarse_maxit = coarse_rtol, coarse_maxit
        self.invdiag = []
        for V, A in zip(self.spaces, self.ops):
            d = A.getDiagonal(V.dat())
            op2.par_loop(CStringKernel("static void recip(double *w) { *w = 1.0 / *w; }", "recip"),
                         V.node_set, d(op2.RW))
            self.invdiag.append(d)
        # per level: r residual, t smoother scratch, e prolonged correction; as a COARSE level also
        # b (restricted residual = its right-hand side) and x (its solution) -- x and e must be
        # distinct Dats: level l-1's solution is the input of the prolongation into level l's e
        self._work = [dict(r=V.dat(), e=V.dat(), t=V.dat(), b=V.dat(), x=V.dat()) for V in self.spaces]

    def _smooth(self, l, b, x):
        """x += omega D^-1 (b - A x), nu times."""
        from . import _lib
        L = _lib.lib()
        A, w = self.ops[l], self._work[l]
        n = b._data.size
        for _ in range(self.nu):
            A.mult(x, w["t"])
            _lib.check(L.fdb_vec_aypx(n, -1.0, b.device_ptr, w["t"].device_ptr))        # t = b - A x
            _lib.check(L.fdb_vec_pointwise_mult(n, w["t"].device_ptr, self.invdiag[l].device_ptr,
                                                w["t"].device_ptr))
            _touched(w["t"])
            _lib.check(L.fdb_vec_axpy(n, self.omega, w["t"].device_ptr, x.device_ptr))
            _touched(x)

    def apply(self, l, b, x):
        """One V-cycle on level l for A x = b, starting from the x passed in."""
        from . import _lib
        from .assemble import cg
        L = _lib.lib()
        A, w = self.ops[l], self._work[l]
        n = b._data.size
        if l == 0:
            cg(A, b, x, rtol=self.coarse_rtol, maxit=self.coarse_maxit, allreduce=self.allreduce)
            _touched(x)
            return x
        self._smooth(l, b, x)
        A.mult(x, w["r"])
        _lib.check(L.fdb_vec_aypx(n, -1.0, b.device_ptr, w["r"].device_ptr))             # r = b - A x
        _touched(w["r"])
        for bc in self.bcs[l]:
            bc.zero(w["r"])
        wc = self._work[l - 1]
        T = self.transfers[l - 1]
        T.restrict(w["r"], wc["b"])
        for bc in self.bcs[l - 1]:
            bc.zero(wc["b"])
        wc["x"].zero()
        wc["x"].device_ptr
        self.apply(l - 1, wc["b"], wc["x"])
        T.prolong(wc["x"], w["e"])
        for bc in self.bcs[l]:
            bc.zero(w["e"])
        _lib.check(L.fdb_vec_axpy(n, 1.0, w["e"].device_ptr, x.device_ptr))
        _touched(x)
        self._smooth(l, b, x)
        return x


def pcg(A, b, x, M, rtol=1e-8, maxit=200, allreduce=None):
    """Preconditioned CG (``ksp_type cg`` with a multigrid ``pc``): ``M(r, z)`` applies the
    preconditioner (e.g. ``lambda r, z: vcycle.apply(top, r, z)`` from z = 0).  ``allreduce``:
    callable summing a float over the ranks; inner products then run over the OWNED dofs."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    V = b.dataset
    r, z, p, Ap = (op2.Dat(V) for _ in range(4))
    n = b._data.size
    n_owned = b.dataset.set.size * b.cdim

    def dot(u, v):
        out = C.c_double()
        _lib.check(L.fdb_vec_dot(n_owned, u.device_ptr, v.device_ptr, C.byref(out)))
        return allreduce(out.value) if allreduce else out.value
    A.mult(x, Ap)
    _lib.check(L.fdb_memcpy_d2d(r.device_ptr, b.device_ptr, b.nbytes))
    r._device_written()
    _lib.check(L.fdb_vec_axpy(n, -1.0, Ap.device_ptr, r.device_ptr))
    _touched(r)
    r0 = np.sqrt(dot(r, r))
    hist = [r0]
    z.zero(); z.device_ptr
    M(r, z)
    _lib.check(L.fdb_memcpy_d2d(p.device_ptr, z.device_ptr, z.nbytes))
    p._device_written()
    rz = dot(r, z)
    it = 0
    while it < maxit and hist[-1] > rtol * r0:
        A.mult(p, Ap)
        alpha = rz / dot(p, Ap)
        _lib.check(L.fdb_vec_axpy(n, alpha, p.device_ptr, x.device_ptr))
        _lib.check(L.fdb_vec_axpy(n, -alpha, Ap.device_ptr, r.device_ptr))
        _touched(x, r)
        hist.append(np.sqrt(dot(r, r)))
        z.zero(); z.device_ptr
        M(r, z)
        rz_new = dot(r, z)
        _lib.check(L.fdb_vec_aypx(n, rz_new / rz, z.device_ptr, p.device_ptr))            # p = z + beta p
        _touched(p)
        rz = rz_new
        it += 1
    return it, hist
