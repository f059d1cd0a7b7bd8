"""The assembly surface: ``assemble()`` for the supported forms, Dirichlet
conditions, and the matrix-free operator context.

Mirrors (same names / call protocol, thin Python over the engine):

* ``firedrake.assemble.assemble`` for 1-forms (``OneFormAssembler``:
  zero tensor -> parloops -> ``bc.zero``; firedrake/assemble.py:1197-1293) and
  2-forms (``ExplicitMatrixAssembler``: sparsity + Mat allocation, parloop with
  BC-masked lgmaps, unit diagonal on BC rows; :1296-1307, 1377-1409, 1484-1525)
* ``firedrake.bcs.DirichletBC.zero/set/apply`` (firedrake/bcs.py:192-221, 404-457)
* ``firedrake.matrix_free.operators.ImplicitMatrixContext`` (operators.py:74-242)

Forms are described by :class:`Form` (the Helmholtz family on a
:class:`FunctionSpace`) instead of UFL: UFL/TSFC are not available here, and the
engine keys its kernels on a form descriptor.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import op2
from .halo import Halo


class FunctionSpace:
    """Scalar/vector CG_p space on an extruded hex mesh, bundling the PyOP2
    objects ``assemble`` needs (node set, cell set, maps, coordinates), i.e.
    what ``V.cell_node_map()``, ``mesh.coordinates.dat`` and friends give in
    Firedrake (firedrake/functionspaceimpl.py:803-812)."""

    def __init__(self, mesh, degree, cdim=1, partition=None):
        self.mesh, self.degree, self.cdim = mesh, degree, cdim
        self.V = V = mesh.function_space(degree)
        if partition is not None:
            cell_sizes, node_sizes = partition.cell_sizes, partition.node_sizes
            halo = Halo(partition.halo_lists(V), max_cdim=cdim) if partition.nranks > 1 else None
        else:
            cell_sizes, node_sizes, halo = mesh.num_base_cells, V.node_count, None
        self.cell_set = op2.ExtrudedSet(op2.Set(cell_sizes), mesh.layers)
        # exec-halo partition: a property of the WHOLE distributed set (the last rank holds no
        # exec cells but must follow the same protocol)
        self.cell_set.owner_computes = bool(getattr(partition, "exec_halo", False))
        self.node_set = op2.Set(node_sizes)
        self.dof_dset = op2.DataSet(self.node_set, cdim, halo=halo)
        self.vertex_set = op2.Set(mesh.coord_space.node_count)
        self.cell_node_map = op2.Map(self.cell_set, self.node_set, V.arity, V.cell_node_map,
                                     offset=V.offset)
        self.coord_map = op2.Map(self.cell_set, self.vertex_set, 8, mesh.coord_map,
                                 offset=mesh.coord_offset)
        self.coordinates = op2.Dat(op2.DataSet(self.vertex_set, 3), mesh.coordinates)

    def dat(self, data=None, pinned=False):
        return op2.Dat(self.dof_dset, data, pinned=pinned)

    def cells_are_affine(self):
        """True iff every cell is a parallelepiped (checked on the device, cached per version of
        the coordinate Dat): lets the assembler pick the per-cell-metric kernel variant."""
        import ctypes as C
        from . import _lib
        key = self.coordinates.dat_version
        if getattr(self, "_affine", (None, None))[0] != key:
            off = np.ascontiguousarray(self.mesh.coord_offset, dtype=np.int32)
            res = C.c_int()
            _lib.check(_lib.lib().fdb_cells_are_affine(
                self.coordinates.device_ptr, self.coord_map.device_ptr, off.ctypes.data, 0,
                self.cell_set.total_size, self.mesh.nz, C.byref(res)), "fdb_cells_are_affine")
            self._affine = (key, bool(res.value))
        return self._affine[1]

    @property
    def node_count(self):
        return self.V.node_count

    def boundary_nodes(self, sub_domain):
        return self.V.boundary_nodes(sub_domain)


def interpolate_q1(V: "FunctionSpace", source: op2.Dat, target: op2.Dat = None):
    """``Function(V).interpolate(w)`` for a Q1 (x) P1 source ``w`` (scalar or
    vector, e.g. the mesh coordinates -> the physical position of every node of
    V): the dual-evaluation parloop of firedrake/interpolation.py:977-1171 with
    WRITE access on the target.  Returns the target Dat (device resident)."""
    from . import _lib
    from .fiat_lite import interval_element
    cdim = source.cdim
    if target is None:
        target = op2.Dat(op2.DataSet(V.node_set, cdim))
    if not hasattr(V, "_dev_offsets"):
        V._dev_offsets = (op2.DeviceArray.from_host(np.ascontiguousarray(V.V.offset, dtype=np.int32)),
                          op2.DeviceArray.from_host(np.ascontiguousarray(V.mesh.coord_offset, dtype=np.int32)))
    nodes = np.ascontiguousarray(interval_element(V.degree).nodes, dtype=np.float64)
    _lib.check(_lib.lib().fdb_interpolate_q1(
        target.device_ptr, source.device_ptr, V.cell_node_map.device_ptr, V.coord_map.device_ptr,
        V._dev_offsets[0].ptr, V._dev_offsets[1].ptr, V.cell_set.total_size, V.mesh.nz, V.degree + 1,
        cdim, nodes.ctypes.data), "fdb_interpolate_q1")
    target._device_written()
    return target


def interpolation_kernel(degree, expressions, name="interpolate_expr"):
    """C source of the dual-evaluation kernel of ``Function(V).interpolate(expr(x))`` on
    Q_p (x) P_p with GLL nodes (point evaluation: firedrake/interpolation.py:977-1171
    builds it with tsfc.compile_expression_dual_evaluation, tsfc/driver.py:225-386).
    ``expressions``: one C expression per component in ``x[0], x[1], x[2]`` (the physical
    position of the node, the trilinear image of its reference position) -- the syntax of
    the reference's former ``Expression("sin(x[0])")``.  Arguments: out (WRITE), coords."""
    from .fiat_lite import interval_element
    from .codegen import CStringKernel
    exprs = [expressions] if isinstance(expressions, str) else list(expressions)
    n = degree + 1
    xi = ", ".join(repr(float(v)) for v in interval_element(degree).nodes)
    body = "\n".join(f"        out[i * {len(exprs)} + {c}] = {e};" for c, e in enumerate(exprs))
    code = f"""
static void {name}(double *out, const double *X)
{{
    const double xi[{n}] = {{{xi}}};                 /* 1-D node positions, dof numbering */
    for (int ax = 0; ax < {n}; ++ax)
    for (int ay = 0; ay < {n}; ++ay)
    for (int az = 0; az < {n}; ++az) {{
        const int i = (ax * {n} + ay) * {n} + az;
        double x[3] = {{0.0, 0.0, 0.0}};
        for (int v = 0; v < 8; ++v) {{
            const double w = ((v & 4) ? xi[ax] : 1.0 - xi[ax]) * ((v & 2) ? xi[ay] : 1.0 - xi[ay])
                           * ((v & 1) ? xi[az] : 1.0 - xi[az]);
            for (int c = 0; c < 3; ++c) x[c] += w * X[v * 3 + c];
        }}
{body}
    }}
}}
"""
    return CStringKernel(code, name)


def interpolate(V: "FunctionSpace", expressions, target: op2.Dat = None):
    """``Function(V).interpolate(expr)`` for C expressions of the physical coordinates
    (SURVEY.md section 8f row f2), run as a WRITE parloop through the engine's generic
    wrapper builder.  Nodes shared by several cells are written by each of them with the
    same value, as in the reference's sequential loop."""
    exprs = [expressions] if isinstance(expressions, str) else list(expressions)
    if len(exprs) != V.cdim:
        raise ValueError(f"need {V.cdim} expressions for this space, got {len(exprs)}")
    if target is None:
        target = V.dat()
    k = interpolation_kernel(V.degree, exprs)
    op2.par_loop(k, V.cell_set, target(op2.WRITE, V.cell_node_map), V.coordinates(op2.READ, V.coord_map))
    return target


def functional_kernel(degree, measure, facet=None, name=None, integrand="avg"):
    """C source of the 0-form kernels ``f*dx`` and ``f*ds`` on Q_p (x) P_p hexes with
    trilinear geometry (what TSFC emits for a rank-0 form: ``A[0] += w*|J|*f(q)``,
    tsfc/kernel_interface/common.py:139-239; facet kernels get the local facet number
    as ``uint facet[1]``, firedrake_loopy.py:317-381).  Gauss-Legendre p+1 points per
    direction.  ``measure``: "dx" (args: out, coords, f) or "ds" (exterior facet; the
    local facet 2*direction + side is baked in when ``facet`` is given -- the
    extruded ds_b / ds_t kernels -- else read from a 4th argument: ds_v), or "dS" (interior
    facet; ``integrand`` "avg" = avg(f), "jump2" = (f('+') - f('-'))**2; ``facet`` = the pair of
    local facet numbers or None to read uint[2])."""
    from .fiat_lite import interval_element
    from .codegen import CStringKernel
    el = interval_element(degree)
    n = degree + 1
    Bend, _ = el.tabulate([0.0, 1.0])
    tab = lambda a: "{" + ", ".join("{" + ", ".join(repr(float(v)) for v in r) + "}" for r in a) + "}"
    vec = lambda a: "{" + ", ".join(repr(float(v)) for v in a) + "}"
    head = f"""
static const double FB[{n}][{n}] = {tab(el.B)};      /* basis a at Gauss point q: FB[q][a] */
static const double FE[2][{n}] = {tab(Bend)};        /* basis at the interval's ends */
static const double FX[{n}] = {vec(el.xq)};
static const double FW[{n}] = {vec(el.wq)};
/* columns of the Jacobian of the trilinear map at xi: J[c][d] = dx_c / dxi_d */
static inline void q1_jacobian(const double *X, const double *xi, double J[3][3])
{{
    for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) J[c][d] = 0.0;
    for (int v = 0; v < 8; ++v) {{
        const int b[3] = {{(v >> 2) & 1, (v >> 1) & 1, v & 1}};
        for (int d = 0; d < 3; ++d) {{
            double g = b[d] ? 1.0 : -1.0;
            for (int e = 0; e < 3; ++e) if (e != d) g *= b[e] ? xi[e] : 1.0 - xi[e];
            for (int c = 0; c < 3; ++c) J[c][d] += X[v * 3 + c] * g;
        }}
    }}
}}
"""
    if measure == "dx":
        name = name or "functional_dx"
        code = head + f"""
static void {name}(double *out, const double *X, const double *f)
{{
    for (int qx = 0; qx < {n}; ++qx) for (int qy = 0; qy < {n}; ++qy) for (int qz = 0; qz < {n}; ++qz) {{
        const double xi[3] = {{FX[qx], FX[qy], FX[qz]}};
        double J[3][3], v = 0.0;
        q1_jacobian(X, xi, J);
        const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
                         - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                         + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        for (int a = 0; a < {n}; ++a) for (int b = 0; b < {n}; ++b) for (int c = 0; c < {n}; ++c)
            v += f[(a * {n} + b) * {n} + c] * FB[qx][a] * FB[qy][b] * FB[qz][c];
        out[0] += FW[qx] * FW[qy] * FW[qz] * fabs(det) * v;
    }}
}}
"""
        return CStringKernel(code, name)
    if measure == "dS":
        # interior facets: coefficient and coordinate arrays are doubled, cell '+' then cell '-'
        # (tsfc/kernel_interface/common.py:518-522); local facet numbers uint[2]
        # (firedrake_loopy.py:317-381).  The two cells are conforming and equally oriented, so a
        # quadrature point has the same tangential reference coordinates in both.
        name = name or ("functional_dS" if facet is None else f"functional_dS{facet[0]}{facet[1]}")
        sig = ", const unsigned int *facet" if facet is None else ""
        getp, getm = ("facet[0]", "facet[1]") if facet is None else (str(int(facet[0])), str(int(facet[1])))
        expr = {"avg": "0.5 * (v[0] + v[1])", "jump2": "(v[0] - v[1]) * (v[0] - v[1])"}[integrand]
        code = head + f"""
static void {name}(double *out, const double *X, const double *f{sig})
{{
    const int fac[2] = {{(int)({getp}), (int)({getm})}};
    const int fd = fac[0] / 2, d1 = (fd + 1) % 3, d2 = (fd + 2) % 3;
    for (int q1 = 0; q1 < {n}; ++q1) for (int q2 = 0; q2 < {n}; ++q2) {{
        double xi[3], J[3][3], v[2];
        const double *T[3];
        for (int s = 0; s < 2; ++s) {{                 /* restriction '+' (s = 0) and '-' (s = 1) */
            T[fd] = FE[fac[s] % 2]; T[d1] = FB[q1]; T[d2] = FB[q2];
            v[s] = 0.0;
            for (int a = 0; a < {n}; ++a) for (int b = 0; b < {n}; ++b) for (int c = 0; c < {n}; ++c)
                v[s] += f[s * {n ** 3} + (a * {n} + b) * {n} + c] * T[0][a] * T[1][b] * T[2][c];
        }}
        xi[fd] = (double)(fac[0] % 2); xi[d1] = FX[q1]; xi[d2] = FX[q2];
        q1_jacobian(X, xi, J);                          /* geometry from the '+' cell */
        const double cx = J[1][d1] * J[2][d2] - J[2][d1] * J[1][d2];
        const double cy = J[2][d1] * J[0][d2] - J[0][d1] * J[2][d2];
        const double cz = J[0][d1] * J[1][d2] - J[1][d1] * J[0][d2];
        out[0] += FW[q1] * FW[q2] * sqrt(cx * cx + cy * cy + cz * cz) * ({expr});
    }}
}}
"""
        return CStringKernel(code, name)
    if measure != "ds":
        raise ValueError(f"unknown measure {measure!r}")
    name = name or ("functional_ds" if facet is None else f"functional_ds{facet}")
    sig = "const unsigned int *facet" if facet is None else ""
    get = "facet[0]" if facet is None else str(int(facet))
    code = head + f"""
static void {name}(double *out, const double *X, const double *f{", " + sig if sig else ""})
{{
    const int fd = (int)({get}) / 2, fs = (int)({get}) % 2;     /* normal direction, side */
    const int d1 = (fd + 1) % 3, d2 = (fd + 2) % 3;
    for (int q1 = 0; q1 < {n}; ++q1) for (int q2 = 0; q2 < {n}; ++q2) {{
        double xi[3], J[3][3], v = 0.0;
        const double *T[3];                       /* 1-D basis rows per direction */
        xi[fd] = (double)fs; xi[d1] = FX[q1]; xi[d2] = FX[q2];
        T[fd] = FE[fs]; T[d1] = FB[q1]; T[d2] = FB[q2];
        q1_jacobian(X, xi, J);
        /* surface element |dx/dxi_d1 x dx/dxi_d2| */
        const double cx = J[1][d1] * J[2][d2] - J[2][d1] * J[1][d2];
        const double cy = J[2][d1] * J[0][d2] - J[0][d1] * J[2][d2];
        const double cz = J[0][d1] * J[1][d2] - J[1][d1] * J[0][d2];
        for (int a = 0; a < {n}; ++a) for (int b = 0; b < {n}; ++b) for (int c = 0; c < {n}; ++c)
            v += f[(a * {n} + b) * {n} + c] * T[0][a] * T[1][b] * T[2][c];
        out[0] += FW[q1] * FW[q2] * sqrt(cx * cx + cy * cy + cz * cz) * v;
    }}
}}
"""
    return CStringKernel(code, name)


def variable_coefficient_kernel(degree, beta=0.0, name="varcoef_action"):
    """C source of the 1-form ``action(inner(kappa*grad(u), grad(v))*dx + beta*inner(u, v)*dx, u)``
    with a COEFFICIENT FIELD kappa in the same Q_p (x) P_p space -- a form outside the hand-written
    set, written the way TSFC's spectral mode would (sum factorisation, tsfc/spectral.py:24-191;
    coefficients are tabulated like arguments, tsfc/fem.py:710-804) and run through the generic
    wrapper builder.  Arguments: y (INC), coords, u, kappa."""
    from .fiat_lite import interval_element
    from .codegen import CStringKernel
    el = interval_element(degree)
    n = degree + 1
    tab = lambda a: "{" + ", ".join("{" + ", ".join(repr(float(v)) for v in r) + "}" for r in a) + "}"
    vec = lambda a: "{" + ", ".join(repr(float(v)) for v in a) + "}"
    code = f"""
#define VN {n}
static const double VB[VN][VN] = {tab(el.B)};      /* VB[q][a] */
static const double VD[VN][VN] = {tab(el.D)};      /* VD[q][a] */
static const double VX[VN] = {vec(el.xq)};
static const double VW[VN] = {vec(el.wq)};
/* out = (T applied along direction dir) in;  tr = 0: out[q] = sum_a T[q][a] in[a];  tr = 1: transpose */
static inline void vc_apply(const double T[VN][VN], int dir, int tr, const double *in, double *out)
{{
    const int st = dir == 0 ? VN * VN : (dir == 1 ? VN : 1);
    for (int i = 0; i < VN * VN * VN; ++i) {{
        const int k = (i / st) % VN, base = i - k * st;
        double s = 0.0;
        for (int a = 0; a < VN; ++a) s += (tr ? T[a][k] : T[k][a]) * in[base + a * st];
        out[i] = s;
    }}
}}
static inline void vc_tensor(const double (*T0)[VN], const double (*T1)[VN], const double (*T2)[VN], int tr,
                             const double *in, double *out)
{{
    double t1[VN * VN * VN], t2[VN * VN * VN];
    vc_apply(T0, 0, tr, in, t1);
    vc_apply(T1, 1, tr, t1, t2);
    vc_apply(T2, 2, tr, t2, out);
}}
static void {name}(double *y, const double *X, const double *u, const double *kappa)
{{
    double U[VN * VN * VN], G[3][VN * VN * VN], K[VN * VN * VN], t[VN * VN * VN];
    vc_tensor(VB, VB, VB, 0, u, U);
    vc_tensor(VD, VB, VB, 0, u, G[0]);
    vc_tensor(VB, VD, VB, 0, u, G[1]);
    vc_tensor(VB, VB, VD, 0, u, G[2]);
    vc_tensor(VB, VB, VB, 0, kappa, K);
    for (int qx = 0; qx < VN; ++qx) for (int qy = 0; qy < VN; ++qy) for (int qz = 0; qz < VN; ++qz) {{
        const int q = (qx * VN + qy) * VN + qz;
        const double xi[3] = {{VX[qx], VX[qy], VX[qz]}};
        double J[3][3] = {{{{0}}}};
        for (int v = 0; v < 8; ++v) {{
            const int b[3] = {{(v >> 2) & 1, (v >> 1) & 1, v & 1}};
            for (int d = 0; d < 3; ++d) {{
                double g = b[d] ? 1.0 : -1.0;
                for (int e = 0; e < 3; ++e) if (e != d) g *= b[e] ? xi[e] : 1.0 - xi[e];
                for (int c = 0; c < 3; ++c) J[c][d] += X[v * 3 + c] * g;
            }}
        }}
        double R[3][3];                              /* cofactor rows: R[k] . J[:, m] = det * delta_km */
        R[0][0] = J[1][1] * J[2][2] - J[2][1] * J[1][2]; R[0][1] = J[2][1] * J[0][2] - J[0][1] * J[2][2];
        R[0][2] = J[0][1] * J[1][2] - J[1][1] * J[0][2];
        R[1][0] = J[1][2] * J[2][0] - J[2][2] * J[1][0]; R[1][1] = J[2][2] * J[0][0] - J[0][2] * J[2][0];
        R[1][2] = J[0][2] * J[1][0] - J[1][2] * J[0][0];
        R[2][0] = J[1][0] * J[2][1] - J[2][0] * J[1][1]; R[2][1] = J[2][0] * J[0][1] - J[0][0] * J[2][1];
        R[2][2] = J[0][0] * J[1][1] - J[1][0] * J[0][1];
        const double det = J[0][0] * R[0][0] + J[1][0] * R[0][1] + J[2][0] * R[0][2];
        const double w = VW[qx] * VW[qy] * VW[qz];
        double h[3], f[3];
        for (int c = 0; c < 3; ++c) h[c] = R[0][c] * G[0][q] + R[1][c] * G[1][q] + R[2][c] * G[2][q];
        for (int k = 0; k < 3; ++k)
            f[k] = K[q] * w / fabs(det) * (R[k][0] * h[0] + R[k][1] * h[1] + R[k][2] * h[2]);
        G[0][q] = f[0]; G[1][q] = f[1]; G[2][q] = f[2];
        U[q] *= {float(beta)!r} * w * fabs(det);
    }}
    vc_tensor(VD, VB, VB, 1, G[0], t);  for (int i = 0; i < VN * VN * VN; ++i) y[i] += t[i];
    vc_tensor(VB, VD, VB, 1, G[1], t);  for (int i = 0; i < VN * VN * VN; ++i) y[i] += t[i];
    vc_tensor(VB, VB, VD, 1, G[2], t);  for (int i = 0; i < VN * VN * VN; ++i) y[i] += t[i];
    vc_tensor(VB, VB, VB, 1, U, t);     for (int i = 0; i < VN * VN * VN; ++i) y[i] += t[i];
}}
#undef VN
"""
    return CStringKernel(code, name)


def assemble_variable_coefficient(V: "FunctionSpace", kappa: op2.Dat, u: op2.Dat, beta=0.0, tensor=None, bcs=()):
    """``assemble(action(inner(kappa*grad(u), grad(v))*dx + beta*inner(u, v)*dx, u))`` for a scalar
    coefficient field ``kappa`` in V (generic path)."""
    if V.cdim != 1:
        raise NotImplementedError("scalar spaces only")
    if tensor is None:
        tensor = V.dat()
    tensor.zero()
    tensor.device_ptr
    op2.par_loop(variable_coefficient_kernel(V.degree, beta), V.cell_set, tensor(op2.INC, V.cell_node_map),
                 V.coordinates(op2.READ, V.coord_map), u(op2.READ, V.cell_node_map),
                 kappa(op2.READ, V.cell_node_map))
    for bc in bcs:
        bc.zero(tensor)
    return tensor


def assemble_functional(V: "FunctionSpace", f: op2.Dat, measure="dx", integrand="avg"):
    """``assemble(f*dx)`` / ``assemble(f*ds_b)`` / ``ds_t`` / ``ds_v`` / ``ds`` for a scalar
    ``f`` in V: rank-0 parloops with a Global INC argument (firedrake/assemble.py
    ZeroFormAssembler :1170-1194; the reduction of pyop2/parloop.py:411-455), through the
    generic wrapper builder.  Extruded exterior facets follow the reference's split:
    bottom / top = iteration regions ON_BOTTOM / ON_TOP over the cells, vertical = the base
    mesh's exterior facets x all layers (firedrake/assemble.py:1810-1850)."""
    from . import codegen
    if V.cdim != 1:
        raise NotImplementedError("functionals of scalar fields only")
    if measure == "ds":
        return sum(assemble_functional(V, f, m) for m in ("ds_b", "ds_t", "ds_v"))
    g = op2.Global(1, 0.0)
    p = V.degree
    if measure == "dx":
        codegen.par_loop(functional_kernel(p, "dx"), V.cell_set, g(op2.INC),
                         V.coordinates(op2.READ, V.coord_map), f(op2.READ, V.cell_node_map))
    elif measure in ("ds_b", "ds_t"):
        k = functional_kernel(p, "ds", facet=4 if measure == "ds_b" else 5)
        codegen.par_loop(k, V.cell_set, g(op2.INC), V.coordinates(op2.READ, V.coord_map),
                         f(op2.READ, V.cell_node_map),
                         iteration_region="ON_BOTTOM" if measure == "ds_b" else "ON_TOP",
                         interior_horizontal=False)
    elif measure == "ds_v":
        if not hasattr(V, "_ext_facets"):
            cells, local = V.mesh.exterior_vertical_facets()
            fset = op2.ExtrudedSet(op2.Set(len(cells)), V.mesh.layers)
            V._ext_facets = (
                fset,
                op2.Map(fset, V.node_set, V.V.arity, V.V.cell_node_map[cells], offset=V.V.offset),
                op2.Map(fset, V.vertex_set, 8, V.mesh.coord_map[cells], offset=V.mesh.coord_offset),
                op2.Dat(op2.DataSet(fset, 1), local, dtype=np.uint32))
        fset, fmap, cmap, local = V._ext_facets
        if fset.total_size:
            codegen.par_loop(functional_kernel(p, "ds"), fset, g(op2.INC), V.coordinates(op2.READ, cmap),
                             f(op2.READ, fmap), local(op2.READ))
    elif measure == "dS_h":
        # horizontal interior facets: ON_INTERIOR_FACETS over the cells, every argument packs the
        # cell below ('+') and the cell above ('-') (pyop2/codegen/builder.py:779-800, 840-844)
        k = functional_kernel(p, "dS", facet=(5, 4), integrand=integrand)
        codegen.par_loop(k, V.cell_set, g(op2.INC), V.coordinates(op2.READ, V.coord_map),
                         f(op2.READ, V.cell_node_map), iteration_region="ON_INTERIOR_FACETS")
    elif measure == "dS_v":
        # vertical interior facets: the base mesh's interior facets x all layers; maps list the
        # nodes of cell '+' then of cell '-' (firedrake/cython/dmcommon.pyx:1636-1677)
        if V.dof_dset.halo is not None:
            raise NotImplementedError("dS_v on a partitioned mesh needs an exec halo of cells")
        if not hasattr(V, "_int_facets"):
            cp, cm, local = V.mesh.interior_vertical_facets()
            fset = op2.ExtrudedSet(op2.Set(len(cp)), V.mesh.layers)
            cat = lambda m: np.concatenate([m[cp], m[cm]], axis=1)
            V._int_facets = (
                fset,
                op2.Map(fset, V.node_set, 2 * V.V.arity, cat(V.V.cell_node_map), offset=np.tile(V.V.offset, 2)),
                op2.Map(fset, V.vertex_set, 16, cat(V.mesh.coord_map), offset=np.tile(V.mesh.coord_offset, 2)),
                op2.Dat(op2.DataSet(fset, 2), local, dtype=np.uint32))
        fset, fmap, cmap, local = V._int_facets
        if fset.total_size:
            codegen.par_loop(functional_kernel(p, "dS", integrand=integrand), fset, g(op2.INC),
                             V.coordinates(op2.READ, cmap), f(op2.READ, fmap), local(op2.READ))
    elif measure == "dS":
        return sum(assemble_functional(V, f, m, integrand) for m in ("dS_h", "dS_v"))
    else:
        raise ValueError(f"unknown measure {measure!r}")
    return float(g.data_ro[0])


class DirichletBC:
    """``DirichletBC(V, g, sub_domain)``: node subset + value
    (firedrake/bcs.py:260-457)."""

    def __init__(self, V: FunctionSpace, g, sub_domain):
        self.V = V
        subs = sub_domain if isinstance(sub_domain, (list, tuple)) else [sub_domain]
        self.sub_domains = tuple(subs)
        nodes = np.unique(np.concatenate([V.boundary_nodes(s) for s in subs])).astype(np.int32)
        self.nodes = nodes
        self.node_set = op2.Subset(V.node_set, nodes)
        self.g = g

    def zero(self, dat):
        dat.zero(self.node_set)

    def set(self, dat, val):
        """dat[nodes] = val[nodes] (val a Dat) or the scalar val."""
        from . import _lib
        L = _lib.lib()
        if not hasattr(self, "_dev_nodes"):
            self._dev_nodes = op2.DeviceArray.from_host(self.nodes)
        if isinstance(val, op2.Dat):
            _lib.check(L.fdb_dat_set_nodes(dat.device_ptr, val.device_ptr, dat.cdim,
                                           self._dev_nodes.ptr, len(self.nodes)))
        else:
            _lib.check(L.fdb_dat_set_nodes_scalar(dat.device_ptr, float(val), dat.cdim,
                                                  self._dev_nodes.ptr, len(self.nodes)))
        dat._device_written()

    def apply(self, dat):
        self.set(dat, self.g)

    def lgmap(self):
        lg = np.arange(self.V.node_count, dtype=np.int32)
        lg[self.nodes] = -1
        return lg


@dataclass
class Form:
    """alpha*inner(grad(u), grad(v))*dx + beta*inner(u, v)*dx on ``V``."""
    V: FunctionSpace
    alpha: float = 1.0
    beta: float = 0.0

    def kernel(self, rank):
        import os
        # per-cell metric on meshes whose cells are all parallelepipeds (checked on the device;
        # GPU-validated in round 2, FDB_AFFINE=0 opts out)
        affine = rank == 1 and os.environ.get("FDB_AFFINE", "1") != "0" and self.V.cells_are_affine()
        return op2.Kernel("helmholtz", degree=self.V.degree, alpha=self.alpha, beta=self.beta,
                          rank=rank, cdim=self.V.cdim, affine=affine)


def poisson(V):
    return Form(V, 1.0, 0.0)


def helmholtz(V):
    return Form(V, 1.0, 1.0)


def mass(V):
    return Form(V, 0.0, 1.0)


class OneFormAssembler:
    """Cached assembler of ``action(a, u)`` (firedrake/assemble.py:950-977,
    1073-1096: parloops are built once and re-run)."""

    def __init__(self, form: Form, u: op2.Dat, bcs=(), scatter="atomic"):
        self.form, self.u, self.bcs = form, u, tuple(bcs)
        V = form.V
        self._gk = op2.GlobalKernel(form.kernel(1), [V.cell_node_map, V.coord_map], extruded=True,
                                    scatter=scatter)
        self._loop = None

    def assemble(self, tensor=None):
        V = self.form.V
        if tensor is None:
            tensor = V.dat()
        if self._loop is None or self._tensor is not tensor:
            self._tensor = tensor
            self._loop = op2.Parloop(self._gk, V.cell_set,
                                     [tensor(op2.INC, V.cell_node_map),
                                      V.coordinates(op2.READ, V.coord_map),
                                      self.u(op2.READ, V.cell_node_map)], location="device")
        tensor.zero()
        self._loop()
        for bc in self.bcs:
            bc.zero(tensor)
        return tensor


def assemble(form: Form, u=None, tensor=None, bcs=(), mat_type="aij"):
    """``assemble(action(a, u))`` when ``u`` is given (-> Dat), else the
    bilinear form: ``mat_type="aij"`` -> :class:`op2.Mat`, ``"is"`` -> :class:`ISMat` (distributed,
    unassembled), ``"matfree"`` -> :class:`ImplicitMatrixContext`."""
    V = form.V
    bcs = tuple(bcs)
    if u is not None:
        return OneFormAssembler(form, u, bcs).assemble(tensor)
    if mat_type == "matfree":
        return ImplicitMatrixContext(form, bcs)
    if tensor is None:
        dsets = (V.node_set, V.node_set) if V.cdim == 1 else (V.dof_dset, V.dof_dset)
        tensor = op2.Mat(op2.Sparsity(dsets, [(V.cell_node_map, V.cell_node_map, None)]))
    tensor.zero()
    lg = None
    if bcs and V.cdim == 1:
        lgm = np.arange(V.node_count, dtype=np.int32)
        for bc in bcs:
            lgm[bc.nodes] = -1
        lg = (lgm, lgm)
    elif bcs:
        # vector-valued space: dof-level lgmap, every component of a constrained node masked
        lgm = np.arange(V.node_count * V.cdim, dtype=np.int32).reshape(-1, V.cdim)
        for bc in bcs:
            lgm[bc.nodes, :] = -1
        lgm = np.ascontiguousarray(lgm.ravel())
        lg = (lgm, lgm)
    op2.par_loop(form.kernel(2), V.cell_set,
                 tensor(op2.INC, (V.cell_node_map, V.cell_node_map), lgmaps=lg),
                 V.coordinates(op2.READ, V.coord_map))
    owned = V.node_set.size
    for bc in bcs:
        # on a partitioned space a constrained node gets its unit diagonal from its OWNER only:
        # the ghost copies' rows stay empty and add nothing in the local->global sum of ISMat.mult
        rows = bc.nodes[bc.nodes < owned] if mat_type == "is" else bc.nodes
        tensor.set_local_diagonal_entries(rows, 1.0)
    tensor.assemble()
    if mat_type == "is":
        return ISMat(V, tensor)
    if V.dof_dset.halo is not None:
        raise NotImplementedError("assembled matrices on a partitioned space are mat_type='is' "
                                  "(unassembled, one block per GPU) or 'matfree'")
    return tensor


class ISMat:
    """``mat_type="is"``: the distributed matrix kept UNASSEMBLED, A = sum_r R_r^T A_r R_r with A_r the
    matrix of rank r's owned cells over its owned + ghost dofs (PETSc MATIS, which the reference supports:
    firedrake/assemble.py:1330-1345, pyop2/types/mat.py:930-933).  No matrix entries ever cross GPUs (the
    reference's MatAssembly stash exchange, SURVEY.md section 2.3 C4, disappears); ``mult`` is a local
    SpMV followed by the local->global halo sum that vectors use anyway."""

    def __init__(self, V: FunctionSpace, local: op2.Mat):
        self.V, self.local = V, local

    def mult(self, X: op2.Dat, Y: op2.Dat):
        halo = self.V.dof_dset.halo
        if halo is not None and not X.halo_valid:
            halo.global_to_local_begin(X)
            halo.global_to_local_end(X)
        self.local.mult(X, Y)
        if halo is not None:
            halo.local_to_global_begin(Y)
            halo.local_to_global_end(Y)
        return Y


class ImplicitMatrixContext:
    """Matrix-free operator (firedrake/matrix_free/operators.py:74-242): ``mult``
    = zero the column-BC entries of x, assemble ``action(a, x)``, write x back on
    the row-BC entries (identity on constrained rows).  x and y stay on the
    device across calls."""

    def __init__(self, form: Form, bcs=()):
        self.form, self.bcs = form, tuple(bcs)
        V = form.V
        self._x = V.dat()
        self._assembler = OneFormAssembler(form, self._x, ())

    def getDiagonal(self, D: op2.Dat):
        """``assemble(a, diagonal=True)`` then 1 on the constrained rows
        (matrix_free/operators.py:199-205; firedrake/assemble.py:1226-1241)."""
        V = self.form.V
        k = op2.Kernel("helmholtz", degree=V.degree, alpha=self.form.alpha, beta=self.form.beta,
                       diagonal=True)
        D.zero()
        op2.par_loop(k, V.cell_set, D(op2.INC, V.cell_node_map), V.coordinates(op2.READ, V.coord_map))
        for bc in self.bcs:
            bc.set(D, 1.0)
        return D

    def duplicate(self, copy=True):
        """``MatDuplicate`` of the python-context matrix (matrix_free/operators.py:451-470): a new
        context on the same form and conditions (nothing is assembled, so nothing is copied)."""
        if not copy:
            raise NotImplementedError("cannot duplicate a matrix-free operator without its values (copy=0)")
        return ImplicitMatrixContext(self.form, self.bcs)

    def createSubMatrix(self, row_is, col_is=None):
        """``MatCreateSubMatrix`` (matrix_free/operators.py:380-447).  The spaces here have ONE field,
        so an index set is either the whole dof range -- the reference then rebuilds the context on the
        extracted sub-form, i.e. on the same form -- or an arbitrary one, for which the reference falls
        back to PETSc's virtual sub-matrix (``MatCreateSubMatrixVirtual``: scatter the sub-vector into
        a zero full vector, apply, gather the rows): :class:`SubMatrixContext`."""
        col_is = row_is if col_is is None else col_is
        n = self.form.V.node_count * self.form.V.cdim
        whole = lambda s: len(s) == n and np.array_equal(np.asarray(s), np.arange(n))
        if whole(row_is) and whole(col_is):
            return self.duplicate()
        return SubMatrixContext(self, row_is, col_is)

    def multTranspose(self, X: op2.Dat, Y: op2.Dat):
        """``Y = A^T X`` (matrix_free/operators.py:245-330: the action of ``adjoint(a)`` with the
        row and column conditions exchanged).  Every form of the supported family is
        symmetric and row/column DirichletBCs coincide here (no EquationBC), so A^T = A."""
        return self.mult(X, Y)

    def mult(self, X: op2.Dat, Y: op2.Dat):
        from . import _lib
        L = _lib.lib()
        _lib.check(L.fdb_memcpy_d2d(self._x.device_ptr, X.device_ptr, X.nbytes))
        self._x._device_written()
        self._x.halo_valid = False
        for bc in self.bcs:
            bc.zero(self._x)
        self._assembler.assemble(tensor=Y)
        for bc in self.bcs:
            bc.set(Y, X)
        return Y


class SubMatrixContext:
    """Virtual sub-matrix ``A[rows, cols]`` of a matrix-free operator: ``mult(xs, ys)`` with compact
    sub-vectors (plain device Dats of len(cols) / len(rows) entries; dof indices = node*cdim + comp)."""

    def __init__(self, parent, rows, cols):
        self.parent = parent
        self.rows = np.ascontiguousarray(rows, dtype=np.int32)
        self.cols = np.ascontiguousarray(cols, dtype=np.int32)
        self._drows = op2.DeviceArray.from_host(self.rows)
        self._dcols = op2.DeviceArray.from_host(self.cols)
        V = parent.form.V
        self._x, self._y = V.dat(), V.dat()
        self.row_set, self.col_set = op2.Set(len(self.rows)), op2.Set(len(self.cols))

    def mult(self, xs: op2.Dat, ys: op2.Dat):
        from . import _lib
        L = _lib.lib()
        self._x.zero()
        _lib.check(L.fdb_vec_scatter(len(self.cols), self._dcols.ptr, xs.device_ptr, self._x.device_ptr))
        self._x._device_written()
        self.parent.mult(self._x, self._y)
        _lib.check(L.fdb_vec_gather(len(self.rows), self._drows.ptr, self._y.device_ptr, ys.device_ptr))
        ys._device_written()
        return ys


def cg(A, b: op2.Dat, x: op2.Dat, rtol=1e-8, atol=0.0, maxit=1000, allreduce=None):
    """Unpreconditioned conjugate gradients on device-resident Dats (the solve
    of demos/matrix_free/poisson.py.rst:38-47 with ``ksp_type cg, pc_type
    none``).  ``A`` needs ``mult(X, Y)``.  Returns (iterations, residual norms).
    ``allreduce``: callable summing a scalar over ranks (owned dofs only)."""
    V = b.dataset
    r = op2.Dat(V)
    p = op2.Dat(V)
    Ap = op2.Dat(V)
    n_owned = b.dataset.set.size * b.cdim

    def dot(a, c):
        import ctypes as C
        from . import _lib
        out = C.c_double()
        _lib.check(_lib.lib().fdb_vec_dot(n_owned, a.device_ptr, c.device_ptr, C.byref(out)))
        return allreduce(out.value) if allreduce else out.value

    from . import _lib
    L = _lib.lib()
    n = b._data.size
    A.mult(x, Ap)
    _lib.check(L.fdb_memcpy_d2d(r.device_ptr, b.device_ptr, b.nbytes))
    r._device_written()
    _lib.check(L.fdb_vec_axpy(n, -1.0, Ap.device_ptr, r.device_ptr))
    r._device_written()
    _lib.check(L.fdb_memcpy_d2d(p.device_ptr, r.device_ptr, r.nbytes))
    p._device_written()
    rr = dot(r, r)
    r0 = np.sqrt(rr)
    hist = [r0]
    it = 0
    while it < maxit and np.sqrt(rr) > max(rtol * r0, atol):
        A.mult(p, Ap)
        alpha = rr / dot(p, Ap)
        # raw vector updates over all local rows: Ap's ghost rows hold partial sums, so the
        # ghost rows of x, r and p are NOT current afterwards (_device_written invalidates them;
        # A.mult refreshes p's through global_to_local when the operator reads ghosts)
        _lib.check(L.fdb_vec_axpy(n, alpha, p.device_ptr, x.device_ptr))
        x._device_written()
        _lib.check(L.fdb_vec_axpy(n, -alpha, Ap.device_ptr, r.device_ptr))
        r._device_written()
        rr_new = dot(r, r)
        _lib.check(L.fdb_vec_aypx(n, rr_new / rr, r.device_ptr, p.device_ptr))   # p = r + beta p
        p._device_written()
        rr = rr_new
        hist.append(np.sqrt(rr))
        it += 1
    x._device_written()
    return it, hist


def solve(form: Form, L: op2.Dat, u: op2.Dat, bcs=(), solver_parameters=None, hierarchy=None, allreduce=None):
    """``solve(a == L, u, bcs=bcs, solver_parameters=...)`` for the supported forms
    (firedrake/solving.py:128-260 -> LinearVariationalSolver): assemble the
    operator, lift the Dirichlet values, run the Krylov solver on the device.

    ``L``: the assembled right-hand side (a Dat, e.g. ``assemble(mass(V), u=f)``).
    ``solver_parameters`` (PETSc option names, the subset that makes sense here):
    ``mat_type`` "matfree" (default) | "aij" | "is"; ``ksp_type`` "cg"; ``pc_type`` "none" (default) |
    "jacobi" | "mg" (needs ``hierarchy``, a mg.MeshHierarchy whose finest mesh is ``form.V.mesh``);
    ``ksp_rtol`` (1e-8), ``ksp_max_it`` (1000).  Returns (iterations, residual history)."""
    from . import _lib
    sp = {"mat_type": "matfree", "ksp_type": "cg", "pc_type": "none", "ksp_rtol": 1e-8, "ksp_max_it": 1000}
    sp.update(solver_parameters or {})
    if sp["ksp_type"] != "cg":
        raise NotImplementedError("ksp_type cg only (the supported forms are symmetric positive definite)")
    V = form.V
    bcs = tuple(bcs)
    lib = _lib.lib()
    n = L._data.size
    # lifting (firedrake/assemble.py:1243-1254 + linear solver's rhs): u = g on the constrained nodes,
    # solve A (u - g) = L - K g on the free rows with homogeneous conditions
    g = V.dat()
    g.device_ptr
    for bc in bcs:
        bc.apply(g)
    lift = any(not (np.isscalar(bc.g) and bc.g == 0.0) for bc in bcs)
    b = V.dat()
    L.copy(b)
    if lift:
        Kg = OneFormAssembler(form, g, ()).assemble()
        b.axpy(-1.0, Kg)
    for bc in bcs:
        bc.zero(b)
    A = assemble(form, bcs=bcs, mat_type=sp["mat_type"])
    u.zero()
    u.device_ptr
    pc = sp["pc_type"]
    if pc == "none":
        its, hist = cg(A, b, u, rtol=sp["ksp_rtol"], maxit=sp["ksp_max_it"], allreduce=allreduce)
    else:
        from . import mg as _mg
        if pc == "jacobi":
            ctx = A if isinstance(A, ImplicitMatrixContext) else ImplicitMatrixContext(form, bcs)
            d = ctx.getDiagonal(V.dat())
            op2.par_loop(op2.Kernel("static void recip(double *w) { *w = 1.0 / *w; }", "recip"), V.node_set,
                         d(op2.RW))

            def M(r, z):
                _lib.check(lib.fdb_vec_pointwise_mult(n, r.device_ptr, d.device_ptr, z.device_ptr))
                z._device_written()
        elif pc == "mg":
            if hierarchy is None:
                raise ValueError("pc_type mg needs the mesh hierarchy")
            vc = _mg.VCycle(hierarchy, V.degree, lambda W: Form(W, form.alpha, form.beta),
                            bc_domains=tuple(s for bc in bcs for s in bc.sub_domains), allreduce=allreduce)
            top = len(hierarchy) - 1
            M = lambda r, z: vc.apply(top, r, z)
        else:
            raise NotImplementedError(f"pc_type {pc!r}")
        its, hist = _mg.pcg(A, b, u, M, rtol=sp["ksp_rtol"], maxit=sp["ksp_max_it"], allreduce=allreduce)
    if lift:
        u.axpy(1.0, g)
    else:
        for bc in bcs:
            bc.apply(u)
    return its, hist


class DGAdvection:
    """Right-hand side ``assemble(L1)`` of the DG advection demo (reference
    demos/DG_advection/DG_advection.py.rst:182-217) on a :class:`QuadMesh`.

    ``fused=False``: three parloops -- cells, exterior facets, interior facets --
    exactly the kernels ``OneFormAssembler`` would run
    (firedrake/assemble.py:1069-1096), with the reference's facet-kernel ABI.
    ``fused=True``: ONE owner-computes pass over the cells (each cell adds its
    cell term and the fluxes through its four facets into its own dofs): no
    atomics, deterministic, ~3x less traffic; the only off-cell data is the q of
    neighbouring cells, i.e. a ghost-cell halo when the mesh is a slab of a
    partitioned run (``halo`` = ``firedrake_b200.halo.Halo`` on the DQ dofs).
    """

    def __init__(self, mesh, dt, q_in=1.0, nq=3, fused=False, halo=None):
        self.mesh, self.fused = mesh, fused
        C_ = mesh.num_cells
        owned = mesh.num_owned_cells
        self.cell_set = op2.Set((owned, owned, C_))
        self.dq_nodes = op2.Set((4 * owned, 4 * owned, 4 * C_))
        self.dq_dset = op2.DataSet(self.dq_nodes, 1, halo=halo)
        self.vertices = op2.Set(mesh.node_count)
        dg, cg = mesh.dg1_map, mesh.coord_map
        self.cell_dq = op2.Map(self.cell_set, self.dq_nodes, 4, dg)
        self.cell_cg = op2.Map(self.cell_set, self.vertices, 4, cg)
        self.coordinates = op2.Dat(op2.DataSet(self.vertices, 2), mesh.coordinates)
        self.consts = op2.Global(2, [dt, q_in])
        self.nq = nq
        self._loops = None
        if fused:
            self.nbr = op2.Dat(op2.DataSet(self.cell_set, 4), mesh.nbr, dtype=np.int32)
            self.nbr_facet = op2.Dat(op2.DataSet(self.cell_set, 4), mesh.nbr_facet, dtype=np.uint32)
            return
        if halo is not None:
            raise NotImplementedError("the facet-loop form runs unpartitioned; use fused=True")
        self.ext_set = op2.Set(len(mesh.ext_facet_cells))
        self.int_set = op2.Set(len(mesh.int_facet_cells))
        # facet->node maps: the nodes of cell '+' then of cell '-'
        # (firedrake/cython/dmcommon.pyx:1636-1677)
        self.ext_dq = op2.Map(self.ext_set, self.dq_nodes, 4, dg[mesh.ext_facet_cells])
        self.ext_cg = op2.Map(self.ext_set, self.vertices, 4, cg[mesh.ext_facet_cells])
        ic = mesh.int_facet_cells
        self.int_dq = op2.Map(self.int_set, self.dq_nodes, 8, np.concatenate([dg[ic[:, 0]], dg[ic[:, 1]]], axis=1))
        self.int_cg = op2.Map(self.int_set, self.vertices, 8, np.concatenate([cg[ic[:, 0]], cg[ic[:, 1]]], axis=1))
        self.ext_facet = op2.Dat(op2.DataSet(self.ext_set, 1), mesh.ext_facet_local, dtype=np.uint32)
        self.int_facet = op2.Dat(op2.DataSet(self.int_set, 2), mesh.int_facet_local, dtype=np.uint32)

    def function(self, data=None):
        return op2.Dat(self.dq_dset, data)

    def velocity(self, data):
        return op2.Dat(op2.DataSet(self.vertices, 2), data)

    def assemble(self, q, u, tensor=None):
        if tensor is None:
            tensor = self.function()
        key = (id(q), id(u), id(tensor))
        if self._loops is None or self._key != key:
            self._key = key
            mk = lambda integral: op2.Kernel("dg_advection", degree=1, integral=integral, nq=self.nq)
            X, G = self.coordinates, self.consts
            gk = lambda k, maps: op2.GlobalKernel(k, maps, extruded=False)
            if self.fused:
                self._loops = [
                    op2.Parloop(gk(mk("fused"), [self.cell_dq, self.cell_cg]), self.cell_set,
                                [tensor(op2.INC, self.cell_dq), X(op2.READ, self.cell_cg),
                                 q(op2.READ, self.cell_dq), u(op2.READ, self.cell_cg), G(op2.READ),
                                 self.nbr_facet(op2.READ), self.nbr(op2.READ)])]
            else:
                self._loops = [
                    op2.Parloop(gk(mk("cell"), [self.cell_dq, self.cell_cg]), self.cell_set,
                                [tensor(op2.INC, self.cell_dq), X(op2.READ, self.cell_cg), q(op2.READ, self.cell_dq),
                                 u(op2.READ, self.cell_cg), G(op2.READ)]),
                    op2.Parloop(gk(mk("exterior_facet"), [self.ext_dq, self.ext_cg]), self.ext_set,
                                [tensor(op2.INC, self.ext_dq), X(op2.READ, self.ext_cg), q(op2.READ, self.ext_dq),
                                 u(op2.READ, self.ext_cg), G(op2.READ), self.ext_facet(op2.READ)]),
                    op2.Parloop(gk(mk("interior_facet"), [self.int_dq, self.int_cg]), self.int_set,
                                [tensor(op2.INC, self.int_dq), X(op2.READ, self.int_cg), q(op2.READ, self.int_dq),
                                 u(op2.READ, self.int_cg), G(op2.READ), self.int_facet(op2.READ)]),
                ]
        tensor.zero()
        if self.fused and q.dataset.halo is not None and not q.halo_valid:
            # ghost-cell q must be current before ANY cell runs (every cell may border a ghost)
            q.dataset.halo.global_to_local_begin(q)
            q.dataset.halo.global_to_local_end(q)
        tensor.frozen_halo = True          # owner-computes: nothing to send back
        for loop in self._loops:
            loop()
        tensor.frozen_halo = False
        return tensor


def dg_slab(nx, ny, rank, nranks):
    """Slab of the nx x ny quad mesh for ``rank`` with ghost-cell columns, and
    the (rank, send, recv) halo lists of its DQ1 dofs."""
    from .partition import slab_bounds
    from .utility_meshes import QuadMesh
    x0, x1 = slab_bounds(nx, nranks, rank)
    mesh = QuadMesh(x1 - x0, ny, ix0=x0, nx_global=nx, ghost_left=rank > 0,
                    ghost_right=rank < nranks - 1)
    dofs = lambda cells: (np.asarray(cells)[:, None] * 4 + np.arange(4)[None, :]).ravel().astype(np.int32)
    neigh = []
    if rank > 0:
        neigh.append((rank - 1, dofs(mesh.first_owned_column), dofs(mesh.ghost_cells_left)))
    if rank < nranks - 1:
        neigh.append((rank + 1, dofs(mesh.last_owned_column), dofs(mesh.ghost_cells_right)))
    return mesh, neigh
