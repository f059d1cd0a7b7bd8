// DG advection right-hand side on quadrilaterals (config 3): the cell,
// exterior-facet and interior-facet kernels of
//   L1 = dtc*( q*div(phi*u)*dx - [u.n<0] phi u.n q_in ds - [u.n>0] phi u.n q ds
//              - (phi('+') - phi('-'))*(un('+') q('+') - un('-') q('-')) dS )
// (reference demos/DG_advection/DG_advection.py.rst:182-217), q/phi in DQ1,
// u in vector CG1, Q1 coordinates.  Facet-kernel ABI as in the reference
// (tsfc/kernel_interface/firedrake_loopy.py:317-381): map rows carry the '+'
// cell's nodes then the '-' cell's, and a uint32 Dat holds the local facet
// numbers (0: x=0, 1: x=1, 2: y=0, 3: y=1).
//
// These are small, HBM/latency-bound kernels (~50-100 flop per dof): one
// thread per cell / facet, coalesced 128-bit loads of the map rows, grid sized
// to the SM count; the scatter is a handful of RED.ADD.F64 per thread.
#include "common.cuh"

namespace {

struct DgParams {
    double *out;
    const double *coords;
    const double *q;
    const double *u;
    const int *dgmap;      // (n, 4) or (n, 8)
    const int *cgmap;      // (n, 4) or (n, 8)
    const unsigned *facet; // (n, 1) or (n, 2)
    const int *nbr;        // fused kernel: (ncells, 4) neighbour cell per local facet, -1 = boundary
    const unsigned *nbr_facet;   // (ncells, 4) the neighbour's local facet number
    int start, end;
    const int *subset;
    int nq;
    double dt, q_in;
    double Bend[4];        // DQ1 1-D basis at x = 0 / x = 1: Bend[e*2 + i]
    double wq[FDB_MAX_1D], xq[FDB_MAX_1D];
};

struct Q1Cell {
    double c[8];   // coords, local ax*2+ay, component fastest
    double u[8];
};

__device__ __forceinline__ void jac(const double *c, double x, double y, double J[2][2])
{
    const double bx[2] = {1.0 - x, x}, by[2] = {1.0 - y, y};
    J[0][0] = J[0][1] = J[1][0] = J[1][1] = 0.0;
#pragma unroll
    for (int ax = 0; ax < 2; ax++)
#pragma unroll
        for (int ay = 0; ay < 2; ay++)
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const double X = c[(ax * 2 + ay) * 2 + a];
                J[a][0] += X * (ax ? 1.0 : -1.0) * by[ay];
                J[a][1] += X * bx[ax] * (ay ? 1.0 : -1.0);
            }
}

__device__ __forceinline__ void p1(const double *v, double x, double y, double out[2])
{
    const double bx[2] = {1.0 - x, x}, by[2] = {1.0 - y, y};
    out[0] = out[1] = 0.0;
#pragma unroll
    for (int ax = 0; ax < 2; ax++)
#pragma unroll
        for (int ay = 0; ay < 2; ay++) {
            out[0] += v[(ax * 2 + ay) * 2] * bx[ax] * by[ay];
            out[1] += v[(ax * 2 + ay) * 2 + 1] * bx[ax] * by[ay];
        }
}

__device__ __forceinline__ double dq(const DgParams &P, int i, double x)
{
    return P.Bend[i] * (1.0 - x) + P.Bend[2 + i] * x;
}
__device__ __forceinline__ double ddq(const DgParams &P, int i) { return P.Bend[2 + i] - P.Bend[i]; }

__device__ __forceinline__ void load_cell(const DgParams &P, const int *cg, Q1Cell &K)
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int v = cg[i];
        const double2 X = *reinterpret_cast<const double2 *>(P.coords + 2 * (long long)v);
        const double2 U = *reinterpret_cast<const double2 *>(P.u + 2 * (long long)v);
        K.c[2 * i] = X.x; K.c[2 * i + 1] = X.y;
        K.u[2 * i] = U.x; K.u[2 * i + 1] = U.y;
    }
}

__device__ __forceinline__ void facet_point(int f, double s, double &x, double &y, double nref[2],
                                            double tref[2])
{
    const bool vert = f < 2;               // x = const facets
    x = vert ? (double)(f & 1) : s;
    y = vert ? s : (double)(f & 1);
    const double sgn = (f & 1) ? 1.0 : -1.0;
    nref[0] = vert ? sgn : 0.0;
    nref[1] = vert ? 0.0 : sgn;
    tref[0] = vert ? 0.0 : 1.0;
    tref[1] = vert ? 1.0 : 0.0;
}

__device__ __forceinline__ void facet_geometry(const double *c, double x, double y, const double nref[2],
                                               const double tref[2], double n[2], double &ds)
{
    double J[2][2];
    jac(c, x, y, J);
    // J^{-T} nref is parallel to (cofactor matrix) nref; normalised below
    double nn0 = J[1][1] * nref[0] - J[1][0] * nref[1];
    double nn1 = -J[0][1] * nref[0] + J[0][0] * nref[1];
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    if (det < 0.0) { nn0 = -nn0; nn1 = -nn1; }
    const double inv = rsqrt(nn0 * nn0 + nn1 * nn1);
    n[0] = nn0 * inv;
    n[1] = nn1 * inv;
    const double t0 = J[0][0] * tref[0] + J[0][1] * tref[1], t1 = J[1][0] * tref[0] + J[1][1] * tref[1];
    ds = sqrt(t0 * t0 + t1 * t1);
}

__global__ void __launch_bounds__(128) dg_cell_kernel(const __grid_constant__ DgParams P)
{
    for (int i = P.start + blockIdx.x * blockDim.x + threadIdx.x; i < P.end; i += gridDim.x * blockDim.x) {
        const int n = P.subset ? P.subset[i] : i;
        const int4 dg = *reinterpret_cast<const int4 *>(P.dgmap + 4 * (long long)n);
        const int4 cg = *reinterpret_cast<const int4 *>(P.cgmap + 4 * (long long)n);
        const int dgi[4] = {dg.x, dg.y, dg.z, dg.w}, cgi[4] = {cg.x, cg.y, cg.z, cg.w};
        Q1Cell K;
        load_cell(P, cgi, K);
        double ql[4], A[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; k++) ql[k] = P.q[dgi[k]];
        for (int qx = 0; qx < P.nq; qx++)
            for (int qy = 0; qy < P.nq; qy++) {
                const double x = P.xq[qx], y = P.xq[qy];
                double J[2][2];
                jac(K.c, x, y, J);
                const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
                const double id = 1.0 / det;
                const double Ki[2][2] = {{J[1][1] * id, -J[0][1] * id}, {-J[1][0] * id, J[0][0] * id}};
                const double w = fabs(det) * P.wq[qx] * P.wq[qy];
                double uv[2];
                p1(K.u, x, y, uv);
                const double bx[2] = {1.0 - x, x}, by[2] = {1.0 - y, y};
                double divu = 0.0, qv = 0.0;
#pragma unroll
                for (int ax = 0; ax < 2; ax++)
#pragma unroll
                    for (int ay = 0; ay < 2; ay++) {
                        const double g0 = (ax ? 1.0 : -1.0) * by[ay], g1 = bx[ax] * (ay ? 1.0 : -1.0);
                        divu += K.u[(ax * 2 + ay) * 2] * (Ki[0][0] * g0 + Ki[1][0] * g1)
                              + K.u[(ax * 2 + ay) * 2 + 1] * (Ki[0][1] * g0 + Ki[1][1] * g1);
                        qv += ql[ax * 2 + ay] * dq(P, ax, x) * dq(P, ay, y);
                    }
#pragma unroll
                for (int ax = 0; ax < 2; ax++)
#pragma unroll
                    for (int ay = 0; ay < 2; ay++) {
                        const double ph = dq(P, ax, x) * dq(P, ay, y);
                        const double g0 = ddq(P, ax) * dq(P, ay, y), g1 = dq(P, ax, x) * ddq(P, ay);
                        const double gp0 = Ki[0][0] * g0 + Ki[1][0] * g1, gp1 = Ki[0][1] * g0 + Ki[1][1] * g1;
                        A[ax * 2 + ay] += P.dt * w * qv * (gp0 * uv[0] + gp1 * uv[1] + ph * divu);
                    }
            }
#pragma unroll
        for (int k = 0; k < 4; k++) atomicAdd(P.out + dgi[k], A[k]);
    }
}

__global__ void __launch_bounds__(128) dg_exterior_kernel(const __grid_constant__ DgParams P)
{
    for (int i = P.start + blockIdx.x * blockDim.x + threadIdx.x; i < P.end; i += gridDim.x * blockDim.x) {
        const int f = P.subset ? P.subset[i] : i;
        const int4 dg = *reinterpret_cast<const int4 *>(P.dgmap + 4 * (long long)f);
        const int4 cg = *reinterpret_cast<const int4 *>(P.cgmap + 4 * (long long)f);
        const int dgi[4] = {dg.x, dg.y, dg.z, dg.w}, cgi[4] = {cg.x, cg.y, cg.z, cg.w};
        Q1Cell K;
        load_cell(P, cgi, K);
        double ql[4], A[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; k++) ql[k] = P.q[dgi[k]];
        const int lf = (int)P.facet[f];
        for (int k = 0; k < P.nq; k++) {
            double x, y, nref[2], tref[2], n[2], ds, uv[2];
            facet_point(lf, P.xq[k], x, y, nref, tref);
            facet_geometry(K.c, x, y, nref, tref, n, ds);
            p1(K.u, x, y, uv);
            const double udn = uv[0] * n[0] + uv[1] * n[1];
            double qv = 0.0;
#pragma unroll
            for (int ax = 0; ax < 2; ax++)
#pragma unroll
                for (int ay = 0; ay < 2; ay++) qv += ql[ax * 2 + ay] * dq(P, ax, x) * dq(P, ay, y);
            const double flux = (udn < 0.0 ? udn * P.q_in : 0.0) + (udn > 0.0 ? udn * qv : 0.0);
#pragma unroll
            for (int ax = 0; ax < 2; ax++)
#pragma unroll
                for (int ay = 0; ay < 2; ay++)
                    A[ax * 2 + ay] -= P.dt * ds * P.wq[k] * dq(P, ax, x) * dq(P, ay, y) * flux;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) atomicAdd(P.out + dgi[k], A[k]);
    }
}

__global__ void __launch_bounds__(128) dg_interior_kernel(const __grid_constant__ DgParams P)
{
    for (int i = P.start + blockIdx.x * blockDim.x + threadIdx.x; i < P.end; i += gridDim.x * blockDim.x) {
        const int f = P.subset ? P.subset[i] : i;
        int dgi[8], cgi[8];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int4 a = *reinterpret_cast<const int4 *>(P.dgmap + 8 * (long long)f + 4 * h);
            const int4 b = *reinterpret_cast<const int4 *>(P.cgmap + 8 * (long long)f + 4 * h);
            dgi[4 * h] = a.x; dgi[4 * h + 1] = a.y; dgi[4 * h + 2] = a.z; dgi[4 * h + 3] = a.w;
            cgi[4 * h] = b.x; cgi[4 * h + 1] = b.y; cgi[4 * h + 2] = b.z; cgi[4 * h + 3] = b.w;
        }
        Q1Cell Kp, Km;
        load_cell(P, cgi, Kp);
        load_cell(P, cgi + 4, Km);
        double ql[8], A[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 8; k++) ql[k] = P.q[dgi[k]];
        const int fp = (int)P.facet[2 * (long long)f], fm = (int)P.facet[2 * (long long)f + 1];
        for (int k = 0; k < P.nq; k++) {
            double xp, yp, xm, ym, nrp[2], trp[2], nrm[2], trm[2], np_[2], nm[2], dsp, dsm, up[2], um[2];
            facet_point(fp, P.xq[k], xp, yp, nrp, trp);
            facet_point(fm, P.xq[k], xm, ym, nrm, trm);
            facet_geometry(Kp.c, xp, yp, nrp, trp, np_, dsp);
            facet_geometry(Km.c, xm, ym, nrm, trm, nm, dsm);
            p1(Kp.u, xp, yp, up);
            p1(Km.u, xm, ym, um);
            const double udnp = up[0] * np_[0] + up[1] * np_[1], udnm = um[0] * nm[0] + um[1] * nm[1];
            const double unp = 0.5 * (udnp + fabs(udnp)), unm = 0.5 * (udnm + fabs(udnm));
            double qp = 0.0, qm = 0.0;
#pragma unroll
            for (int ax = 0; ax < 2; ax++)
#pragma unroll
                for (int ay = 0; ay < 2; ay++) {
                    qp += ql[ax * 2 + ay] * dq(P, ax, xp) * dq(P, ay, yp);
                    qm += ql[4 + ax * 2 + ay] * dq(P, ax, xm) * dq(P, ay, ym);
                }
            const double jump = (unp * qp - unm * qm) * P.dt * dsp * P.wq[k];
#pragma unroll
            for (int ax = 0; ax < 2; ax++)
#pragma unroll
                for (int ay = 0; ay < 2; ay++) {
                    A[ax * 2 + ay] -= dq(P, ax, xp) * dq(P, ay, yp) * jump;
                    A[4 + ax * 2 + ay] += dq(P, ax, xm) * dq(P, ay, ym) * jump;
                }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) atomicAdd(P.out + dgi[k], A[k]);
    }
}

// Owner-computes fusion of the three integrals: one thread per cell adds the
// cell term and, for each of its four facets, the flux into ITS OWN test
// functions -- exterior inflow/outflow, or the upwind interior flux using the
// neighbour's q (read only).  From the form, for either side of an interior
// facet:  -phi_me (un_me q_me - un_nb q_nb),  un_me = max(u.n_me, 0),
// un_nb = max(-u.n_me, 0)  (u continuous, n_nb = -n_me).  Every output dof is
// written by exactly one thread: no atomics, deterministic, and q of
// neighbouring cells is the only off-cell data (= the ghost-cell halo of a
// partitioned run).  Same result as the cell + exterior + interior parloops.
__global__ void __launch_bounds__(128) dg_fused_kernel(const __grid_constant__ DgParams P)
{
    for (int i = P.start + blockIdx.x * blockDim.x + threadIdx.x; i < P.end; i += gridDim.x * blockDim.x) {
        const int n = P.subset ? P.subset[i] : i;
        const int4 dg = *reinterpret_cast<const int4 *>(P.dgmap + 4 * (long long)n);
        const int4 cg = *reinterpret_cast<const int4 *>(P.cgmap + 4 * (long long)n);
        const int dgi[4] = {dg.x, dg.y, dg.z, dg.w}, cgi[4] = {cg.x, cg.y, cg.z, cg.w};
        Q1Cell K;
        load_cell(P, cgi, K);
        double ql[4], A[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; k++) ql[k] = P.q[dgi[k]];
        // ---- cell integral
        for (int qx = 0; qx < P.nq; qx++)
            for (int qy = 0; qy < P.nq; qy++) {
                const double x = P.xq[qx], y = P.xq[qy];
                double J[2][2];
                jac(K.c, x, y, J);
                const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
                const double id = 1.0 / det;
                const double Ki[2][2] = {{J[1][1] * id, -J[0][1] * id}, {-J[1][0] * id, J[0][0] * id}};
                const double w = fabs(det) * P.wq[qx] * P.wq[qy];
                double uv[2];
                p1(K.u, x, y, uv);
                const double bx[2] = {1.0 - x, x}, by[2] = {1.0 - y, y};
                double divu = 0.0, qv = 0.0;
#pragma unroll
                for (int ax = 0; ax < 2; ax++)
#pragma unroll
                    for (int ay = 0; ay < 2; ay++) {
                        const double g0 = (ax ? 1.0 : -1.0) * by[ay], g1 = bx[ax] * (ay ? 1.0 : -1.0);
                        divu += K.u[(ax * 2 + ay) * 2] * (Ki[0][0] * g0 + Ki[1][0] * g1)
                              + K.u[(ax * 2 + ay) * 2 + 1] * (Ki[0][1] * g0 + Ki[1][1] * g1);
                        qv += ql[ax * 2 + ay] * dq(P, ax, x) * dq(P, ay, y);
                    }
#pragma unroll
                for (int ax = 0; ax < 2; ax++)
#pragma unroll
                    for (int ay = 0; ay < 2; ay++) {
                        const double ph = dq(P, ax, x) * dq(P, ay, y);
                        const double g0 = ddq(P, ax) * dq(P, ay, y), g1 = dq(P, ax, x) * ddq(P, ay);
                        const double gp0 = Ki[0][0] * g0 + Ki[1][0] * g1, gp1 = Ki[0][1] * g0 + Ki[1][1] * g1;
                        A[ax * 2 + ay] += P.dt * w * qv * (gp0 * uv[0] + gp1 * uv[1] + ph * divu);
                    }
            }
        // ---- the four facets
        const int4 nb4 = *reinterpret_cast<const int4 *>(P.nbr + 4 * (long long)n);
        const uint4 nf4 = *reinterpret_cast<const uint4 *>(P.nbr_facet + 4 * (long long)n);
        const int nb[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
        const unsigned nf[4] = {nf4.x, nf4.y, nf4.z, nf4.w};
#pragma unroll
        for (int f = 0; f < 4; f++) {
            double qn[4] = {0, 0, 0, 0};
            if (nb[f] >= 0) {
                const int4 dn = *reinterpret_cast<const int4 *>(P.dgmap + 4 * (long long)nb[f]);
                qn[0] = P.q[dn.x]; qn[1] = P.q[dn.y]; qn[2] = P.q[dn.z]; qn[3] = P.q[dn.w];
            }
            for (int k = 0; k < P.nq; k++) {
                double x, y, nref[2], tref[2], nn[2], ds, uv[2];
                facet_point(f, P.xq[k], x, y, nref, tref);
                facet_geometry(K.c, x, y, nref, tref, nn, ds);
                p1(K.u, x, y, uv);
                const double udn = uv[0] * nn[0] + uv[1] * nn[1];
                double qv = 0.0;
#pragma unroll
                for (int ax = 0; ax < 2; ax++)
#pragma unroll
                    for (int ay = 0; ay < 2; ay++) qv += ql[ax * 2 + ay] * dq(P, ax, x) * dq(P, ay, y);
                double flux;
                if (nb[f] < 0) {
                    flux = (udn < 0.0 ? udn * P.q_in : 0.0) + (udn > 0.0 ? udn * qv : 0.0);
                } else {
                    double xn, yn, nr2[2], tr2[2];
                    facet_point((int)nf[f], P.xq[k], xn, yn, nr2, tr2);
                    double qnv = 0.0;
#pragma unroll
                    for (int ax = 0; ax < 2; ax++)
#pragma unroll
                        for (int ay = 0; ay < 2; ay++) qnv += qn[ax * 2 + ay] * dq(P, ax, xn) * dq(P, ay, yn);
                    const double un_me = 0.5 * (udn + fabs(udn)), un_nb = 0.5 * (-udn + fabs(udn));
                    flux = un_me * qv - un_nb * qnv;
                }
#pragma unroll
                for (int ax = 0; ax < 2; ax++)
#pragma unroll
                    for (int ay = 0; ay < 2; ay++)
                        A[ax * 2 + ay] -= P.dt * ds * P.wq[k] * dq(P, ax, x) * dq(P, ay, y) * flux;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) P.out[dgi[k]] += A[k];
    }
}

}  // namespace

int fdb_launch_dg_advection(fdb_kernel_s *k, fdb_int start, fdb_int end, const fdb_int *subset,
                            double *out, const double *coords, const double *q, const double *u,
                            const double *consts_host, const unsigned *facet, const fdb_int *dgmap,
                            const fdb_int *cgmap, const fdb_int *nbr)
{
    fdb::Context &c = fdb::ctx();
    if (end <= start) return 0;
    DgParams P;
    P.out = out;
    P.coords = coords;
    P.q = q;
    P.u = u;
    P.dgmap = dgmap;
    P.cgmap = cgmap;
    P.facet = facet;
    P.nbr = nbr;
    P.nbr_facet = facet;   // fused kernel: the "facet" argument is the (ncells, 4) neighbour-facet table
    P.start = start;
    P.end = end;
    P.subset = subset;
    P.nq = k->desc.nq;
    P.dt = consts_host[0];
    P.q_in = consts_host[1];
    for (int i = 0; i < 4; i++) P.Bend[i] = k->desc.B[i];
    for (int i = 0; i < FDB_MAX_1D; i++) {
        P.wq[i] = k->desc.wq[i];
        P.xq[i] = k->desc.xq[i];
    }
    long long blocks = ((long long)(end - start) + 127) / 128;
    long long cap = (long long)c.sm_count * 16;
    if (blocks > cap) blocks = cap;
    switch (k->desc.integral) {
    case FDB_INTEGRAL_CELL: dg_cell_kernel<<<(int)blocks, 128, 0, c.stream>>>(P); break;
    case FDB_INTEGRAL_EXTERIOR_FACET: dg_exterior_kernel<<<(int)blocks, 128, 0, c.stream>>>(P); break;
    case FDB_INTEGRAL_INTERIOR_FACET: dg_interior_kernel<<<(int)blocks, 128, 0, c.stream>>>(P); break;
    case FDB_INTEGRAL_FUSED: dg_fused_kernel<<<(int)blocks, 128, 0, c.stream>>>(P); break;
    default: fdb::set_error("dg advection: bad integral type"); return 1;
    }
    FDB_LAUNCH_CHECK();
    return 0;
}
