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
#include <stdlib.h>

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

// ---- round 2: the fused kernel specialised on the number of quadrature points -----------------
// Same integrals as dg_fused_kernel, reorganised so that the straight-line code per cell drops from
// ~2300 to ~900 fp64 instructions (SASS count in profiles/r02_dg_advection.txt):
//   * NQ is a template parameter: every loop unrolls, the DQ1 tabulation phi_i(x_k) is 2*NQ registers;
//   * bilinear geometry and velocity are expanded once per cell (X = C0 + C1 x + C2 y + C3 xy);
//   * |det J| w K^{-T} grad = sign(det) w adj(J)^T grad: the cell term needs NO division;
//   * the scaled facet normal cof(J) n_ref has length ds and is CONSTANT along a straight edge:
//     u.n ds = u.(cof(J) n_ref), so the facet terms need no sqrt / rsqrt and no per-point geometry;
//   * the neighbour's trace is reduced to its two edge coefficients before the point loop.
// Results agree with the generic kernel to rounding (tests/test_dg_advection_gpu.py, 1e-12).
template <int NQ>
__global__ void __launch_bounds__(128) dg_fused_nq_kernel(const __grid_constant__ DgParams P)
{
    double Phi[2][NQ];                       // phi_i(x_k)
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int k = 0; k < NQ; k++) Phi[i][k] = P.Bend[i] * (1.0 - P.xq[k]) + P.Bend[2 + i] * P.xq[k];
    const double dPhi[2] = {P.Bend[2] - P.Bend[0], P.Bend[3] - P.Bend[1]};
    for (int i = P.start + blockIdx.x * blockDim.x + threadIdx.x; i < P.end; i += gridDim.x * blockDim.x) {
        const int n = P.subset ? P.subset[i] : i;
        const int4 dg = *reinterpret_cast<const int4 *>(P.dgmap + 4 * (long long)n);
        const int4 cg = *reinterpret_cast<const int4 *>(P.cgmap + 4 * (long long)n);
        const int4 nb4 = *reinterpret_cast<const int4 *>(P.nbr + 4 * (long long)n);
        const uint4 nf4 = *reinterpret_cast<const uint4 *>(P.nbr_facet + 4 * (long long)n);
        const int dgi[4] = {dg.x, dg.y, dg.z, dg.w}, cgi[4] = {cg.x, cg.y, cg.z, cg.w};
        const int nb[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
        const unsigned nf[4] = {nf4.x, nf4.y, nf4.z, nf4.w};
        // neighbour rows first: the longest dependent chain (map row -> q values)
        double qn[4][4];
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const int4 dn = *reinterpret_cast<const int4 *>(P.dgmap + 4 * (long long)(nb[f] >= 0 ? nb[f] : n));
            qn[f][0] = P.q[dn.x]; qn[f][1] = P.q[dn.y]; qn[f][2] = P.q[dn.z]; qn[f][3] = P.q[dn.w];
        }
        Q1Cell K;
        load_cell(P, cgi, K);
        double ql[2][2], A[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
        for (int k = 0; k < 4; k++) ql[k >> 1][k & 1] = P.q[dgi[k]];
        // bilinear expansions  f(x, y) = F0 + F1 x + F2 y + F3 x y  (component a)
        double C1[2], C2[2], C3[2], U0[2], U1[2], U2[2], U3[2];
#pragma unroll
        for (int a = 0; a < 2; a++) {
            C1[a] = K.c[4 + a] - K.c[a];
            C2[a] = K.c[2 + a] - K.c[a];
            C3[a] = K.c[6 + a] - K.c[4 + a] - K.c[2 + a] + K.c[a];
            U0[a] = K.u[a];
            U1[a] = K.u[4 + a] - K.u[a];
            U2[a] = K.u[2 + a] - K.u[a];
            U3[a] = K.u[6 + a] - K.u[4 + a] - K.u[2 + a] + K.u[a];
        }
        // orientation of the cell (constant sign of det J on a valid cell): at the centre
        const double detc = (C1[0] + 0.5 * C3[0]) * (C2[1] + 0.5 * C3[1]) - (C2[0] + 0.5 * C3[0]) * (C1[1] + 0.5 * C3[1]);
        const double sg = detc < 0.0 ? -1.0 : 1.0;
        // ---- cell integral
#pragma unroll
        for (int qx = 0; qx < NQ; qx++) {
            const double x = P.xq[qx];
            const double J01 = fma(C3[0], x, C2[0]), J11 = fma(C3[1], x, C2[1]);      // dX/dy
            const double uy0 = fma(U3[0], x, U2[0]), uy1 = fma(U3[1], x, U2[1]);      // du/dy (reference)
            const double ub0 = fma(U1[0], x, U0[0]), ub1 = fma(U1[1], x, U0[1]);      // u = ub + uy * y
            const double qa = ql[0][0] * Phi[0][qx] + ql[1][0] * Phi[1][qx];           // q = qa phi_0(y) + qb phi_1(y)
            const double qb = ql[0][1] * Phi[0][qx] + ql[1][1] * Phi[1][qx];
            double T0[2] = {0, 0}, T1[2] = {0, 0};       // sums over qy of phi_ay(y) * (.) and dphi_ay * (.)
#pragma unroll
            for (int qy = 0; qy < NQ; qy++) {
                const double y = P.xq[qy];
                const double J00 = fma(C3[0], y, C1[0]), J10 = fma(C3[1], y, C1[1]);  // dX/dx
                const double ux0 = fma(U3[0], y, U1[0]), ux1 = fma(U3[1], y, U1[1]);  // du/dx (reference)
                const double uv0 = fma(uy0, y, ub0), uv1 = fma(uy1, y, ub1);
                const double qv = qa * Phi[0][qy] + qb * Phi[1][qy];
                const double sc = P.dt * sg * P.wq[qx] * P.wq[qy] * qv;
                const double a0 = J11 * uv0 - J01 * uv1;                 // coefficient of d/dx(phi)
                const double a1 = J00 * uv1 - J10 * uv0;                 // coefficient of d/dy(phi)
                const double dd = J11 * ux0 - J10 * uy0 - J01 * ux1 + J00 * uy1;   // det * div u
                // A[ax][ay] += sc * (a0 dphi_ax phi_ay + a1 phi_ax dphi_ay + dd phi_ax phi_ay)
#pragma unroll
                for (int ay = 0; ay < 2; ay++) {
                    T0[ay] = fma(sc * a0, Phi[ay][qy], T0[ay]);
                    T1[ay] = fma(sc, fma(a1, dPhi[ay], dd * Phi[ay][qy]), T1[ay]);
                }
            }
#pragma unroll
            for (int ax = 0; ax < 2; ax++)
#pragma unroll
                for (int ay = 0; ay < 2; ay++) A[ax][ay] += dPhi[ax] * T0[ay] + Phi[ax][qx] * T1[ay];
        }
        // ---- the four facets (f = 0: x = 0, 1: x = 1, 2: y = 0, 3: y = 1)
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const int e = f & 1;
            const double xe = (double)e, sgn = e ? 1.0 : -1.0;
            // scaled outward normal (constant along the edge), |nn| = ds
            double nn0, nn1;
            if (f < 2) {        // n_ref = (sgn, 0): cof(J) n_ref = sgn (J11, -J01) at x = xe
                nn0 = sg * sgn * fma(C3[1], xe, C2[1]);
                nn1 = -sg * sgn * fma(C3[0], xe, C2[0]);
            } else {            // n_ref = (0, sgn): cof(J) n_ref = sgn (-J10, J00) at y = xe
                nn0 = -sg * sgn * fma(C3[1], xe, C1[1]);
                nn1 = sg * sgn * fma(C3[0], xe, C1[0]);
            }
            // my trace: q(s) = ta phi_0(s) + tb phi_1(s)
            double ta, tb;
            if (f < 2) {
                ta = ql[0][0] * P.Bend[e * 2] + ql[1][0] * P.Bend[e * 2 + 1];
                tb = ql[0][1] * P.Bend[e * 2] + ql[1][1] * P.Bend[e * 2 + 1];
            } else {
                ta = ql[0][0] * P.Bend[e * 2] + ql[0][1] * P.Bend[e * 2 + 1];
                tb = ql[1][0] * P.Bend[e * 2] + ql[1][1] * P.Bend[e * 2 + 1];
            }
            // the neighbour's trace on ITS local facet nf
            const bool interior = nb[f] >= 0;
            const int en = (int)(nf[f] & 1u);
            const double b0 = P.Bend[en * 2], b1 = P.Bend[en * 2 + 1];
            const bool nvert = nf[f] < 2u;
            const double na = nvert ? qn[f][0] * b0 + qn[f][2] * b1 : qn[f][0] * b0 + qn[f][1] * b1;
            const double nbv = nvert ? qn[f][1] * b0 + qn[f][3] * b1 : qn[f][2] * b0 + qn[f][3] * b1;
            // velocity along the edge: u(s) = ue + us * s
            double ue0, ue1, us0, us1;
            if (f < 2) {
                ue0 = fma(U1[0], xe, U0[0]); ue1 = fma(U1[1], xe, U0[1]);
                us0 = fma(U3[0], xe, U2[0]); us1 = fma(U3[1], xe, U2[1]);
            } else {
                ue0 = fma(U2[0], xe, U0[0]); ue1 = fma(U2[1], xe, U0[1]);
                us0 = fma(U3[0], xe, U1[0]); us1 = fma(U3[1], xe, U1[1]);
            }
            const double un_e = ue0 * nn0 + ue1 * nn1, un_s = us0 * nn0 + us1 * nn1;   // u.n ds = un_e + un_s s
            double E[2] = {0, 0};                   // sum_k w_k phi_j(s_k) flux_k
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                const double uds = fma(un_s, P.xq[k], un_e);
                const double qv = ta * Phi[0][k] + tb * Phi[1][k];
                double flux;
                if (!interior) {
                    flux = (uds < 0.0 ? uds * P.q_in : 0.0) + (uds > 0.0 ? uds * qv : 0.0);
                } else {
                    const double qnv = na * Phi[0][k] + nbv * Phi[1][k];
                    flux = fmax(uds, 0.0) * qv - fmax(-uds, 0.0) * qnv;
                }
                const double wf = P.dt * P.wq[k] * flux;
                E[0] = fma(wf, Phi[0][k], E[0]);
                E[1] = fma(wf, Phi[1][k], E[1]);
            }
            if (f < 2) {
#pragma unroll
                for (int ax = 0; ax < 2; ax++)
#pragma unroll
                    for (int ay = 0; ay < 2; ay++) A[ax][ay] -= P.Bend[e * 2 + ax] * E[ay];
            } else {
#pragma unroll
                for (int ax = 0; ax < 2; ax++)
#pragma unroll
                    for (int ay = 0; ay < 2; ay++) A[ax][ay] -= P.Bend[e * 2 + ay] * E[ax];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) P.out[dgi[k]] += A[k >> 1][k & 1];
    }
}

// ---- round 2: the three reference-ABI kernels specialised on NQ with the same algebra as
// dg_fused_nq_kernel (no division in the cell term, scaled facet normal = cof(J) n_ref, constant
// along a straight edge: no sqrt / rsqrt).  These are what a real Firedrake calls through the
// hook (one parloop per integral type, firedrake/assemble.py:1069-1096).
struct DgCellX {            // bilinear expansions f = F0 + F1 x + F2 y + F3 xy of geometry and velocity
    double C1[2], C2[2], C3[2], U0[2], U1[2], U2[2], U3[2], sg;
};

__device__ __forceinline__ void dg_expand(const Q1Cell &K, DgCellX &G)
{
#pragma unroll
    for (int a = 0; a < 2; a++) {
        G.C1[a] = K.c[4 + a] - K.c[a];
        G.C2[a] = K.c[2 + a] - K.c[a];
        G.C3[a] = K.c[6 + a] - K.c[4 + a] - K.c[2 + a] + K.c[a];
        G.U0[a] = K.u[a];
        G.U1[a] = K.u[4 + a] - K.u[a];
        G.U2[a] = K.u[2 + a] - K.u[a];
        G.U3[a] = K.u[6 + a] - K.u[4 + a] - K.u[2 + a] + K.u[a];
    }
    const double detc = (G.C1[0] + 0.5 * G.C3[0]) * (G.C2[1] + 0.5 * G.C3[1]) -
                        (G.C2[0] + 0.5 * G.C3[0]) * (G.C1[1] + 0.5 * G.C3[1]);
    G.sg = detc < 0.0 ? -1.0 : 1.0;
}

// E[j] = sum_k dt w_k phi_j(s_k) flux_k on local facet f (run-time) of the cell G; q(s) = ta phi_0 + tb phi_1
// is the cell's own trace, (na, nb) the neighbour's (interior) -- flux as in dg_fused_nq_kernel
template <int NQ>
__device__ __forceinline__ void dg_facet_E(const DgParams &P, const double (&Phi)[2][NQ], const DgCellX &G, int f,
                                           double ta, double tb, bool interior, double na, double nbv,
                                           double (&E)[2])
{
    const int e = f & 1;
    const double xe = (double)e, sgn = e ? 1.0 : -1.0;
    const bool vert = f < 2;
    const double nn0 = vert ? G.sg * sgn * fma(G.C3[1], xe, G.C2[1]) : -G.sg * sgn * fma(G.C3[1], xe, G.C1[1]);
    const double nn1 = vert ? -G.sg * sgn * fma(G.C3[0], xe, G.C2[0]) : G.sg * sgn * fma(G.C3[0], xe, G.C1[0]);
    const double ue0 = vert ? fma(G.U1[0], xe, G.U0[0]) : fma(G.U2[0], xe, G.U0[0]);
    const double ue1 = vert ? fma(G.U1[1], xe, G.U0[1]) : fma(G.U2[1], xe, G.U0[1]);
    const double us0 = vert ? fma(G.U3[0], xe, G.U2[0]) : fma(G.U3[0], xe, G.U1[0]);
    const double us1 = vert ? fma(G.U3[1], xe, G.U2[1]) : fma(G.U3[1], xe, G.U1[1]);
    const double un_e = ue0 * nn0 + ue1 * nn1, un_s = us0 * nn0 + us1 * nn1;
    E[0] = E[1] = 0.0;
#pragma unroll
    for (int k = 0; k < NQ; k++) {
        const double uds = fma(un_s, P.xq[k], un_e);
        const double qv = ta * Phi[0][k] + tb * Phi[1][k];
        double flux;
        if (!interior) {
            flux = (uds < 0.0 ? uds * P.q_in : 0.0) + (uds > 0.0 ? uds * qv : 0.0);
        } else {
            const double qnv = na * Phi[0][k] + nbv * Phi[1][k];
            flux = fmax(uds, 0.0) * qv - fmax(-uds, 0.0) * qnv;
        }
        const double wf = P.dt * P.wq[k] * flux;
        E[0] = fma(wf, Phi[0][k], E[0]);
        E[1] = fma(wf, Phi[1][k], E[1]);
    }
}

// the two edge coefficients of q (4 dofs, index ax*2+ay) on local facet f
__device__ __forceinline__ void dg_trace(const DgParams &P, const double *q, int f, double &ta, double &tb)
{
    const int e = f & 1;
    const double b0 = P.Bend[e * 2], b1 = P.Bend[e * 2 + 1];
    if (f < 2) { ta = q[0] * b0 + q[2] * b1; tb = q[1] * b0 + q[3] * b1; }
    else       { ta = q[0] * b0 + q[1] * b1; tb = q[2] * b0 + q[3] * b1; }
}

// A[ax*2+ay] += sign * phi_(normal index)(x_e) * E[(tangential index)]
__device__ __forceinline__ void dg_facet_add(const DgParams &P, int f, const double (&E)[2], double sign, double *A)
{
    const int e = f & 1;
#pragma unroll
    for (int ax = 0; ax < 2; ax++)
#pragma unroll
        for (int ay = 0; ay < 2; ay++)
            A[ax * 2 + ay] += sign * (f < 2 ? P.Bend[e * 2 + ax] * E[ay] : P.Bend[e * 2 + ay] * E[ax]);
}

template <int NQ>
__device__ __forceinline__ void dg_tab(const DgParams &P, double (&Phi)[2][NQ])
{
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int k = 0; k < NQ; k++) Phi[i][k] = P.Bend[i] * (1.0 - P.xq[k]) + P.Bend[2 + i] * P.xq[k];
}

template <int NQ>
__global__ void __launch_bounds__(128) dg_cell_nq_kernel(const __grid_constant__ DgParams P)
{
    double Phi[2][NQ];
    dg_tab<NQ>(P, Phi);
    const double dPhi[2] = {P.Bend[2] - P.Bend[0], P.Bend[3] - P.Bend[1]};
    for (int i = P.start + blockIdx.x * blockDim.x + threadIdx.x; i < P.end; i += gridDim.x * blockDim.x) {
        const int n = P.subset ? P.subset[i] : i;
        const int4 dg = *reinterpret_cast<const int4 *>(P.dgmap + 4 * (long long)n);
        const int4 cg = *reinterpret_cast<const int4 *>(P.cgmap + 4 * (long long)n);
        const int dgi[4] = {dg.x, dg.y, dg.z, dg.w}, cgi[4] = {cg.x, cg.y, cg.z, cg.w};
        Q1Cell K;
        load_cell(P, cgi, K);
        double ql[2][2], A[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
        for (int k = 0; k < 4; k++) ql[k >> 1][k & 1] = P.q[dgi[k]];
        DgCellX G;
        dg_expand(K, G);
#pragma unroll
        for (int qx = 0; qx < NQ; qx++) {
            const double x = P.xq[qx];
            const double J01 = fma(G.C3[0], x, G.C2[0]), J11 = fma(G.C3[1], x, G.C2[1]);
            const double uy0 = fma(G.U3[0], x, G.U2[0]), uy1 = fma(G.U3[1], x, G.U2[1]);
            const double ub0 = fma(G.U1[0], x, G.U0[0]), ub1 = fma(G.U1[1], x, G.U0[1]);
            const double qa = ql[0][0] * Phi[0][qx] + ql[1][0] * Phi[1][qx];
            const double qb = ql[0][1] * Phi[0][qx] + ql[1][1] * Phi[1][qx];
            double T0[2] = {0, 0}, T1[2] = {0, 0};
#pragma unroll
            for (int qy = 0; qy < NQ; qy++) {
                const double y = P.xq[qy];
                const double J00 = fma(G.C3[0], y, G.C1[0]), J10 = fma(G.C3[1], y, G.C1[1]);
                const double ux0 = fma(G.U3[0], y, G.U1[0]), ux1 = fma(G.U3[1], y, G.U1[1]);
                const double uv0 = fma(uy0, y, ub0), uv1 = fma(uy1, y, ub1);
                const double qv = qa * Phi[0][qy] + qb * Phi[1][qy];
                const double sc = P.dt * G.sg * P.wq[qx] * P.wq[qy] * qv;
                const double a0 = J11 * uv0 - J01 * uv1;
                const double a1 = J00 * uv1 - J10 * uv0;
                const double dd = J11 * ux0 - J10 * uy0 - J01 * ux1 + J00 * uy1;
#pragma unroll
                for (int ay = 0; ay < 2; ay++) {
                    T0[ay] = fma(sc * a0, Phi[ay][qy], T0[ay]);
                    T1[ay] = fma(sc, fma(a1, dPhi[ay], dd * Phi[ay][qy]), T1[ay]);
                }
            }
#pragma unroll
            for (int ax = 0; ax < 2; ax++)
#pragma unroll
                for (int ay = 0; ay < 2; ay++) A[ax][ay] += dPhi[ax] * T0[ay] + Phi[ax][qx] * T1[ay];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) atomicAdd(P.out + dgi[k], A[k >> 1][k & 1]);
    }
}

template <int NQ>
__global__ void __launch_bounds__(128) dg_exterior_nq_kernel(const __grid_constant__ DgParams P)
{
    double Phi[2][NQ];
    dg_tab<NQ>(P, Phi);
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
        DgCellX G;
        dg_expand(K, G);
        double ta, tb, E[2];
        dg_trace(P, ql, lf, ta, tb);
        dg_facet_E<NQ>(P, Phi, G, lf, ta, tb, false, 0.0, 0.0, E);
        dg_facet_add(P, lf, E, -1.0, A);
#pragma unroll
        for (int k = 0; k < 4; k++) atomicAdd(P.out + dgi[k], A[k]);
    }
}

template <int NQ>
__global__ void __launch_bounds__(128) dg_interior_nq_kernel(const __grid_constant__ DgParams P)
{
    double Phi[2][NQ];
    dg_tab<NQ>(P, Phi);
    for (int i = P.start + blockIdx.x * blockDim.x + threadIdx.x; i < P.end; i += gridDim.x * blockDim.x) {
        const int f = P.subset ? P.subset[i] : i;
        int dgi[8], cgi[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int4 a = *reinterpret_cast<const int4 *>(P.dgmap + 8 * (long long)f + 4 * h);
            dgi[4 * h] = a.x; dgi[4 * h + 1] = a.y; dgi[4 * h + 2] = a.z; dgi[4 * h + 3] = a.w;
        }
        {
            // geometry and velocity of the '+' cell: u is continuous and the mesh conforming, so the '-'
            // side's u.n is minus this one (what the reference's facet kernel computes from cell '-')
            const int4 b = *reinterpret_cast<const int4 *>(P.cgmap + 8 * (long long)f);
            cgi[0] = b.x; cgi[1] = b.y; cgi[2] = b.z; cgi[3] = b.w;
        }
        Q1Cell Kp;
        load_cell(P, cgi, Kp);
        double ql[8], A[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 8; k++) ql[k] = P.q[dgi[k]];
        const int fp = (int)P.facet[2 * (long long)f], fm = (int)P.facet[2 * (long long)f + 1];
        DgCellX G;
        dg_expand(Kp, G);
        double ta, tb, na, nbv, E[2];
        dg_trace(P, ql, fp, ta, tb);
        dg_trace(P, ql + 4, fm, na, nbv);
        dg_facet_E<NQ>(P, Phi, G, fp, ta, tb, true, na, nbv, E);
        dg_facet_add(P, fp, E, -1.0, A);
        dg_facet_add(P, fm, E, 1.0, A + 4);
#pragma unroll
        for (int k = 0; k < 8; k++) atomicAdd(P.out + dgi[k], A[k]);
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
    static const bool dg_generic = getenv("FDB_DG_GENERIC") && atoi(getenv("FDB_DG_GENERIC"));
    switch (k->desc.integral) {
    case FDB_INTEGRAL_CELL:
        if (!dg_generic && P.nq == 2) dg_cell_nq_kernel<2><<<(int)blocks, 128, 0, c.stream>>>(P);
        else if (!dg_generic && P.nq == 3) dg_cell_nq_kernel<3><<<(int)blocks, 128, 0, c.stream>>>(P);
        else if (!dg_generic && P.nq == 4) dg_cell_nq_kernel<4><<<(int)blocks, 128, 0, c.stream>>>(P);
        else dg_cell_kernel<<<(int)blocks, 128, 0, c.stream>>>(P);
        break;
    case FDB_INTEGRAL_EXTERIOR_FACET:
        if (!dg_generic && P.nq == 2) dg_exterior_nq_kernel<2><<<(int)blocks, 128, 0, c.stream>>>(P);
        else if (!dg_generic && P.nq == 3) dg_exterior_nq_kernel<3><<<(int)blocks, 128, 0, c.stream>>>(P);
        else if (!dg_generic && P.nq == 4) dg_exterior_nq_kernel<4><<<(int)blocks, 128, 0, c.stream>>>(P);
        else dg_exterior_kernel<<<(int)blocks, 128, 0, c.stream>>>(P);
        break;
    case FDB_INTEGRAL_INTERIOR_FACET:
        if (!dg_generic && P.nq == 2) dg_interior_nq_kernel<2><<<(int)blocks, 128, 0, c.stream>>>(P);
        else if (!dg_generic && P.nq == 3) dg_interior_nq_kernel<3><<<(int)blocks, 128, 0, c.stream>>>(P);
        else if (!dg_generic && P.nq == 4) dg_interior_nq_kernel<4><<<(int)blocks, 128, 0, c.stream>>>(P);
        else dg_interior_kernel<<<(int)blocks, 128, 0, c.stream>>>(P);
        break;
    case FDB_INTEGRAL_FUSED: {
        const bool generic = dg_generic;
        if (!generic && P.nq == 2) dg_fused_nq_kernel<2><<<(int)blocks, 128, 0, c.stream>>>(P);
        else if (!generic && P.nq == 3) dg_fused_nq_kernel<3><<<(int)blocks, 128, 0, c.stream>>>(P);
        else if (!generic && P.nq == 4) dg_fused_nq_kernel<4><<<(int)blocks, 128, 0, c.stream>>>(P);
        else dg_fused_kernel<<<(int)blocks, 128, 0, c.stream>>>(P);
        break;
    }
    default: fdb::set_error("dg advection: bad integral type"); return 1;
    }
    FDB_LAUNCH_CHECK();
    return 0;
}
