// Low-order fast path: action of  alpha*inner(grad u, grad v)*dx + beta*inner(u, v)*dx  on
// Q1 (x) P1 hexahedra (degree 1, 2^3 Gauss points), ONE THREAD PER CELL.
//
// The slab-thread kernel of action_hex.cu spends most of its instructions on the machinery that
// pays off at p >= 3 (shared-memory re-orientation of slabs, per-warp staging pipeline): at p = 1 a
// cell is 8 values and 8 quadrature points, which fit in the registers of one thread.  Here a warp
// takes 32 CONSECUTIVE LAYERS of one column: the gathers of a dof column are unit-stride across the
// lanes (coalesced 256-byte requests), nothing goes through shared memory, and the contributions
// of vertically adjacent cells to the dofs they share are combined with one shuffle before the
// scatter, which halves the RED.ADD.F64 count.  Same arithmetic as the reference's generated
// kernel (tsfc/kernel_interface/common.py:139-239; geometry at every point, tsfc/ufl_utils.py:41-85)
// and the same wrapper semantics (pyop2/codegen/builder.py:80-128, 352-429).
//
// Work per cell: 8 + 9 sum-factorised 2x2 contractions (~290 FMAs) and 8 x ~60 geometry
// operations: ~770 fp64 instructions against 16*8/ (shared dofs) ~ 41 bytes of compulsory traffic,
// i.e. fp64-pipe bound on a B200 (0.7 ms floor for 256^3 cells), not HBM bound.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

struct Q1Params {
    double *y;
    const double *x;
    const double *coords;
    const int *map0;
    const int *map1;
    const int *collist;
    int off0[8], off1[8];
    int ncols, col0, nlay;
    int items_per_col;       // ceil(nlay / 32)
    double alpha, beta;
    double B[4], D[4];       // [q][a]
    double wq[2], xq[2];
};

__device__ __forceinline__ double rcp_nr(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    // one third-order step r (1 + e + e^2), e = 1 - x r (|e| <= 2^-20 -> 2^-60): 3 dependent FMAs, not 4
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// out[i][j][k] = sum_s M[i*2+s] in[s][j][k]   (AX = 0),  ... along y (AX = 1), z (AX = 2);
// T: use the transpose of M
template <int AX, bool T>
__device__ __forceinline__ void contract(const double *M, const double (&in)[2][2][2], double (&out)[2][2][2])
{
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                double s = 0.0;
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int o = AX == 0 ? i : (AX == 1 ? j : k);
                    const double m = T ? M[t * 2 + o] : M[o * 2 + t];
                    const double v = AX == 0 ? in[t][j][k] : (AX == 1 ? in[i][t][k] : in[i][j][t]);
                    s = fma(m, v, s);
                }
                out[i][j][k] = s;
            }
}

// AFFINE: the caller's promise (fdb_kernel_desc.affine_cells, checked by fdb_cells_are_affine) that
// every cell is a parallelepiped: c4..c7 vanish, the cofactor rows and det J are formed once per cell.
template <bool MASS, bool AFFINE>
__global__ void __launch_bounds__(128, MASS ? 2 : 3) q1_action_kernel(const __grid_constant__ Q1Params P)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nitems = (long long)P.ncols * P.items_per_col;
    for (long long item = warp0; item < nitems; item += nwarps) {
        const int ci = (int)(item / P.items_per_col);
        const int l0 = (int)(item - (long long)ci * P.items_per_col) * 32;
        const int col = P.collist ? __ldg(P.collist + ci) : P.col0 + ci;
        const int layer = l0 + lane;
        const bool valid = layer < P.nlay;
        const int lay = valid ? layer : P.nlay - 1;          // idle lanes recompute the top cell, scatter nothing
        // ---- gather (unit stride across the lanes for every dof column)
        int g[8];
        double u[2][2][2];
#pragma unroll
        for (int loc = 0; loc < 8; loc++) {
            g[loc] = __ldg(P.map0 + (long long)col * 8 + loc) + P.off0[loc] * lay;
            u[loc >> 2][(loc >> 1) & 1][loc & 1] = __ldg(P.x + g[loc]);
        }
        double X[8][3];
#pragma unroll
        for (int v = 0; v < 8; v++) {
            const long long gv = (long long)(__ldg(P.map1 + (long long)col * 8 + v) + P.off1[v] * lay) * 3;
#pragma unroll
            for (int a = 0; a < 3; a++) X[v][a] = __ldg(P.coords + gv + a);
        }
        // ---- trilinear geometry coefficients
        double c1[3], c2[3], c3[3], c4[3], c5[3], c6[3], c7[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            c1[a] = X[4][a] - X[0][a];
            c2[a] = X[2][a] - X[0][a];
            c3[a] = X[1][a] - X[0][a];
            c4[a] = X[6][a] - X[4][a] - X[2][a] + X[0][a];
            c5[a] = X[3][a] - X[2][a] - X[1][a] + X[0][a];
            c6[a] = X[5][a] - X[4][a] - X[1][a] + X[0][a];
            c7[a] = X[7][a] - X[6][a] - X[5][a] - X[3][a] + X[4][a] + X[2][a] + X[1][a] - X[0][a];
        }
        // ---- reference gradient (and value) at the 8 points, sum factorised
        double bx[2][2][2], dx[2][2][2], t0[2][2][2], gx[2][2][2], gy[2][2][2], gz[2][2][2], val[2][2][2];
        contract<0, false>(P.B, u, bx);
        contract<0, false>(P.D, u, dx);
        contract<1, false>(P.B, dx, t0);
        contract<2, false>(P.B, t0, gx);        // D B B
        contract<1, false>(P.D, bx, t0);
        contract<2, false>(P.B, t0, gy);        // B D B
        contract<1, false>(P.B, bx, t0);
        contract<2, false>(P.D, t0, gz);        // B B D
        if (MASS) contract<2, false>(P.B, t0, val);
        double R0[3], R1[3], R2[3], adet_c = 1.0, rdet_c = 1.0;
        if (AFFINE) {
            R0[0] = c2[1] * c3[2] - c2[2] * c3[1];
            R0[1] = c2[2] * c3[0] - c2[0] * c3[2];
            R0[2] = c2[0] * c3[1] - c2[1] * c3[0];
            R1[0] = c3[1] * c1[2] - c3[2] * c1[1];
            R1[1] = c3[2] * c1[0] - c3[0] * c1[2];
            R1[2] = c3[0] * c1[1] - c3[1] * c1[0];
            R2[0] = c1[1] * c2[2] - c1[2] * c2[1];
            R2[1] = c1[2] * c2[0] - c1[0] * c2[2];
            R2[2] = c1[0] * c2[1] - c1[1] * c2[0];
            adet_c = fabs(c1[0] * R0[0] + c1[1] * R0[1] + c1[2] * R0[2]);
            rdet_c = rcp_nr(adet_c);
        }
        // ---- fluxes at the points (in place)
#pragma unroll
        for (int qx = 0; qx < 2; qx++)
#pragma unroll
            for (int qy = 0; qy < 2; qy++)
#pragma unroll
                for (int qz = 0; qz < 2; qz++) {
                    const double xi = P.xq[qx], eta = P.xq[qy], zeta = P.xq[qz];
                    double r0[3], r1[3], r2[3], adet, rdet;
                    if (AFFINE) {
#pragma unroll
                        for (int a = 0; a < 3; a++) { r0[a] = R0[a]; r1[a] = R1[a]; r2[a] = R2[a]; }
                        adet = adet_c;
                        rdet = rdet_c;
                    } else {
                        double ja[3], jb[3], jc[3];
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            ja[a] = fma(fma(c7[a], eta, c6[a]), zeta, fma(c4[a], eta, c1[a]));
                            jb[a] = fma(fma(c7[a], zeta, c4[a]), xi, fma(c5[a], zeta, c2[a]));
                            jc[a] = fma(fma(c7[a], eta, c6[a]), xi, fma(c5[a], eta, c3[a]));
                        }
                        r0[0] = jb[1] * jc[2] - jb[2] * jc[1];
                        r0[1] = jb[2] * jc[0] - jb[0] * jc[2];
                        r0[2] = jb[0] * jc[1] - jb[1] * jc[0];
                        r1[0] = jc[1] * ja[2] - jc[2] * ja[1];
                        r1[1] = jc[2] * ja[0] - jc[0] * ja[2];
                        r1[2] = jc[0] * ja[1] - jc[1] * ja[0];
                        r2[0] = ja[1] * jb[2] - ja[2] * jb[1];
                        r2[1] = ja[2] * jb[0] - ja[0] * jb[2];
                        r2[2] = ja[0] * jb[1] - ja[1] * jb[0];
                        adet = fabs(ja[0] * r0[0] + ja[1] * r0[1] + ja[2] * r0[2]);
                        rdet = rcp_nr(adet);
                    }
                    const double w = P.wq[qx] * P.wq[qy] * P.wq[qz];
                    const double s = P.alpha * w * rdet;
                    const double a0 = gx[qx][qy][qz], a1 = gy[qx][qy][qz], a2 = gz[qx][qy][qz];
                    double h[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) h[a] = r0[a] * a0 + r1[a] * a1 + r2[a] * a2;
                    gx[qx][qy][qz] = s * (r0[0] * h[0] + r0[1] * h[1] + r0[2] * h[2]);
                    gy[qx][qy][qz] = s * (r1[0] * h[0] + r1[1] * h[1] + r1[2] * h[2]);
                    gz[qx][qy][qz] = s * (r2[0] * h[0] + r2[1] * h[1] + r2[2] * h[2]);
                    if (MASS) val[qx][qy][qz] *= P.beta * w * adet;
                }
        // ---- transpose path: r = D^T B^T B^T fx + B^T D^T B^T fy + B^T B^T D^T fz (+ B^T B^T B^T m)
        double r[2][2][2];
        contract<2, true>(P.B, gx, t0);
        contract<1, true>(P.B, t0, bx);
        contract<0, true>(P.D, bx, r);
        contract<2, true>(P.B, gy, t0);
        contract<1, true>(P.D, t0, bx);
        contract<2, true>(P.D, gz, t0);
        contract<1, true>(P.B, t0, dx);
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int k = 0; k < 2; k++) bx[i][j][k] += dx[i][j][k];
        if (MASS) {
            contract<2, true>(P.B, val, t0);
            contract<1, true>(P.B, t0, dx);
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int k = 0; k < 2; k++) bx[i][j][k] += dx[i][j][k];
        }
        contract<0, true>(P.B, bx, t0);
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int k = 0; k < 2; k++) r[i][j][k] += t0[i][j][k];
        // ---- scatter-add: the cell above (lane + 1) shares my top dofs (its bottom ones) when the
        //      numbering is extruded; its contribution rides along and it skips those atomics
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int gb = g[b * 2], gt = g[b * 2 + 1];
            double vb = r[b >> 1][b & 1][0], vt = r[b >> 1][b & 1][1];
            const double nb = __shfl_down_sync(0xffffffffu, vb, 1);
            const int gnb = __shfl_down_sync(0xffffffffu, gb, 1);
            const int nvalid = __shfl_down_sync(0xffffffffu, (int)valid, 1);
            const int gpt = __shfl_up_sync(0xffffffffu, gt, 1);
            const bool take = lane < 31 && nvalid && gnb == gt;          // I add my upper neighbour's share
            const bool taken = lane > 0 && gpt == gb;                    // my lower neighbour adds mine
            if (take) vt += nb;
            if (valid) {
                atomicAdd(P.y + gt, vt);
                if (!taken) atomicAdd(P.y + gb, vb);
            }
        }
    }
}

}  // namespace

int fdb_launch_q1_action(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay, const fdb_int *subset,
                         double *y, const double *coords, const double *x, const fdb_int *map0,
                         const fdb_int *map1)
{
    fdb::Context &c = fdb::ctx();
    Q1Params P;
    memset(&P, 0, sizeof(P));
    P.y = y;
    P.x = x;
    P.coords = coords;
    P.map0 = map0;
    P.map1 = map1;
    P.collist = subset;
    const bool extruded = k->desc.cell == FDB_CELL_HEX_EXTRUDED;
    for (int i = 0; i < 8; i++) {
        P.off0[i] = extruded ? k->h_off0[i] : 0;
        P.off1[i] = extruded ? k->h_off1[i] : 0;
    }
    P.ncols = end - start;
    P.col0 = start;
    P.nlay = nlay;
    P.items_per_col = (nlay + 31) / 32;
    P.alpha = k->desc.alpha;
    P.beta = k->desc.beta;
    for (int i = 0; i < 4; i++) {
        P.B[i] = k->desc.B[i];
        P.D[i] = k->desc.D[i];
    }
    for (int i = 0; i < 2; i++) {
        P.wq[i] = k->desc.wq[i];
        P.xq[i] = k->desc.xq[i];
    }
    if (P.ncols <= 0 || nlay <= 0) return 0;
    const long long nitems = (long long)P.ncols * P.items_per_col;
    long long grid = (nitems + 3) / 4;                       // 4 warps per CTA
    const long long cap = (long long)c.sm_count * 12;        // a few waves of resident CTAs
    if (grid > cap) grid = cap;
    const bool mass = k->desc.beta != 0.0, aff = k->desc.affine_cells != 0;
    if (mass && aff) q1_action_kernel<true, true><<<(int)grid, 128, 0, c.stream>>>(P);
    else if (mass) q1_action_kernel<true, false><<<(int)grid, 128, 0, c.stream>>>(P);
    else if (aff) q1_action_kernel<false, true><<<(int)grid, 128, 0, c.stream>>>(P);
    else q1_action_kernel<false, false><<<(int)grid, 128, 0, c.stream>>>(P);
    FDB_LAUNCH_CHECK();
    return 0;
}
