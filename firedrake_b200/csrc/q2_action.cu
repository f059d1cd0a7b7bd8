// Degree 2: action of  alpha*inner(grad u, grad v)*dx + beta*inner(u, v)*dx  on Q2 (x) P2
// hexahedra (27 dofs, 3^3 Gauss points), ONE THREAD PER CELL -- the degree-2 sibling of q1_action.cu.
//
// 27 values do not leave room for three separate gradient tensors in one thread's registers, so the
// arithmetic is the collocated one of the slab kernel: interpolate u to the Gauss points (three 3x3
// contractions, in place, every index static because all loops are unrolled), then at each point
// take the three derivatives straight from the collocated values with Dt = D B^{-1} (9 FMAs), apply
// the metric of the trilinear geometry (recomputed at the point, as TSFC does), and accumulate the
// transposed derivative into a second 27-value tensor; three transposed contractions bring it back
// to the dofs.  Two 27-double tensors + the 21 geometry coefficients live in registers (255, with
// ~85 doubles spilled by the fully unrolled point loop -- parking the collocated values in shared
// memory instead did not reduce the spills: they come from the scheduler's hoisting of the geometry
// across the 27 unrolled points); nothing goes through shared memory.  A warp = 32 consecutive layers of one column (strided-coalesced gathers),
// vertically adjacent cells merge the contributions to their shared face dofs with one shuffle.
//
// Work: 2 x 3 x 81 FMAs of contractions + 27 x (18 + ~60) per point = ~2600 fp64 instructions per
// cell -> 2.3 ms floor for 256^3 cells on a B200.  Reference semantics as in action_hex.cu.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int N = 3, ND = 27;

struct Q2Params {
    double *y;
    const double *x;
    const double *coords;
    const int *map0;
    const int *map1;
    const int *collist;
    int off0[ND], off1[8];
    int ncols, col0, nlay;
    int items_per_col;
    double alpha, beta;
    double B[N * N], Dt[N * N];      // [q][a], [q][q']
    double wq[N], xq[N];
};

__device__ __forceinline__ double rcp_nr2(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(e, r, r);
    e = fma(-x, r, 1.0);
    r = fma(e, r, r);
    return r;
}

// in-place contraction of t[N][N][N] along axis AX with M (T: its transpose), all indices static
template <int AX, bool T>
__device__ __forceinline__ void contract3(const double *M, double (&t)[N][N][N])
{
#pragma unroll
    for (int a = 0; a < N; a++)
#pragma unroll
        for (int b = 0; b < N; b++) {
            double in[N], out[N];
#pragma unroll
            for (int s = 0; s < N; s++) in[s] = AX == 0 ? t[s][a][b] : (AX == 1 ? t[a][s][b] : t[a][b][s]);
#pragma unroll
            for (int o = 0; o < N; o++) {
                double acc = 0.0;
#pragma unroll
                for (int s = 0; s < N; s++) acc = fma(T ? M[s * N + o] : M[o * N + s], in[s], acc);
                out[o] = acc;
            }
#pragma unroll
            for (int o = 0; o < N; o++) {
                if (AX == 0) t[o][a][b] = out[o];
                else if (AX == 1) t[a][o][b] = out[o];
                else t[a][b][o] = out[o];
            }
        }
}

template <bool MASS>
__global__ void __launch_bounds__(128, 2) q2_action_kernel(const __grid_constant__ Q2Params P)
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
        const int lay = valid ? layer : P.nlay - 1;
        // ---- gather
        double U[N][N][N];
#pragma unroll
        for (int loc = 0; loc < ND; loc++) {
            const int g = __ldg(P.map0 + (long long)col * ND + loc) + P.off0[loc] * lay;
            U[loc / 9][(loc / 3) % 3][loc % 3] = __ldg(P.x + g);
        }
        double c1[3], c2[3], c3[3], c4[3], c5[3], c6[3], c7[3];
        {
            double X[8][3];
#pragma unroll
            for (int v = 0; v < 8; v++) {
                const long long gv = (long long)(__ldg(P.map1 + (long long)col * 8 + v) + P.off1[v] * lay) * 3;
#pragma unroll
                for (int a = 0; a < 3; a++) X[v][a] = __ldg(P.coords + gv + a);
            }
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
        }
        // ---- to the Gauss points (collocated values)
        contract3<0, false>(P.B, U);
        contract3<1, false>(P.B, U);
        contract3<2, false>(P.B, U);
        double V[N][N][N];
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++)
#pragma unroll
                for (int k = 0; k < N; k++) V[i][j][k] = 0.0;
        // ---- quadrature points
#pragma unroll
        for (int qx = 0; qx < N; qx++)
#pragma unroll
            for (int qy = 0; qy < N; qy++)
#pragma unroll
                for (int qz = 0; qz < N; qz++) {
                    double gx = 0.0, gy = 0.0, gz = 0.0;
#pragma unroll
                    for (int s = 0; s < N; s++) {
                        gx = fma(P.Dt[qx * N + s], U[s][qy][qz], gx);
                        gy = fma(P.Dt[qy * N + s], U[qx][s][qz], gy);
                        gz = fma(P.Dt[qz * N + s], U[qx][qy][s], gz);
                    }
                    const double xi = P.xq[qx], eta = P.xq[qy], zeta = P.xq[qz];
                    double ja[3], jb[3], jc[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        ja[a] = fma(fma(c7[a], eta, c6[a]), zeta, fma(c4[a], eta, c1[a]));
                        jb[a] = fma(fma(c7[a], zeta, c4[a]), xi, fma(c5[a], zeta, c2[a]));
                        jc[a] = fma(fma(c7[a], eta, c6[a]), xi, fma(c5[a], eta, c3[a]));
                    }
                    double r0[3], r1[3], r2[3];
                    r0[0] = jb[1] * jc[2] - jb[2] * jc[1];
                    r0[1] = jb[2] * jc[0] - jb[0] * jc[2];
                    r0[2] = jb[0] * jc[1] - jb[1] * jc[0];
                    r1[0] = jc[1] * ja[2] - jc[2] * ja[1];
                    r1[1] = jc[2] * ja[0] - jc[0] * ja[2];
                    r1[2] = jc[0] * ja[1] - jc[1] * ja[0];
                    r2[0] = ja[1] * jb[2] - ja[2] * jb[1];
                    r2[1] = ja[2] * jb[0] - ja[0] * jb[2];
                    r2[2] = ja[0] * jb[1] - ja[1] * jb[0];
                    const double adet = fabs(ja[0] * r0[0] + ja[1] * r0[1] + ja[2] * r0[2]);
                    const double w = P.wq[qx] * P.wq[qy] * P.wq[qz];
                    const double s = P.alpha * w * rcp_nr2(adet);
                    double h[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) h[a] = r0[a] * gx + r1[a] * gy + r2[a] * gz;
                    const double fx = s * (r0[0] * h[0] + r0[1] * h[1] + r0[2] * h[2]);
                    const double fy = s * (r1[0] * h[0] + r1[1] * h[1] + r1[2] * h[2]);
                    const double fz = s * (r2[0] * h[0] + r2[1] * h[1] + r2[2] * h[2]);
#pragma unroll
                    for (int t = 0; t < N; t++) {
                        V[t][qy][qz] = fma(P.Dt[qx * N + t], fx, V[t][qy][qz]);
                        V[qx][t][qz] = fma(P.Dt[qy * N + t], fy, V[qx][t][qz]);
                        V[qx][qy][t] = fma(P.Dt[qz * N + t], fz, V[qx][qy][t]);
                    }
                    if (MASS) V[qx][qy][qz] = fma(P.beta * w * adet, U[qx][qy][qz], V[qx][qy][qz]);
                }
        // ---- back to the dofs
        contract3<2, true>(P.B, V);
        contract3<1, true>(P.B, V);
        contract3<0, true>(P.B, V);
        // ---- scatter-add; local dof (ax, ay, az): az = 0 bottom, 1 top, 2 interior
#pragma unroll
        for (int b = 0; b < N * N; b++) {
            const int ax = b / N, ay = b % N;
            const int gb = __ldg(P.map0 + (long long)col * ND + b * N) + P.off0[b * N] * lay;
            const int gt = __ldg(P.map0 + (long long)col * ND + b * N + 1) + P.off0[b * N + 1] * lay;
            const int gi = __ldg(P.map0 + (long long)col * ND + b * N + 2) + P.off0[b * N + 2] * lay;
            double vb = V[ax][ay][0], vt = V[ax][ay][1];
            const double nb = __shfl_down_sync(0xffffffffu, vb, 1);
            const int gnb = __shfl_down_sync(0xffffffffu, gb, 1);
            const int nvalid = __shfl_down_sync(0xffffffffu, (int)valid, 1);
            const int gpt = __shfl_up_sync(0xffffffffu, gt, 1);
            const bool take = lane < 31 && nvalid && gnb == gt;
            const bool taken = lane > 0 && gpt == gb;
            if (take) vt += nb;
            if (valid) {
                atomicAdd(P.y + gi, V[ax][ay][2]);
                atomicAdd(P.y + gt, vt);
                if (!taken) atomicAdd(P.y + gb, vb);
            }
        }
    }
}

}  // namespace

int fdb_launch_q2_action(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay, const fdb_int *subset,
                         double *y, const double *coords, const double *x, const fdb_int *map0,
                         const fdb_int *map1)
{
    fdb::Context &c = fdb::ctx();
    Q2Params P;
    memset(&P, 0, sizeof(P));
    P.y = y;
    P.x = x;
    P.coords = coords;
    P.map0 = map0;
    P.map1 = map1;
    P.collist = subset;
    for (int i = 0; i < ND; i++) P.off0[i] = k->h_off0[i];
    for (int i = 0; i < 8; i++) P.off1[i] = k->h_off1[i];
    P.ncols = end - start;
    P.col0 = start;
    P.nlay = nlay;
    P.items_per_col = (nlay + 31) / 32;
    P.alpha = k->desc.alpha;
    P.beta = k->desc.beta;
    for (int i = 0; i < N * N; i++) {
        P.B[i] = k->desc.B[i];
        P.Dt[i] = k->Dt[i];
    }
    for (int i = 0; i < N; i++) {
        P.wq[i] = k->desc.wq[i];
        P.xq[i] = k->desc.xq[i];
    }
    if (P.ncols <= 0 || nlay <= 0) return 0;
    const long long nitems = (long long)P.ncols * P.items_per_col;
    long long grid = (nitems + 3) / 4;
    const long long cap = (long long)c.sm_count * 8;
    if (grid > cap) grid = cap;
    if (k->desc.beta != 0.0)
        q2_action_kernel<true><<<(int)grid, 128, 0, c.stream>>>(P);
    else
        q2_action_kernel<false><<<(int)grid, 128, 0, c.stream>>>(P);
    FDB_LAUNCH_CHECK();
    return 0;
}
