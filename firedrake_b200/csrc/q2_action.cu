// Degree 2: action of  alpha*inner(grad u, grad v)*dx + beta*inner(u, v)*dx  on Q2 (x) P2
// hexahedra (27 dofs, 3^3 Gauss points), ONE THREAD PER CELL -- the degree-2 sibling of q1_action.cu.
//
// 27 values do not leave room for three separate gradient tensors in one thread's registers, so the
// arithmetic is the collocated one of the slab kernel: interpolate u to the Gauss points (three 3x3
// contractions, in place, every index static because all loops are unrolled), then at each point
// take the three derivatives straight from the collocated values with Dt = D B^{-1} (9 FMAs), apply
// the metric of the trilinear geometry (recomputed at the point, as TSFC does), and accumulate the
// transposed derivative into a second 27-value tensor; three transposed contractions bring it back
// to the dofs (storage: see the comment above the kernel).  A warp = 32 consecutive layers of one column (strided-coalesced gathers),
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
    // one third-order step r (1 + e + e^2), e = 1 - x r (|e| <= 2^-20 -> 2^-60): 3 dependent FMAs, not 4
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
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

// Both 27-value tensors live in a thread-private shared-memory scratch ([27 slots][128 threads]:
// a warp's access to one slot is 32 consecutive doubles, conflict free); the point loops over qy and
// qx are ROLLED (run-time slot indices are fine in shared memory), only qz is unrolled.  Registers
// hold the 21 geometry coefficients and 12 accumulators: 3 CTAs per SM instead of 2, no spills
// (the all-register version spilled ~85 doubles at the 255-register cap: 5.05 ms at 256^3).
// AFFINE: all cells are parallelepipeds (fdb_kernel_desc.affine_cells): cofactor rows and det J once per cell.
template <bool MASS, bool AFFINE>
__global__ void __launch_bounds__(128, 3) q2_action_kernel(const __grid_constant__ Q2Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *su = reinterpret_cast<double *>(smem_raw) + threadIdx.x;     // collocated values: su[slot * 128]
    double *sv = su + ND * 128;                                          // accumulator:       sv[slot * 128]
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
        double c1[3], c2[3], c3[3], c4[3], c5[3], c6[3], c7[3];
        {
            // ---- gather, interpolate to the Gauss points in registers, park in shared memory
            double U[N][N][N];
#pragma unroll
            for (int loc = 0; loc < ND; loc++) {
                const int g = __ldg(P.map0 + (long long)col * ND + loc) + P.off0[loc] * lay;
                U[loc / 9][(loc / 3) % 3][loc % 3] = __ldg(P.x + g);
            }
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
            contract3<0, false>(P.B, U);
            contract3<1, false>(P.B, U);
            contract3<2, false>(P.B, U);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int j = 0; j < N; j++)
#pragma unroll
                    for (int k = 0; k < N; k++) {
                        su[((i * N + j) * N + k) * 128] = U[i][j][k];
                        sv[((i * N + j) * N + k) * 128] = 0.0;
                    }
        }
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
            rdet_c = rcp_nr2(adet_c);
        }
        // ---- quadrature points: qy, qx rolled; qz unrolled
#pragma unroll 1
        for (int qy = 0; qy < N; qy++) {
            const double eta = P.xq[qy];
            double A1[3], A3[3], A6[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                A1[a] = fma(c4[a], eta, c1[a]);
                A3[a] = fma(c5[a], eta, c3[a]);
                A6[a] = fma(c7[a], eta, c6[a]);
            }
            double Vx[N][N];                       // x-part of the accumulator for this qy: [t][qz]
#pragma unroll
            for (int t = 0; t < N; t++)
#pragma unroll
                for (int k = 0; k < N; k++) Vx[t][k] = 0.0;
#pragma unroll 1
            for (int qx = 0; qx < N; qx++) {
                const double xi = P.xq[qx];
                const double dx0 = P.Dt[qx * N], dx1 = P.Dt[qx * N + 1], dx2 = P.Dt[qx * N + 2];
                const double dy0 = P.Dt[qy * N], dy1 = P.Dt[qy * N + 1], dy2 = P.Dt[qy * N + 2];
                double jc[3];
#pragma unroll
                for (int a = 0; a < 3; a++) jc[a] = fma(A6[a], xi, A3[a]);               // dX/dzeta
                // the z-line of collocated values through (qx, qy)
                double uz[N];
#pragma unroll
                for (int k = 0; k < N; k++) uz[k] = su[((qx * N + qy) * N + k) * 128];
                double Vz[N] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int qz = 0; qz < N; qz++) {
                    const double zeta = P.xq[qz];
                    const double gx = dx0 * su[((0 * N + qy) * N + qz) * 128] + dx1 * su[((1 * N + qy) * N + qz) * 128] +
                                      dx2 * su[((2 * N + qy) * N + qz) * 128];
                    const double gy = dy0 * su[((qx * N + 0) * N + qz) * 128] + dy1 * su[((qx * N + 1) * N + qz) * 128] +
                                      dy2 * su[((qx * N + 2) * N + qz) * 128];
                    const double gz = P.Dt[qz * N] * uz[0] + P.Dt[qz * N + 1] * uz[1] + P.Dt[qz * N + 2] * uz[2];
                    double r0[3], r1[3], r2[3], adet, rdet;
                    if (AFFINE) {
#pragma unroll
                        for (int a = 0; a < 3; a++) { r0[a] = R0[a]; r1[a] = R1[a]; r2[a] = R2[a]; }
                        adet = adet_c;
                        rdet = rdet_c;
                    } else {
                        double ja[3], jb[3];
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            ja[a] = fma(A6[a], zeta, A1[a]);
                            jb[a] = fma(fma(c7[a], zeta, c4[a]), xi, fma(c5[a], zeta, c2[a]));
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
                        rdet = rcp_nr2(adet);
                    }
                    const double w = P.wq[qx] * P.wq[qy] * P.wq[qz];
                    const double s = P.alpha * w * rdet;
                    double h[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) h[a] = r0[a] * gx + r1[a] * gy + r2[a] * gz;
                    const double fx = s * (r0[0] * h[0] + r0[1] * h[1] + r0[2] * h[2]);
                    const double fy = s * (r1[0] * h[0] + r1[1] * h[1] + r1[2] * h[2]);
                    const double fz = s * (r2[0] * h[0] + r2[1] * h[1] + r2[2] * h[2]);
                    Vx[0][qz] = fma(dx0, fx, Vx[0][qz]);
                    Vx[1][qz] = fma(dx1, fx, Vx[1][qz]);
                    Vx[2][qz] = fma(dx2, fx, Vx[2][qz]);
                    // y-part: V[qx][t][qz] += Dt[qy][t] fy   (run-time qx: shared memory)
                    sv[((qx * N + 0) * N + qz) * 128] = fma(dy0, fy, sv[((qx * N + 0) * N + qz) * 128]);
                    sv[((qx * N + 1) * N + qz) * 128] = fma(dy1, fy, sv[((qx * N + 1) * N + qz) * 128]);
                    sv[((qx * N + 2) * N + qz) * 128] = fma(dy2, fy, sv[((qx * N + 2) * N + qz) * 128]);
#pragma unroll
                    for (int t = 0; t < N; t++) Vz[t] = fma(P.Dt[qz * N + t], fz, Vz[t]);
                    if (MASS) Vz[qz] = fma(P.beta * w * adet, uz[qz], Vz[qz]);
                }
#pragma unroll
                for (int t = 0; t < N; t++)
                    sv[((qx * N + qy) * N + t) * 128] += Vz[t];
            }
#pragma unroll
            for (int t = 0; t < N; t++)
#pragma unroll
                for (int k = 0; k < N; k++) sv[((t * N + qy) * N + k) * 128] += Vx[t][k];
        }
        // ---- back to the dofs
        double V[N][N][N];
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++)
#pragma unroll
                for (int k = 0; k < N; k++) V[i][j][k] = sv[((i * N + j) * N + k) * 128];
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
    constexpr int SMEM = 2 * ND * 128 * (int)sizeof(double);      // 55 KB: U and V scratch of the 128 threads
    static bool configured = false;
    if (!configured) {
        FDB_CUDA(cudaFuncSetAttribute(q2_action_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        FDB_CUDA(cudaFuncSetAttribute(q2_action_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        FDB_CUDA(cudaFuncSetAttribute(q2_action_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        FDB_CUDA(cudaFuncSetAttribute(q2_action_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    const bool mass = k->desc.beta != 0.0, aff = k->desc.affine_cells != 0;
    if (mass && aff) q2_action_kernel<true, true><<<(int)grid, 128, SMEM, c.stream>>>(P);
    else if (mass) q2_action_kernel<true, false><<<(int)grid, 128, SMEM, c.stream>>>(P);
    else if (aff) q2_action_kernel<false, true><<<(int)grid, 128, SMEM, c.stream>>>(P);
    else q2_action_kernel<false, false><<<(int)grid, 128, SMEM, c.stream>>>(P);
    FDB_LAUNCH_CHECK();
    return 0;
}
