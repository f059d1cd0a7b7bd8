// Explicit element matrices of  alpha*inner(grad u, grad v)*dx + beta*inner(u, v)*dx  on
// Q_p (x) P_p hexahedra as a DENSE  B^T D B  contraction on the fp64 tensor pipe
// (mma.sync.m8n8k4.f64 -> DMMA), scattered into the device CSR with MatSetValuesLocal
// semantics.  BASELINE.json config 4 (vector Helmholtz CG4) and every 2-form of degree >= 3.
//
// Reference: the element tensor is what TSFC's generated kernel computes
// (tsfc/kernel_interface/common.py:139-239, tsfc/fem.py:710-804); the scatter is
// MatSetValues[Blocked]Local(ADD_VALUES) with masked local-to-global maps
// (pyop2/codegen/builder.py:520-625, pyop2/parloop.py:279-314).
//
// Per cell (n = (p+1)^3 dofs, Q = (p+1)^3 Gauss points):
//     A[i][j] = sum_q  sum_{a,b} dphi_i/dxi_a(q) G_q[a][b] dphi_j/dxi_b(q)  +  m_q phi_i(q) phi_j(q)
// with  G_q = alpha w_q / |det J_q| * K_q K_q^T  (K = cofactor rows of J, trilinear Q1 geometry
// recomputed at every point as TSFC does, tsfc/ufl_utils.py:41-85)  and  m_q = beta w_q |det J_q|.
// Written as a GEMM over k = (q, r), r = 0..3:
//     L[k][i] = { dphi_i/dxi_r(q), r < 3 ;  phi_i(q), r = 3 }        cell independent (table T)
//     R[k][j] = { sum_b G_q[r][b] L[(q,b)][j], r < 3 ;  m_q L[(q,3)][j] }   per cell, 10 FMAs per entry
//     A = L^T R                                                       4 Q n^2 FMAs: the DMMA part
// One CTA (8 warps) owns one cell at a time; the n x n accumulator (padded to NP) lives in
// registers as m8n8 DMMA tiles, the k dimension streams through shared memory in chunks of
// KQ quadrature points: L-chunk by cp.async from the L2-resident table (double buffered), R-chunk
// computed in place.  Shared-memory rows are padded to a stride of NP + 4 doubles so that the
// DMMA fragment loads (8 consecutive i for each of 4 consecutive k) are bank-conflict free.
//
// Roofline: DMMA-pipe bound.  Per cell 4*Qp*NP^2 FMAs issued (Qp, NP: padded sizes); CG4: 8.39 M
// against 7.81 M useful (93 %).  The scatter is n^2 RED.ADD.F64 per cell into L2.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int KQ = 4;               // quadrature points per k-chunk (16 k-rows = 4 DMMA k-steps)
constexpr int KR = 4 * KQ;          // k-rows per chunk

struct BdbParams {
    const double *coords;
    const int *map0;
    const int *map1;
    const int *collist;
    const int *off0;         // device, n entries
    const int *off1;         // device, 8 entries
    int ncols, col0, nlay;
    const double *table;     // [Qp*4][NP]
    const long long *rowptr;
    const int *colidx;
    double *vals;
    const int *row_lg, *col_lg;
    const unsigned short *rank_tab;
    int nvar;
    double alpha, beta;
    double xq[FDB_MAX_1D], wq[FDB_MAX_1D];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// metric G_q (6 entries, alpha w / |det J| K K^T) and mass factor beta w |det J| of one quadrature point
template <int N>
__device__ __forceinline__ void geometry_point(const BdbParams &P, const double *sX, double *g, int q, bool real)
{
    if (!real) {
#pragma unroll
        for (int e = 0; e < 7; e++) g[e] = 0.0;
        return;
    }
    const int qx = q / (N * N), qy = (q / N) % N, qz = q % N;
    const double xi = P.xq[qx], eta = P.xq[qy], zeta = P.xq[qz];
    double ja[3], jb[3], jc[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const double X000 = sX[0 * 3 + a], X001 = sX[1 * 3 + a], X010 = sX[2 * 3 + a],
                     X011 = sX[3 * 3 + a], X100 = sX[4 * 3 + a], X101 = sX[5 * 3 + a],
                     X110 = sX[6 * 3 + a], X111 = sX[7 * 3 + a];
        const double c1 = X100 - X000, c2 = X010 - X000, c3 = X001 - X000;
        const double c4 = X110 - X100 - X010 + X000, c5 = X011 - X010 - X001 + X000,
                     c6 = X101 - X100 - X001 + X000;
        const double c7 = X111 - X110 - X101 - X011 + X100 + X010 + X001 - X000;
        ja[a] = c1 + c4 * eta + (c6 + c7 * eta) * zeta;      // dx/dxi
        jb[a] = c2 + c5 * zeta + (c4 + c7 * zeta) * xi;      // dx/deta
        jc[a] = c3 + c5 * eta + (c6 + c7 * eta) * xi;        // dx/dzeta
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
    const double s = P.alpha * w / adet;
    g[0] = s * (r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]);
    g[1] = s * (r0[0] * r1[0] + r0[1] * r1[1] + r0[2] * r1[2]);
    g[2] = s * (r0[0] * r2[0] + r0[1] * r2[1] + r0[2] * r2[2]);
    g[3] = s * (r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
    g[4] = s * (r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2]);
    g[5] = s * (r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
    g[6] = P.beta * w * adet;
}

template <int N>
struct BdbCfg {
    static constexpr int ND = N * N * N;
    static constexpr int NP = (ND + 31) / 32 * 32;            // padded dofs
    static constexpr int Q3 = N * N * N;
    static constexpr int NCHUNK = (Q3 + KQ - 1) / KQ;
    static constexpr int WARPS = 8;
    static constexpr int WM = (NP >= 128) ? 4 : 2;             // warp grid over (i, j)
    static constexpr int WN = WARPS / WM;
    static constexpr int TM = NP / WM, TN = NP / WN;           // warp tile
    static constexpr int MT = TM / 8, NT = TN / 8;             // DMMA tiles per warp
    static constexpr int S = NP + 4;                           // smem row stride (doubles)
    static constexpr int SMEM_DOUBLES = 3 * KR * S             // L (2 buffers) + R
                                        + NCHUNK * KQ * 7      // G (6) + mass factor per point
                                        + 24;                  // vertex coordinates
    static constexpr int SMEM_BYTES = SMEM_DOUBLES * 8 + NP * 8 + NP * 4 + 16;
};

template <int N>
__global__ void __launch_bounds__(256, (N >= 5) ? 1 : 3) bdb_matrix_kernel(const __grid_constant__ BdbParams P)
{
    using C = BdbCfg<N>;
    constexpr int ND = C::ND, NP = C::NP, S = C::S, Q3 = C::Q3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sL = reinterpret_cast<double *>(smem_raw);        // [2][KR][S]
    double *sR = sL + 2 * KR * S;                              // [KR][S]
    double *sG = sR + KR * S;                                  // [NCHUNK*KQ][7]
    double *sX = sG + C::NCHUNK * KQ * 7;                      // [24]
    long long *sRow = reinterpret_cast<long long *>(sX + 24);  // [NP] start of the CSR row, -1 = dropped
    int *sCol = reinterpret_cast<int *>(sRow + NP);            // [NP] global column (after lgmap), -1 = dropped

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / C::WN, wn = warp % C::WN;
    const int gid = lane >> 2, tig = lane & 3;
    const long long ncells = (long long)P.ncols * P.nlay;

    for (long long cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
        const int ci = (int)(cell / P.nlay), layer = (int)(cell - (long long)ci * P.nlay);
        const int col = P.collist ? P.collist[ci] : P.col0 + ci;
        __syncthreads();                                       // previous cell's epilogue is done with smem
        // ---- first L chunk in flight while the geometry is computed
        auto stage_L = [&](int chunk, int buf) {
            const double *src = P.table + (size_t)chunk * KR * NP;
            double *dst = sL + buf * KR * S;
            for (int e = tid; e < KR * NP / 2; e += 256) {
                const int row = e / (NP / 2), c2 = e - row * (NP / 2);
                cp_async16(dst + row * S + 2 * c2, src + row * NP + 2 * c2);
            }
            cp_async_commit();
        };
        stage_L(0, 0);
        if (tid < 24) {
            const int v = tid / 3, a = tid - v * 3;
            const int g = P.map1[(long long)col * 8 + v] + P.off1[v] * layer;
            sX[tid] = P.coords[(long long)g * 3 + a];
        }
        for (int i = tid; i < NP; i += 256) {
            long long rs = -1;
            int gc = -1;
            if (i < ND) {
                const int g = P.map0[(long long)col * ND + i] + P.off0[i] * layer;
                int gr = P.row_lg ? P.row_lg[g] : g;
                gc = P.col_lg ? P.col_lg[g] : g;
                if (gr >= 0) rs = P.rowptr[gr];
                if (!P.rank_tab && gr >= 0) rs = gr;           // binary-search mode keeps the row index
            }
            sRow[i] = rs;
            sCol[i] = gc;
        }
        __syncthreads();
        for (int q = tid; q < C::NCHUNK * KQ; q += 256) geometry_point<N>(P, sX, sG + q * 7, q, q < Q3);

        double acc[C::MT][C::NT][2];
#pragma unroll
        for (int m = 0; m < C::MT; m++)
#pragma unroll
            for (int n = 0; n < C::NT; n++) acc[m][n][0] = acc[m][n][1] = 0.0;

        for (int chunk = 0; chunk < C::NCHUNK; chunk++) {
            const int buf = chunk & 1;
            cp_async_wait_all();
            __syncthreads();              // L[buf] landed; G visible (first trip); previous MMAs done with sR / L[buf^1]
            if (chunk + 1 < C::NCHUNK) stage_L(chunk + 1, buf ^ 1);
            const double *L = sL + buf * KR * S;
            // ---- R chunk: one (point, dof) pair per pass
            for (int e = tid; e < KQ * NP; e += 256) {
                const int ql = e / NP, j = e - ql * NP;
                const double *g = sG + (chunk * KQ + ql) * 7;
                const double l0 = L[(ql * 4 + 0) * S + j], l1 = L[(ql * 4 + 1) * S + j],
                             l2 = L[(ql * 4 + 2) * S + j], l3 = L[(ql * 4 + 3) * S + j];
                sR[(ql * 4 + 0) * S + j] = g[0] * l0 + g[1] * l1 + g[2] * l2;
                sR[(ql * 4 + 1) * S + j] = g[1] * l0 + g[3] * l1 + g[4] * l2;
                sR[(ql * 4 + 2) * S + j] = g[2] * l0 + g[4] * l1 + g[5] * l2;
                sR[(ql * 4 + 3) * S + j] = g[6] * l3;
            }
            __syncthreads();
            // ---- A += L^T R on the tensor pipe
#pragma unroll
            for (int ks = 0; ks < KR / 4; ks++) {
                double af[C::MT], bf[C::NT];
                const double *Lk = L + (ks * 4 + tig) * S + wm * C::TM + gid;
                const double *Rk = sR + (ks * 4 + tig) * S + wn * C::TN + gid;
#pragma unroll
                for (int m = 0; m < C::MT; m++) af[m] = Lk[m * 8];
#pragma unroll
                for (int n = 0; n < C::NT; n++) bf[n] = Rk[n * 8];
#pragma unroll
                for (int m = 0; m < C::MT; m++)
#pragma unroll
                    for (int n = 0; n < C::NT; n++) dmma_m8n8k4(acc[m][n][0], acc[m][n][1], af[m], bf[n]);
            }
        }

        // ---- MatSetValuesLocal(ADD_VALUES): rows/columns with a negative local-to-global index are dropped
        const unsigned short *rk = nullptr;
        if (P.rank_tab) {
            const int v = P.nlay < 3 ? layer : (layer == 0 ? 0 : (layer == P.nlay - 1 ? 2 : 1));
            rk = P.rank_tab + ((long long)col * P.nvar + v) * ND * ND;
        }
#pragma unroll
        for (int m = 0; m < C::MT; m++) {
            const int i = wm * C::TM + m * 8 + gid;
            const long long rs = sRow[i];
            if (i >= ND || rs < 0) continue;
#pragma unroll
            for (int n = 0; n < C::NT; n++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int j = wn * C::TN + n * 8 + tig * 2 + h;
                    if (j >= ND) continue;
                    const int gc = sCol[j];
                    if (gc < 0) continue;
                    long long pos;
                    if (rk) {
                        pos = rs + rk[j * ND + i];
                    } else {
                        long long lo = P.rowptr[rs], hi = P.rowptr[rs + 1];
                        while (hi - lo > 1) {
                            const long long mid = (lo + hi) >> 1;
                            if (P.colidx[mid] <= gc) lo = mid; else hi = mid;
                        }
                        pos = lo;
                    }
                    atomicAdd(P.vals + pos, acc[m][n][h]);
                }
        }
    }
}

// ---- symmetric variant for NP = 128 (p = 4, config 4) -----------------------------------------
// A = L^T (D L) is symmetric (G_q is), so only the upper triangle of the 16 x 16 grid of 8 x 8
// DMMA tiles is computed: 136 tiles instead of 256.  Warps 0..5 own the six off-diagonal 32 x 32
// super-blocks (16 tiles each), warps 6 and 7 the upper tiles of two diagonal super-blocks each
// (2 x 10 tiles); the accumulator shrinks from 64 to 40 doubles per thread, which lets TWO cells
// be resident per SM: one cell's scatter (n^2 RED.ADD.F64, the mirrored entry written from the same
// register) overlaps the other cell's DMMAs.
__device__ __forceinline__ constexpr int dtile(int m, int n)   // index of tile (m <= n) in a 4x4 upper triangle
{
    return m * 4 - m * (m - 1) / 2 + (n - m);
}

template <int N>
__global__ void __launch_bounds__(256, 2) bdb_matrix_sym_kernel(const __grid_constant__ BdbParams P)
{
    using C = BdbCfg<N>;
    constexpr int ND = C::ND, NP = C::NP, S = C::S, Q3 = C::Q3;
    static_assert(NP == 128, "the symmetric tiling is written for a 128 x 128 padded element matrix");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sL = reinterpret_cast<double *>(smem_raw);        // [2][KR][S]
    double *sR = sL + 2 * KR * S;                              // [KR][S]
    double *sG = sR + KR * S;                                  // [NCHUNK*KQ][7]
    double *sX = sG + C::NCHUNK * KQ * 7;                      // [24]
    long long *sRow = reinterpret_cast<long long *>(sX + 24);  // [NP]
    int *sCol = reinterpret_cast<int *>(sRow + NP);            // [NP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const long long ncells = (long long)P.ncols * P.nlay;
    // off-diagonal super-block of warps 0..5: (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    const int bi = warp < 3 ? 0 : (warp < 5 ? 1 : 2);
    const int bj = warp < 3 ? warp + 1 : (warp < 5 ? warp - 1 : 3);
    const bool offdiag = warp < 6;
    const int d0 = (warp - 6) * 2;                             // diagonal super-blocks d0, d0 + 1 (warps 6, 7)

    for (long long cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
        const int ci = (int)(cell / P.nlay), layer = (int)(cell - (long long)ci * P.nlay);
        const int col = P.collist ? P.collist[ci] : P.col0 + ci;
        __syncthreads();
        auto stage_L = [&](int chunk, int buf) {
            const double *src = P.table + (size_t)chunk * KR * NP;
            double *dst = sL + buf * KR * S;
            for (int e = tid; e < KR * NP / 2; e += 256) {
                const int row = e / (NP / 2), c2 = e - row * (NP / 2);
                cp_async16(dst + row * S + 2 * c2, src + row * NP + 2 * c2);
            }
            cp_async_commit();
        };
        stage_L(0, 0);
        if (tid < 24) {
            const int v = tid / 3, a = tid - v * 3;
            const int g = P.map1[(long long)col * 8 + v] + P.off1[v] * layer;
            sX[tid] = P.coords[(long long)g * 3 + a];
        }
        for (int i = tid; i < NP; i += 256) {
            long long rs = -1;
            int gc = -1;
            if (i < ND) {
                const int g = P.map0[(long long)col * ND + i] + P.off0[i] * layer;
                int gr = P.row_lg ? P.row_lg[g] : g;
                gc = P.col_lg ? P.col_lg[g] : g;
                if (gr >= 0) rs = P.rank_tab ? P.rowptr[gr] : (long long)gr;
            }
            sRow[i] = rs;
            sCol[i] = gc;
        }
        __syncthreads();
        for (int q = tid; q < C::NCHUNK * KQ; q += 256) geometry_point<N>(P, sX, sG + q * 7, q, q < Q3);

        double acc[40];
#pragma unroll
        for (int e = 0; e < 40; e++) acc[e] = 0.0;

        for (int chunk = 0; chunk < C::NCHUNK; chunk++) {
            const int buf = chunk & 1;
            cp_async_wait_all();
            __syncthreads();
            if (chunk + 1 < C::NCHUNK) stage_L(chunk + 1, buf ^ 1);
            const double *L = sL + buf * KR * S;
            for (int e = tid; e < KQ * NP; e += 256) {
                const int ql = e / NP, j = e - ql * NP;
                const double *g = sG + (chunk * KQ + ql) * 7;
                const double l0 = L[(ql * 4 + 0) * S + j], l1 = L[(ql * 4 + 1) * S + j],
                             l2 = L[(ql * 4 + 2) * S + j], l3 = L[(ql * 4 + 3) * S + j];
                sR[(ql * 4 + 0) * S + j] = g[0] * l0 + g[1] * l1 + g[2] * l2;
                sR[(ql * 4 + 1) * S + j] = g[1] * l0 + g[3] * l1 + g[4] * l2;
                sR[(ql * 4 + 2) * S + j] = g[2] * l0 + g[4] * l1 + g[5] * l2;
                sR[(ql * 4 + 3) * S + j] = g[6] * l3;
            }
            __syncthreads();
            if (offdiag) {
#pragma unroll
                for (int ks = 0; ks < KR / 4; ks++) {
                    double af[4], bf[4];
                    const double *Lk = L + (ks * 4 + tig) * S + bi * 32 + gid;
                    const double *Rk = sR + (ks * 4 + tig) * S + bj * 32 + gid;
#pragma unroll
                    for (int m = 0; m < 4; m++) af[m] = Lk[m * 8];
#pragma unroll
                    for (int n = 0; n < 4; n++) bf[n] = Rk[n * 8];
#pragma unroll
                    for (int m = 0; m < 4; m++)
#pragma unroll
                        for (int n = 0; n < 4; n++)
                            dmma_m8n8k4(acc[(m * 4 + n) * 2], acc[(m * 4 + n) * 2 + 1], af[m], bf[n]);
                }
            } else {
#pragma unroll
                for (int ks = 0; ks < KR / 4; ks++) {
#pragma unroll
                    for (int d = 0; d < 2; d++) {
                        double af[4], bf[4];
                        const double *Lk = L + (ks * 4 + tig) * S + (d0 + d) * 32 + gid;
                        const double *Rk = sR + (ks * 4 + tig) * S + (d0 + d) * 32 + gid;
#pragma unroll
                        for (int m = 0; m < 4; m++) {
                            af[m] = Lk[m * 8];
                            bf[m] = Rk[m * 8];
                        }
#pragma unroll
                        for (int m = 0; m < 4; m++)
#pragma unroll
                            for (int n = m; n < 4; n++)
                                dmma_m8n8k4(acc[(d * 10 + dtile(m, n)) * 2], acc[(d * 10 + dtile(m, n)) * 2 + 1],
                                            af[m], bf[n]);
                    }
                }
            }
        }

        // ---- scatter: (i, j) and, off the diagonal tiles, the mirrored (j, i)
        const unsigned short *rk = nullptr;
        if (P.rank_tab) {
            const int v = P.nlay < 3 ? layer : (layer == 0 ? 0 : (layer == P.nlay - 1 ? 2 : 1));
            rk = P.rank_tab + ((long long)col * P.nvar + v) * ND * ND;
        }
        auto add = [&](int i, int j, double val) {     // row (test) dof i, column (trial) dof j
            const long long rs = sRow[i];
            const int gc = sCol[j];
            if (rs < 0 || gc < 0) return;
            long long pos;
            if (rk) {
                pos = rs + rk[j * ND + i];
            } else {
                long long lo = P.rowptr[rs], hi = P.rowptr[rs + 1];
                while (hi - lo > 1) {
                    const long long mid = (lo + hi) >> 1;
                    if (P.colidx[mid] <= gc) lo = mid; else hi = mid;
                }
                pos = lo;
            }
            atomicAdd(P.vals + pos, val);
        };
        if (offdiag) {
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int i = bi * 32 + m * 8 + gid;
#pragma unroll
                for (int n = 0; n < 4; n++)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int j = bj * 32 + n * 8 + tig * 2 + h;
                        if (i < ND && j < ND) {
                            add(i, j, acc[(m * 4 + n) * 2 + h]);
                            add(j, i, acc[(m * 4 + n) * 2 + h]);
                        }
                    }
            }
        } else {
#pragma unroll
            for (int d = 0; d < 2; d++)
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int i = (d0 + d) * 32 + m * 8 + gid;
#pragma unroll
                    for (int n = m; n < 4; n++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int j = (d0 + d) * 32 + n * 8 + tig * 2 + h;
                            if (i < ND && j < ND) {
                                const double val = acc[(d * 10 + dtile(m, n)) * 2 + h];
                                add(i, j, val);
                                if (n != m) add(j, i, val);      // diagonal tiles hold both halves themselves
                            }
                        }
                }
        }
    }
}

// ---- symmetric variant for NP = 64 (p = 3): TWO cells per CTA ---------------------------------
// 8 x 8 grid of DMMA tiles, upper triangle = 36 tiles, 4 warps per cell: two warps share the
// off-diagonal 32 x 32 super-block (8 tiles each), two take the upper tiles of a diagonal
// super-block (10 each).  The cell-independent L chunk is staged once for both cells.
template <int N>
__global__ void __launch_bounds__(256, 3) bdb_matrix_sym64_kernel(const __grid_constant__ BdbParams P)
{
    using C = BdbCfg<N>;
    constexpr int ND = C::ND, NP = C::NP, S = C::S, Q3 = C::Q3;
    static_assert(NP == 64, "written for a 64 x 64 element matrix");
    constexpr int QP = C::NCHUNK * KQ;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sL = reinterpret_cast<double *>(smem_raw);        // [2][KR][S]
    double *sR = sL + 2 * KR * S;                              // [2 cells][KR][S]
    double *sG = sR + 2 * KR * S;                              // [2][QP][7]
    double *sX = sG + 2 * QP * 7;                              // [2][24]
    long long *sRow = reinterpret_cast<long long *>(sX + 48);  // [2][NP]
    int *sCol = reinterpret_cast<int *>(sRow + 2 * NP);        // [2][NP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = warp >> 2, w4 = warp & 3;
    const int gid = lane >> 2, tig = lane & 3;
    const long long ncells = (long long)P.ncols * P.nlay;
    const long long npairs = (ncells + 1) / 2;

    for (long long pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        __syncthreads();
        auto stage_L = [&](int chunk, int buf) {
            const double *src = P.table + (size_t)chunk * KR * NP;
            double *dst = sL + buf * KR * S;
            for (int e = tid; e < KR * NP / 2; e += 256) {
                const int row = e / (NP / 2), c2 = e - row * (NP / 2);
                cp_async16(dst + row * S + 2 * c2, src + row * NP + 2 * c2);
            }
            cp_async_commit();
        };
        stage_L(0, 0);
        int colv[2], layv[2];
        bool live[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const long long cell = 2 * pair + c;
            live[c] = cell < ncells;
            const long long cc = live[c] ? cell : ncells - 1;
            const int ci = (int)(cc / P.nlay);
            layv[c] = (int)(cc - (long long)ci * P.nlay);
            colv[c] = P.collist ? P.collist[ci] : P.col0 + ci;
        }
        if (tid < 48) {
            const int c = tid / 24, t = tid - c * 24, v = t / 3, a = t - v * 3;
            const int g = P.map1[(long long)colv[c] * 8 + v] + P.off1[v] * layv[c];
            sX[tid] = P.coords[(long long)g * 3 + a];
        }
        if (tid < 2 * NP) {
            const int c = tid / NP, i = tid - c * NP;
            long long rs = -1;
            int gc = -1;
            if (i < ND && live[c]) {
                const int g = P.map0[(long long)colv[c] * ND + i] + P.off0[i] * layv[c];
                int gr = P.row_lg ? P.row_lg[g] : g;
                gc = P.col_lg ? P.col_lg[g] : g;
                if (gr >= 0) rs = P.rank_tab ? P.rowptr[gr] : (long long)gr;
            }
            sRow[tid] = rs;
            sCol[tid] = gc;
        }
        __syncthreads();
        for (int e = tid; e < 2 * QP; e += 256) {
            const int c = e / QP, q = e - c * QP;
            geometry_point<N>(P, sX + c * 24, sG + e * 7, q, q < Q3);
        }
        double acc[20];
#pragma unroll
        for (int e = 0; e < 20; e++) acc[e] = 0.0;

        for (int chunk = 0; chunk < C::NCHUNK; chunk++) {
            const int buf = chunk & 1;
            cp_async_wait_all();
            __syncthreads();
            if (chunk + 1 < C::NCHUNK) stage_L(chunk + 1, buf ^ 1);
            const double *L = sL + buf * KR * S;
            for (int e = tid; e < 2 * KQ * NP; e += 256) {
                const int c = e / (KQ * NP), r = e - c * (KQ * NP);
                const int ql = r / NP, j = r - ql * NP;
                const double *g = sG + (c * QP + chunk * KQ + ql) * 7;
                double *R = sR + c * KR * S;
                const double l0 = L[(ql * 4 + 0) * S + j], l1 = L[(ql * 4 + 1) * S + j],
                             l2 = L[(ql * 4 + 2) * S + j], l3 = L[(ql * 4 + 3) * S + j];
                R[(ql * 4 + 0) * S + j] = g[0] * l0 + g[1] * l1 + g[2] * l2;
                R[(ql * 4 + 1) * S + j] = g[1] * l0 + g[3] * l1 + g[4] * l2;
                R[(ql * 4 + 2) * S + j] = g[2] * l0 + g[4] * l1 + g[5] * l2;
                R[(ql * 4 + 3) * S + j] = g[6] * l3;
            }
            __syncthreads();
            const double *R = sR + slot * KR * S;
            if (w4 < 2) {
#pragma unroll
                for (int ks = 0; ks < KR / 4; ks++) {
                    double af[2], bf[4];
                    const double *Lk = L + (ks * 4 + tig) * S + w4 * 16 + gid;
                    const double *Rk = R + (ks * 4 + tig) * S + 32 + gid;
#pragma unroll
                    for (int m = 0; m < 2; m++) af[m] = Lk[m * 8];
#pragma unroll
                    for (int n = 0; n < 4; n++) bf[n] = Rk[n * 8];
#pragma unroll
                    for (int m = 0; m < 2; m++)
#pragma unroll
                        for (int n = 0; n < 4; n++)
                            dmma_m8n8k4(acc[(m * 4 + n) * 2], acc[(m * 4 + n) * 2 + 1], af[m], bf[n]);
                }
            } else {
                const int d = w4 - 2;
#pragma unroll
                for (int ks = 0; ks < KR / 4; ks++) {
                    double af[4], bf[4];
                    const double *Lk = L + (ks * 4 + tig) * S + d * 32 + gid;
                    const double *Rk = R + (ks * 4 + tig) * S + d * 32 + gid;
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        af[m] = Lk[m * 8];
                        bf[m] = Rk[m * 8];
                    }
#pragma unroll
                    for (int m = 0; m < 4; m++)
#pragma unroll
                        for (int n = m; n < 4; n++)
                            dmma_m8n8k4(acc[dtile(m, n) * 2], acc[dtile(m, n) * 2 + 1], af[m], bf[n]);
                }
            }
        }

        if (!live[slot]) continue;
        const unsigned short *rk = nullptr;
        if (P.rank_tab) {
            const int layer = layv[slot];
            const int v = P.nlay < 3 ? layer : (layer == 0 ? 0 : (layer == P.nlay - 1 ? 2 : 1));
            rk = P.rank_tab + ((long long)colv[slot] * P.nvar + v) * ND * ND;
        }
        const long long *row = sRow + slot * NP;
        const int *colg = sCol + slot * NP;
        auto add = [&](int i, int j, double val) {
            const long long rs = row[i];
            const int gc = colg[j];
            if (rs < 0 || gc < 0) return;
            long long pos;
            if (rk) {
                pos = rs + rk[j * ND + i];
            } else {
                long long lo = P.rowptr[rs], hi = P.rowptr[rs + 1];
                while (hi - lo > 1) {
                    const long long mid = (lo + hi) >> 1;
                    if (P.colidx[mid] <= gc) lo = mid; else hi = mid;
                }
                pos = lo;
            }
            atomicAdd(P.vals + pos, val);
        };
        if (w4 < 2) {
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const int i = w4 * 16 + m * 8 + gid;
#pragma unroll
                for (int n = 0; n < 4; n++)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int j = 32 + n * 8 + tig * 2 + h;
                        if (i < ND && j < ND) {
                            add(i, j, acc[(m * 4 + n) * 2 + h]);
                            add(j, i, acc[(m * 4 + n) * 2 + h]);
                        }
                    }
            }
        } else {
            const int d = w4 - 2;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int i = d * 32 + m * 8 + gid;
#pragma unroll
                for (int n = m; n < 4; n++)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int j = d * 32 + n * 8 + tig * 2 + h;
                        if (i < ND && j < ND) {
                            const double val = acc[dtile(m, n) * 2 + h];
                            add(i, j, val);
                            if (n != m) add(j, i, val);
                        }
                    }
            }
        }
    }
}

// table T[(q*4 + r)][NP]: r < 3 reference gradient component r of basis i at point q, r = 3 its value
template <int N>
int build_table(fdb_kernel_s *k)
{
    using C = BdbCfg<N>;
    if (k->d_bdb_table) return 0;
    const size_t rows = (size_t)C::NCHUNK * KR;
    std::vector<double> T(rows * C::NP, 0.0);
    const double *B = k->desc.B, *D = k->desc.D;
    for (int qx = 0; qx < N; qx++)
        for (int qy = 0; qy < N; qy++)
            for (int qz = 0; qz < N; qz++) {
                const int q = (qx * N + qy) * N + qz;
                for (int ax = 0; ax < N; ax++)
                    for (int ay = 0; ay < N; ay++)
                        for (int az = 0; az < N; az++) {
                            const int i = (ax * N + ay) * N + az;
                            const double bx = B[qx * N + ax], by = B[qy * N + ay], bz = B[qz * N + az];
                            const double dx = D[qx * N + ax], dy = D[qy * N + ay], dz = D[qz * N + az];
                            T[((size_t)q * 4 + 0) * C::NP + i] = dx * by * bz;
                            T[((size_t)q * 4 + 1) * C::NP + i] = bx * dy * bz;
                            T[((size_t)q * 4 + 2) * C::NP + i] = bx * by * dz;
                            T[((size_t)q * 4 + 3) * C::NP + i] = bx * by * bz;
                        }
            }
    FDB_CUDA(cudaMalloc(&k->d_bdb_table, T.size() * sizeof(double)));
    FDB_CUDA(cudaMemcpy(k->d_bdb_table, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

template <int N>
int launch_bdb(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay, const fdb_int *subset, fdb_mat_t mat,
               const double *coords, const fdb_int *map0, const fdb_int *map1)
{
    using C = BdbCfg<N>;
    fdb::Context &c = fdb::ctx();
    if (build_table<N>(k)) return 1;
    BdbParams P;
    memset(&P, 0, sizeof(P));
    P.coords = coords;
    P.map0 = map0;
    P.map1 = map1;
    P.collist = subset;
    P.off0 = k->d_off0;
    P.off1 = k->d_off1;
    P.ncols = end - start;
    P.col0 = start;
    P.nlay = nlay;
    P.table = k->d_bdb_table;
    fdb_mat_device_view(mat, &P.rowptr, &P.colidx, &P.vals, &P.row_lg, &P.col_lg);
    fdb_mat_rank_table(mat, &P.rank_tab, &P.nvar);
    if (subset || k->desc.cell != FDB_CELL_HEX_EXTRUDED) P.rank_tab = nullptr;   // table is per column of the full set
    P.alpha = k->desc.alpha;
    P.beta = k->desc.beta;
    for (int i = 0; i < N; i++) {
        P.xq[i] = k->desc.xq[i];
        P.wq[i] = k->desc.wq[i];
    }
    if (P.ncols <= 0 || nlay <= 0) return 0;
    static const bool sym_on = !(getenv("FDB_BDB_SYM") && atoi(getenv("FDB_BDB_SYM")) == 0);
    void (*kern)(const BdbParams) = bdb_matrix_kernel<N>;
    int smem = C::SMEM_BYTES;
    int cells_per_cta = 1;
    if (sym_on && C::NP == 128) {
        kern = bdb_matrix_sym_kernel<(C::NP == 128 ? N : 5)>;
    } else if (sym_on && C::NP == 64) {
        kern = bdb_matrix_sym64_kernel<(C::NP == 64 ? N : 4)>;
        using C4 = BdbCfg<4>;
        smem = (4 * KR * C4::S + 2 * C4::NCHUNK * KQ * 7 + 48) * 8 + 2 * C4::NP * 8 + 2 * C4::NP * 4 + 16;
        cells_per_cta = 2;
    }
    static bool configured = false;
    static int occ = 1;
    if (!configured) {
        FDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        if (occ < 1) occ = 1;
        configured = true;
    }
    const long long ncells = (long long)P.ncols * nlay;
    long long grid = (long long)c.sm_count * occ;
    const long long nwork = (ncells + cells_per_cta - 1) / cells_per_cta;
    if (grid > nwork) grid = nwork;
    kern<<<(int)grid, 256, smem, c.stream>>>(P);
    FDB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// Dense B^T D B element matrices on the fp64 tensor pipe; degrees 2..5 (n = 27..216 does not fit
// the register accumulator beyond p = 4: p = 5 stays on the sum-factorised path).
int fdb_launch_helmholtz_matrix_dmma(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay,
                                     const fdb_int *subset, fdb_mat_t mat, const double *coords,
                                     const fdb_int *map0, const fdb_int *map1)
{
    switch (k->n1d) {
    case 3: return launch_bdb<3>(k, start, end, nlay, subset, mat, coords, map0, map1);
    case 4: return launch_bdb<4>(k, start, end, nlay, subset, mat, coords, map0, map1);
    case 5: return launch_bdb<5>(k, start, end, nlay, subset, mat, coords, map0, map1);
    }
    fdb::set_error("helmholtz matrix (DMMA): degree %d not instantiated (2..4)", k->n1d - 1);
    return 1;
}
