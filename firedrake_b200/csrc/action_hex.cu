// Matrix-free action of  alpha*inner(grad u, grad v)*dx + beta*inner(u, v)*dx
// on Q_p (x) P_p hexahedra (extruded or native), fp64, sm_100a.
//
// One launch = gather through the cell->node map + element kernel + scatter-add,
// i.e. the whole PyOP2 wrapper of SURVEY.md section 9.1 (reference
// pyop2/codegen/builder.py:80-128, 352-429, 730-812) fused with the TSFC kernel
// it calls (reference tsfc/kernel_interface/common.py:139-239).
//
// Thread mapping ("slab threads"): N = p+1 lanes cooperate on one cell, each
// lane owning one N x N slab of the N^3 tensor; a warp holds 32/N cells, which
// are CONSECUTIVE LAYERS of one column so that a warp-wide gather instruction
// walks a contiguous run of each dof column.  Two slab orientations are used:
//   layout Z: lane t owns index t along z, holds [x][y]   (gather / scatter)
//   layout Y: lane t owns index t along y, holds [x][z]   (quadrature points)
// Contractions along in-slab axes run in registers (N^2 x N FMAs against a
// constant-bank table); the two orientation changes go through a per-warp
// shared-memory tile with __syncwarp only -- no block-level barrier anywhere.
//
// Arithmetic: the basis is first interpolated to the N Gauss points per axis
// (B (x) B (x) B), gradients are then taken with the collocated derivative
// matrix Dt = D B^{-1}; the transpose path mirrors it.  6 N^4 FMAs each way
// instead of 8 N^4 for the textbook form; identical in exact arithmetic.
// Geometry (trilinear Q1 coordinate field) is recomputed at every quadrature
// point, as TSFC does (reference tsfc/ufl_utils.py:41-85), from the 8 vertex
// coordinates: cofactor rows r_k of J, det = a.(b x c), and the flux in
// reference coordinates is  (alpha w / |det|) r_k . (sum_m r_m ghat_m).
#include "common.cuh"

namespace {

template <int N>
struct HelmParams {
    double *y;
    const double *x;
    const double *coords;
    const int *map0;
    const int *map1;
    const int *collist;      // column indirection (subset / colour list) or NULL
    const int *off0;         // device, N^3 entries (zeros for non-extruded)
    const int *off1;         // device, 8 entries
    int ncols;               // number of columns to process
    int col0;                // first column (when collist == NULL)
    int nlay_items;          // layers to process per column
    int lay_first, lay_step; // layer = lay_first + lay_step * k
    int cdim;
    double alpha, beta;
    double B[N * N];         // B[q][a]
    double Dt[N * N];        // Dt[q][q']
    double wq[N];
    double xq[N];
};

__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(e, r, r);
    e = fma(-x, r, 1.0);
    r = fma(e, r, r);
    return r;
}

// out[i][j] = sum_k M(i,k) in[k][j];  M(i,k) = T ? M[k*N+i] : M[i*N+k]
template <int N, bool T>
__device__ __forceinline__ void apply_first(const double *M, const double (&in)[N][N],
                                            double (&out)[N][N])
{
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < N; k++) s = fma(T ? M[k * N + i] : M[i * N + k], in[k][j], s);
            out[i][j] = s;
        }
}

// out[i][j] = sum_k M(j,k) in[i][k]
template <int N, bool T>
__device__ __forceinline__ void apply_second(const double *M, const double (&in)[N][N],
                                             double (&out)[N][N])
{
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < N; k++) s = fma(T ? M[k * N + j] : M[j * N + k], in[i][k], s);
            out[i][j] = s;
        }
}

// Shared-memory tile used to re-orient slabs.  For N == 4 the (y, z) position
// is rotated by the cell's index in the warp so that both the layout-Z and the
// layout-Y access patterns touch all 32 banks exactly once per half-warp.
template <int N>
struct Tile {
    static constexpr int PAD = (N == 4) ? 0 : ((N * N * N) % 2 == 0 ? 2 : 1);
    static constexpr int STRIDE = N * N * N + PAD;
    double *base;
    int t, r, k[N];
    __device__ __forceinline__ Tile(double *warp_smem, int cw, int t_) : t(t_)
    {
        base = warp_smem + cw * STRIDE;
        if (N == 4) {
            r = (t_ + cw) & 3;
#pragma unroll
            for (int j = 0; j < N; j++) k[j] = (j + cw) & 3;
        } else {
            r = t_;
#pragma unroll
            for (int j = 0; j < N; j++) k[j] = j;
        }
    }
    // lane t == z holds a[x][y]
    __device__ __forceinline__ void store_Z(const double (&a)[N][N]) const
    {
#pragma unroll
        for (int x = 0; x < N; x++)
#pragma unroll
            for (int y = 0; y < N; y++) base[(x * N + k[y]) * N + r] = a[x][y];
    }
    __device__ __forceinline__ void load_Z(double (&a)[N][N]) const
    {
#pragma unroll
        for (int x = 0; x < N; x++)
#pragma unroll
            for (int y = 0; y < N; y++) a[x][y] = base[(x * N + k[y]) * N + r];
    }
    // lane t == y holds a[x][z]
    __device__ __forceinline__ void store_Y(const double (&a)[N][N]) const
    {
#pragma unroll
        for (int x = 0; x < N; x++)
#pragma unroll
            for (int z = 0; z < N; z++) base[(x * N + r) * N + k[z]] = a[x][z];
    }
    __device__ __forceinline__ void load_Y(double (&a)[N][N]) const
    {
#pragma unroll
        for (int x = 0; x < N; x++)
#pragma unroll
            for (int z = 0; z < N; z++) a[x][z] = base[(x * N + r) * N + k[z]];
    }
};

constexpr int WARPS_PER_CTA = 4;

template <int N, bool MASS, bool ATOMIC>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
helmholtz_action_kernel(const __grid_constant__ HelmParams<N> P)
{
    constexpr int CW = 32 / N;            // cells per warp
    constexpr int ND = N * N * N;
    constexpr int TS = Tile<N>::STRIDE;
    __shared__ double s_tile[WARPS_PER_CTA][CW * TS];
    __shared__ double s_coord[WARPS_PER_CTA][CW][24];
    __shared__ int s_off0[ND];

    for (int i = threadIdx.x; i < ND; i += blockDim.x) s_off0[i] = P.off0[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cw = lane / N, t = lane - cw * N;
    const bool lane_active = cw < CW;
    const int cwc = lane_active ? cw : 0;
    Tile<N> tile(s_tile[warp], cwc, t);
    double *sc = s_coord[warp][cwc];

    const int ncells = P.ncols * P.nlay_items;      // < 2^31, checked by the launcher
    const int nitems = (ncells + CW - 1) / CW;
    const double eta = P.xq[lane_active ? t : 0];
    const double wy_alpha = P.wq[lane_active ? t : 0] * P.alpha;
    const double wy_beta = P.wq[lane_active ? t : 0] * P.beta;

    for (int item = blockIdx.x * WARPS_PER_CTA + warp; item < nitems;
         item += gridDim.x * WARPS_PER_CTA) {
        // cells are numbered column-major: consecutive lanes-groups take
        // consecutive layers of one column (nlay_items == 1 for native hexes)
        const int lin = item * CW + cw;
        const bool valid = lane_active && lin < ncells;
        const int ci = valid ? lin / P.nlay_items : 0;
        const int kk = valid ? lin - ci * P.nlay_items : 0;
        const int layer = P.lay_first + P.lay_step * kk;
        const int col = P.collist ? __ldg(P.collist + ci) : (P.col0 + ci);
        const int *mrow = P.map0 + (long long)col * ND;

        // ---- stage the 8 vertex coordinates of each cell in shared memory
        __syncwarp();
        if (valid) {
            for (int i = t; i < 24; i += N) {
                int v = i / 3, a = i - v * 3;
                int g = __ldg(P.map1 + (long long)col * 8 + v) + __ldg(P.off1 + v) * layer;
                sc[i] = __ldg(P.coords + (long long)g * 3 + a);
            }
        }
        __syncwarp();
        // trilinear coefficients reduced at this lane's eta (see header comment)
        double A1[3], A3[3], A6[3], c2[3], c4[3], c5[3], c7[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double X000 = sc[0 * 3 + a], X001 = sc[1 * 3 + a], X010 = sc[2 * 3 + a],
                   X011 = sc[3 * 3 + a], X100 = sc[4 * 3 + a], X101 = sc[5 * 3 + a],
                   X110 = sc[6 * 3 + a], X111 = sc[7 * 3 + a];
            if (!valid) {   // keep idle lanes finite: unit cube
                X000 = 0; X001 = (a == 2); X010 = (a == 1); X011 = (a >= 1);
                X100 = (a == 0); X101 = (a != 1); X110 = (a != 2); X111 = 1;
            }
            double c1 = X100 - X000;
            c2[a] = X010 - X000;
            double c3 = X001 - X000;
            c4[a] = X110 - X100 - X010 + X000;
            c5[a] = X011 - X010 - X001 + X000;
            double c6 = X101 - X100 - X001 + X000;
            c7[a] = X111 - X110 - X101 - X011 + X100 + X010 + X001 - X000;
            A1[a] = fma(c4[a], eta, c1);
            A3[a] = fma(c5[a], eta, c3);
            A6[a] = fma(c7[a], eta, c6);
        }

        for (int comp = 0; comp < P.cdim; comp++) {
            // ---- gather, layout Z (lane t == a_z)
            double u[N][N];
#pragma unroll
            for (int x = 0; x < N; x++)
#pragma unroll
                for (int yy = 0; yy < N; yy++) {
                    const int loc = (x * N + yy) * N + t;
                    double v = 0.0;
                    if (valid) {
                        int g = __ldg(mrow + loc) + s_off0[loc] * layer;
                        v = __ldg(P.x + (long long)g * P.cdim + comp);
                    }
                    u[x][yy] = v;
                }
            // ---- forward: interpolate to the quadrature points
            double tmp[N][N], U[N][N], Gy[N][N];
            apply_first<N, false>(P.B, u, tmp);          // a_x -> q_x
            apply_second<N, false>(P.B, tmp, u);         // a_y -> q_y     u = w[qx][qy] @ a_z
            __syncwarp();
            tile.store_Z(u);
            __syncwarp();
            tile.load_Y(tmp);                            // tmp = w[qx][az] @ q_y
            apply_second<N, false>(P.B, tmp, U);         // a_z -> q_z     U[qx][qz] @ q_y
            __syncwarp();
            tile.store_Y(U);
            __syncwarp();
            tile.load_Z(tmp);                            // U[qx][qy] @ q_z
            apply_second<N, false>(P.Dt, tmp, u);        // d/d eta, still layout Z
            __syncwarp();
            tile.store_Z(u);
            __syncwarp();
            tile.load_Y(Gy);                             // Gy[qx][qz] @ q_y

            // ---- quadrature points (layout Y), fused with the x/z derivative
            //      and its transpose so only U, Gy and Vp stay live
            double Vp[N][N];
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int j = 0; j < N; j++) Vp[i][j] = 0.0;
#pragma unroll
            for (int qz = 0; qz < N; qz++) {
                const double zeta = P.xq[qz];
                double ca[3], pb[3], qb[3];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    ca[a] = fma(A6[a], zeta, A1[a]);       // dx/dxi
                    pb[a] = fma(c5[a], zeta, c2[a]);
                    qb[a] = fma(c7[a], zeta, c4[a]);
                }
                const double wyz_a = wy_alpha * P.wq[qz];
                const double wyz_b = wy_beta * P.wq[qz];
#pragma unroll
                for (int qx = 0; qx < N; qx++) {
                    const double xi = P.xq[qx];
                    double cb[3], cc[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        cb[a] = fma(qb[a], xi, pb[a]);     // dx/deta
                        cc[a] = fma(A6[a], xi, A3[a]);     // dx/dzeta
                    }
                    double gx = 0.0, gz = 0.0;
#pragma unroll
                    for (int q = 0; q < N; q++) {
                        gx = fma(P.Dt[qx * N + q], U[q][qz], gx);
                        gz = fma(P.Dt[qz * N + q], U[qx][q], gz);
                    }
                    const double gy = Gy[qx][qz];
                    // cofactor rows: r0 = b x c, r1 = c x a, r2 = a x b
                    double r0[3], r1[3], r2[3];
                    r0[0] = cb[1] * cc[2] - cb[2] * cc[1];
                    r0[1] = cb[2] * cc[0] - cb[0] * cc[2];
                    r0[2] = cb[0] * cc[1] - cb[1] * cc[0];
                    r1[0] = cc[1] * ca[2] - cc[2] * ca[1];
                    r1[1] = cc[2] * ca[0] - cc[0] * ca[2];
                    r1[2] = cc[0] * ca[1] - cc[1] * ca[0];
                    r2[0] = ca[1] * cb[2] - ca[2] * cb[1];
                    r2[1] = ca[2] * cb[0] - ca[0] * cb[2];
                    r2[2] = ca[0] * cb[1] - ca[1] * cb[0];
                    const double det = ca[0] * r0[0] + ca[1] * r0[1] + ca[2] * r0[2];
                    const double adet = fabs(det);
                    const double s = wyz_a * P.wq[qx] * fast_rcp(adet);
                    double h[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) h[a] = r0[a] * gx + r1[a] * gy + r2[a] * gz;
                    const double fx = s * (r0[0] * h[0] + r0[1] * h[1] + r0[2] * h[2]);
                    const double fy = s * (r1[0] * h[0] + r1[1] * h[1] + r1[2] * h[2]);
                    const double fz = s * (r2[0] * h[0] + r2[1] * h[1] + r2[2] * h[2]);
                    Gy[qx][qz] = fy;
#pragma unroll
                    for (int q = 0; q < N; q++) {
                        Vp[q][qz] = fma(P.Dt[qx * N + q], fx, Vp[q][qz]);
                        Vp[qx][q] = fma(P.Dt[qz * N + q], fz, Vp[qx][q]);
                    }
                    if (MASS) Vp[qx][qz] = fma(wyz_b * P.wq[qx] * adet, U[qx][qz], Vp[qx][qz]);
                }
            }

            // ---- backward
            __syncwarp();
            tile.store_Y(Gy);                            // Fy[qx][qz] @ q_y
            __syncwarp();
            tile.load_Z(tmp);                            // Fy[qx][qy] @ q_z
            apply_second<N, true>(P.Dt, tmp, u);         // Dt^T along eta
            __syncwarp();
            tile.store_Z(u);
            __syncwarp();
            tile.load_Y(tmp);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int j = 0; j < N; j++) Vp[i][j] += tmp[i][j];
            apply_second<N, true>(P.B, Vp, tmp);         // q_z -> a_z     W[qx][az] @ q_y
            __syncwarp();
            tile.store_Y(tmp);
            __syncwarp();
            tile.load_Z(u);                              // W[qx][qy] @ a_z
            apply_first<N, true>(P.B, u, tmp);           // q_x -> a_x
            apply_second<N, true>(P.B, tmp, u);          // q_y -> a_y     R[ax][ay] @ a_z

            // ---- scatter-add, layout Z
            if (valid) {
#pragma unroll
                for (int x = 0; x < N; x++)
#pragma unroll
                    for (int yy = 0; yy < N; yy++) {
                        const int loc = (x * N + yy) * N + t;
                        int g = __ldg(mrow + loc) + s_off0[loc] * layer;
                        double *dst = P.y + (long long)g * P.cdim + comp;
                        if (ATOMIC) atomicAdd(dst, u[x][yy]);
                        else *dst += u[x][yy];
                    }
            }
        }
    }
}

template <int N>
int launch_n(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay, const fdb_int *subset,
             double *y, const double *coords, const double *x, const fdb_int *map0,
             const fdb_int *map1)
{
    fdb::Context &c = fdb::ctx();
    HelmParams<N> P;
    P.y = y;
    P.x = x;
    P.coords = coords;
    P.map0 = map0;
    P.map1 = map1;
    P.off0 = k->d_off0;
    P.off1 = k->d_off1;
    P.cdim = k->desc.cdim;
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
    const bool mass = k->desc.beta != 0.0;
    constexpr int CW = 32 / N;
    auto grid_for = [&](long long ncols, int nitems_lay) {
        long long nitems = (ncols * nitems_lay + CW - 1) / CW;
        long long blocks = (nitems + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        long long cap = (long long)c.sm_count * 16;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        return (int)blocks;
    };
    if (k->desc.scatter == FDB_SCATTER_ATOMIC) {
        P.collist = subset;
        P.col0 = start;
        P.ncols = end - start;
        P.nlay_items = nlay;
        P.lay_first = 0;
        P.lay_step = 1;
        if (P.ncols <= 0 || nlay <= 0) return 0;
        int grid = grid_for(P.ncols, nlay);
        if (mass)
            helmholtz_action_kernel<N, true, true><<<grid, WARPS_PER_CTA * 32, 0, c.stream>>>(P);
        else
            helmholtz_action_kernel<N, false, true><<<grid, WARPS_PER_CTA * 32, 0, c.stream>>>(P);
        FDB_LAUNCH_CHECK();
        return 0;
    }
    // deterministic: one launch per (colour, layer parity); within a launch no
    // two cells share a dof, so plain read-modify-write is race free and the
    // summation order is fixed.
    if (subset) {
        fdb::set_error("coloured scatter does not support subsets yet");
        return 1;
    }
    for (int col = 0; col < k->ncolours; col++) {
        int cbeg = k->colour_start[col], cend = k->colour_start[col + 1];
        for (int par = 0; par < (nlay > 1 ? 2 : 1); par++) {
            P.collist = k->d_colour_cols + cbeg;
            P.col0 = 0;
            P.ncols = cend - cbeg;
            P.lay_first = par;
            P.lay_step = 2;
            P.nlay_items = (nlay - par + 1) / 2;
            if (P.ncols <= 0 || P.nlay_items <= 0) continue;
            int grid = grid_for(P.ncols, P.nlay_items);
            if (mass)
                helmholtz_action_kernel<N, true, false><<<grid, WARPS_PER_CTA * 32, 0, c.stream>>>(P);
            else
                helmholtz_action_kernel<N, false, false><<<grid, WARPS_PER_CTA * 32, 0, c.stream>>>(P);
            FDB_LAUNCH_CHECK();
        }
    }
    return 0;
}

}  // namespace

int fdb_launch_helmholtz_action(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay,
                                const fdb_int *subset, double *y, const double *coords,
                                const double *x, const fdb_int *map0, const fdb_int *map1)
{
    switch (k->n1d) {
    case 2: return launch_n<2>(k, start, end, nlay, subset, y, coords, x, map0, map1);
    case 3: return launch_n<3>(k, start, end, nlay, subset, y, coords, x, map0, map1);
    case 4: return launch_n<4>(k, start, end, nlay, subset, y, coords, x, map0, map1);
    case 5: return launch_n<5>(k, start, end, nlay, subset, y, coords, x, map0, map1);
    case 6: return launch_n<6>(k, start, end, nlay, subset, y, coords, x, map0, map1);
    }
    fdb::set_error("helmholtz action: degree %d not instantiated (1..5)", k->n1d - 1);
    return 1;
}
