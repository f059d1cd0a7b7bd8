// P1 triangles (BASELINE config 1: Poisson CG1 on UnitSquareMesh(64, 64)):
// alpha*inner(grad u, grad v)*dx + beta*u*v*dx on affine cells, bilinear form
// into CSR and its matrix-free action.  TSFC hoists the constant Jacobian of
// affine cells out of the quadrature loop (reference tsfc/fem.py:793-797); the
// quadrature table (basis values at the points, weights) is a runtime input.
// One thread per cell: these launches are latency/HBM-bound (~100 flop/cell).
#include "common.cuh"

namespace {

struct TriParams {
    const double *coords;   // (nv, 2)
    const double *x;
    double *y;
    const int *map;         // (ncells, 3)
    const int *subset;
    int start, end;
    int nq;
    double alpha, beta;
    double tab[3 * 16];     // tab[i*nq + q]
    double dtab[6];         // reference gradients
    double w[16];
    const long long *rowptr;
    const int *colidx;
    double *vals;
    const int *row_lg, *col_lg;
};

__device__ __forceinline__ void element_matrix(const TriParams &P, const int v[3], double A[3][3])
{
    double c[3][2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double2 X = *reinterpret_cast<const double2 *>(P.coords + 2 * (long long)v[i]);
        c[i][0] = X.x;
        c[i][1] = X.y;
    }
    double J[2][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int d = 0; d < 2; d++)
            J[a][d] = c[0][a] * P.dtab[0 * 2 + d] + c[1][a] * P.dtab[1 * 2 + d] + c[2][a] * P.dtab[2 * 2 + d];
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det, ad = fabs(det);
    const double K[2][2] = {{J[1][1] * id, -J[0][1] * id}, {-J[1][0] * id, J[0][0] * id}};
    double g[3][2], wsum = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int a = 0; a < 2; a++) g[i][a] = K[0][a] * P.dtab[i * 2] + K[1][a] * P.dtab[i * 2 + 1];
    for (int q = 0; q < P.nq; q++) wsum += P.w[q];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double m = 0.0;
            if (P.beta != 0.0)
                for (int q = 0; q < P.nq; q++) m += P.tab[i * P.nq + q] * P.tab[j * P.nq + q] * P.w[q];
            A[i][j] = ad * (P.alpha * wsum * (g[i][0] * g[j][0] + g[i][1] * g[j][1]) + P.beta * m);
        }
}

template <bool MATRIX>
__global__ void __launch_bounds__(128) tri_kernel(const __grid_constant__ TriParams P)
{
    for (int i = P.start + blockIdx.x * blockDim.x + threadIdx.x; i < P.end; i += gridDim.x * blockDim.x) {
        const int n = P.subset ? P.subset[i] : i;
        const int v[3] = {P.map[3 * (long long)n], P.map[3 * (long long)n + 1], P.map[3 * (long long)n + 2]};
        double A[3][3];
        element_matrix(P, v, A);
        if (MATRIX) {
#pragma unroll
            for (int a = 0; a < 3; a++) {
                int r = P.row_lg ? P.row_lg[v[a]] : v[a];
                if (r < 0) continue;
                const long long lo0 = P.rowptr[r], hi0 = P.rowptr[r + 1];
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    int cc = P.col_lg ? P.col_lg[v[b]] : v[b];
                    if (cc < 0) continue;
                    long long lo = lo0, hi = hi0;
                    while (hi - lo > 1) {
                        long long mid = (lo + hi) >> 1;
                        if (P.colidx[mid] <= cc) lo = mid; else hi = mid;
                    }
                    atomicAdd(P.vals + lo, A[a][b]);
                }
            }
        } else {
            const double xl[3] = {P.x[v[0]], P.x[v[1]], P.x[v[2]]};
#pragma unroll
            for (int a = 0; a < 3; a++)
                atomicAdd(P.y + v[a], A[a][0] * xl[0] + A[a][1] * xl[1] + A[a][2] * xl[2]);
        }
    }
}

}  // namespace

int fdb_launch_tri_p1(fdb_kernel_s *k, fdb_int start, fdb_int end, const fdb_int *subset, double *y,
                      const double *coords, const double *x, const fdb_int *map, fdb_mat_t mat)
{
    fdb::Context &c = fdb::ctx();
    if (end <= start) return 0;
    TriParams P;
    memset(&P, 0, sizeof(P));
    P.coords = coords;
    P.x = x;
    P.y = y;
    P.map = map;
    P.subset = subset;
    P.start = start;
    P.end = end;
    P.nq = k->desc.nq;
    P.alpha = k->desc.alpha;
    P.beta = k->desc.beta;
    for (int i = 0; i < 3 * P.nq; i++) P.tab[i] = k->desc.B[i];
    for (int i = 0; i < 6; i++) P.dtab[i] = k->desc.D[i];
    for (int i = 0; i < P.nq; i++) P.w[i] = k->desc.wq[i];
    long long blocks = ((long long)(end - start) + 127) / 128;
    long long cap = (long long)c.sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (mat) {
        fdb_mat_device_view(mat, &P.rowptr, &P.colidx, &P.vals, &P.row_lg, &P.col_lg);
        tri_kernel<true><<<(int)blocks, 128, 0, c.stream>>>(P);
    } else {
        tri_kernel<false><<<(int)blocks, 128, 0, c.stream>>>(P);
    }
    FDB_LAUNCH_CHECK();
    return 0;
}
