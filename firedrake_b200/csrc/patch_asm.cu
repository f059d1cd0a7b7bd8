// Batched small dense linear algebra on the device: the additive Schwarz patch smoother of
// TinyASM.
//
// Reference semantics (tinyasm/tinyasm.cpp:27-120, class BlockJacobi):
//   updateValuesPerBlock(P): for every patch p with dof list d_p, extract the dense block
//                            P[d_p, d_p] (MatCreateSubMatrices + MatConvert(MATDENSE)) and replace
//                            it by its inverse (LAPACK getrf/getri, "mymatinvert");
//   solve(b, x):             x[d_p] += inv(P[d_p, d_p]) b[d_p]   for every patch (additive).
// Here: one gather kernel over all (patch, i, j) entries (binary search in the sorted CSR rows),
// one CTA per patch for an in-place Gauss-Jordan inversion with partial pivoting (in shared
// memory when the block fits, else in place in global memory), and one warp-per-row dense
// mat-vec per patch with RED.ADD.F64 into x (patches overlap).  Patch sizes are arbitrary; the
// blocks are stored back to back, row-major.
#include <stdlib.h>

#include <vector>

#include "common.cuh"

using namespace fdb;

struct fdb_asm_s {
    int npatch = 0;
    int max_n = 0;
    long long total_dofs = 0, total_entries = 0;
    long long *d_ptr = nullptr;       // [npatch + 1] offsets into d_dofs
    long long *d_mptr = nullptr;      // [npatch + 1] offsets into d_inv (sum of n_p^2)
    fdb_int *d_dofs = nullptr;
    double *d_inv = nullptr;
    int *d_info = nullptr;            // number of singular patches met by the last update
};

namespace {

__global__ void k_extract(int npatch, const long long *__restrict__ ptr, const long long *__restrict__ mptr,
                          const fdb_int *__restrict__ dofs, const long long *__restrict__ rowptr,
                          const fdb_int *__restrict__ colidx, const double *__restrict__ vals, int bs,
                          double *__restrict__ out)
{
    // one CTA per patch, threads stride over its n^2 entries; dof index = node * bs + component
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        const int n = (int)(ptr[p + 1] - ptr[p]);
        const fdb_int *d = dofs + ptr[p];
        double *A = out + mptr[p];
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e / n, j = e - i * n;
            const fdb_int ri = d[i], cj = d[j];
            const fdb_int rn = ri / bs, ra = ri - rn * bs, cn = cj / bs, ca = cj - cn * bs;
            long long lo = rowptr[rn], hi = rowptr[rn + 1];
            double v = 0.0;
            if (hi > lo) {
                while (hi - lo > 1) {
                    const long long mid = (lo + hi) >> 1;
                    if (colidx[mid] <= cn) lo = mid; else hi = mid;
                }
                if (colidx[lo] == cn) v = vals[lo * bs * bs + ra * bs + ca];
            }
            A[e] = v;
        }
    }
}

// In-place Gauss-Jordan inversion with partial (row) pivoting of the n x n row-major block A.
// All threads of the CTA cooperate; `piv` (n ints) and `red` (blockDim doubles + ints) in shared memory.
__device__ void gauss_jordan(double *A, int n, int *piv, double *red_v, int *red_i, int *singular)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 0; k < n; k++) {
        // pivot search in column k, rows k..n-1
        double best = -1.0;
        int bi = k;
        for (int i = k + tid; i < n; i += nt) {
            const double v = fabs(A[i * n + k]);
            if (v > best) { best = v; bi = i; }
        }
        red_v[tid] = best;
        red_i[tid] = bi;
        __syncthreads();
        for (int s = nt >> 1; s > 0; s >>= 1) {
            if (tid < s && (red_v[tid + s] > red_v[tid] ||
                            (red_v[tid + s] == red_v[tid] && red_i[tid + s] < red_i[tid]))) {
                red_v[tid] = red_v[tid + s];
                red_i[tid] = red_i[tid + s];
            }
            __syncthreads();
        }
        const int r = red_i[0];
        const double pmax = red_v[0];
        __syncthreads();
        if (tid == 0) {
            piv[k] = r;
            if (!(pmax > 0.0)) *singular = 1;
        }
        if (r != k)
            for (int j = tid; j < n; j += nt) {
                const double t = A[k * n + j];
                A[k * n + j] = A[r * n + j];
                A[r * n + j] = t;
            }
        __syncthreads();
        const double pinv = 1.0 / A[k * n + k];
        __syncthreads();
        // scale the pivot row (its diagonal entry becomes 1/pivot)
        for (int j = tid; j < n; j += nt) A[k * n + j] = (j == k) ? pinv : A[k * n + j] * pinv;
        __syncthreads();
        // eliminate column k from every other row; the column is saved in red_v chunks via registers
        for (int e = tid; e < n * n; e += nt) {
            const int i = e / n, j = e - i * n;
            if (i == k || j == k) continue;
            A[e] = fma(-A[i * n + k], A[k * n + j], A[e]);
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt)
            if (i != k) A[i * n + k] = -A[i * n + k] * pinv;
        __syncthreads();
    }
    // undo the row interchanges as column interchanges, in reverse order
    for (int k = n - 1; k >= 0; k--) {
        const int r = piv[k];
        if (r != k)
            for (int i = tid; i < n; i += nt) {
                const double t = A[i * n + k];
                A[i * n + k] = A[i * n + r];
                A[i * n + r] = t;
            }
        __syncthreads();
    }
}

// shared-memory layout of k_invert: reduction scratch, pivots, then the block itself
__host__ __device__ inline size_t invert_head_bytes(int max_n)
{
    return ((256 * 12 + 4 * (size_t)max_n + 15) / 16) * 16;
}

__global__ void __launch_bounds__(256) k_invert(int npatch, const long long *__restrict__ ptr,
                                                const long long *__restrict__ mptr, double *__restrict__ blocks,
                                                int max_n, int smem_max_n, int *__restrict__ info)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *red_v = reinterpret_cast<double *>(smem_raw);
    int *red_i = reinterpret_cast<int *>(red_v + 256);
    int *piv = red_i + 256;                                      // [max_n]
    double *As = reinterpret_cast<double *>(smem_raw + invert_head_bytes(max_n));
    __shared__ int singular;
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        const int n = (int)(ptr[p + 1] - ptr[p]);
        double *Ag = blocks + mptr[p];
        if (threadIdx.x == 0) singular = 0;
        __syncthreads();
        if (n <= smem_max_n) {
            for (int e = threadIdx.x; e < n * n; e += blockDim.x) As[e] = Ag[e];
            __syncthreads();
            gauss_jordan(As, n, piv, red_v, red_i, &singular);
            for (int e = threadIdx.x; e < n * n; e += blockDim.x) Ag[e] = As[e];
        } else {
            gauss_jordan(Ag, n, piv, red_v, red_i, &singular);
        }
        __syncthreads();
        if (threadIdx.x == 0 && singular) atomicAdd(info, 1);
        __syncthreads();
    }
}

// x[d_p] += Ainv_p b[d_p]: one CTA per patch, one warp per output row
__global__ void __launch_bounds__(256) k_apply(int npatch, const long long *__restrict__ ptr,
                                               const long long *__restrict__ mptr, const fdb_int *__restrict__ dofs,
                                               const double *__restrict__ inv, const double *__restrict__ b,
                                               double *__restrict__ x)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *bl = reinterpret_cast<double *>(smem_raw);            // [max_n]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        const int n = (int)(ptr[p + 1] - ptr[p]);
        const fdb_int *d = dofs + ptr[p];
        const double *A = inv + mptr[p];
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += blockDim.x) bl[j] = b[d[j]];
        __syncthreads();
        for (int i = warp; i < n; i += nw) {
            double s = 0.0;
            for (int j = lane; j < n; j += 32) s = fma(A[(long long)i * n + j], bl[j], s);
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) atomicAdd(x + d[i], s);
        }
    }
}

}  // namespace

extern "C" {

int fdb_asm_create(int npatch, const long long *patch_ptr_host, const fdb_int *patch_dofs_host, fdb_asm_t *out)
{
    if (require_init()) return 1;
    if (npatch < 0 || !patch_ptr_host || !out) {
        set_error("fdb_asm_create: bad arguments");
        return 1;
    }
    fdb_asm_s *a = new fdb_asm_s;
    a->npatch = npatch;
    std::vector<long long> mptr(npatch + 1, 0);
    for (int p = 0; p < npatch; p++) {
        const long long n = patch_ptr_host[p + 1] - patch_ptr_host[p];
        if (n < 0) {
            set_error("fdb_asm_create: patch offsets must be non-decreasing");
            delete a;
            return 1;
        }
        if (n > a->max_n) a->max_n = (int)n;
        mptr[p + 1] = mptr[p] + n * n;
    }
    a->total_dofs = patch_ptr_host[npatch];
    a->total_entries = mptr[npatch];
    cudaStream_t st = ctx().stream;
    FDB_CUDA(cudaMalloc(&a->d_ptr, sizeof(long long) * (npatch + 1)));
    FDB_CUDA(cudaMalloc(&a->d_mptr, sizeof(long long) * (npatch + 1)));
    FDB_CUDA(cudaMalloc(&a->d_dofs, sizeof(fdb_int) * (size_t)(a->total_dofs + 1)));
    FDB_CUDA(cudaMalloc(&a->d_inv, sizeof(double) * (size_t)(a->total_entries + 1)));
    FDB_CUDA(cudaMalloc(&a->d_info, sizeof(int)));
    FDB_CUDA(cudaMemcpyAsync(a->d_ptr, patch_ptr_host, sizeof(long long) * (npatch + 1), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(a->d_mptr, mptr.data(), sizeof(long long) * (npatch + 1), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(a->d_dofs, patch_dofs_host, sizeof(fdb_int) * (size_t)a->total_dofs,
                             cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    *out = a;
    return 0;
}

int fdb_asm_destroy(fdb_asm_t a)
{
    if (!a) return 0;
    if (ctx().ready) {
        cudaStreamSynchronize(ctx().stream);
        cudaFree(a->d_ptr);
        cudaFree(a->d_mptr);
        cudaFree(a->d_dofs);
        cudaFree(a->d_inv);
        cudaFree(a->d_info);
    }
    delete a;
    return 0;
}

// updateValuesPerBlock: extract the patch blocks of `mat` and invert them; *nsingular (may be NULL)
// receives the number of patches whose elimination met a zero pivot.
int fdb_asm_update(fdb_asm_t a, fdb_mat_t mat, int *nsingular)
{
    if (require_init()) return 1;
    if (a->npatch == 0) return 0;
    cudaStream_t st = ctx().stream;
    const long long *rowptr;
    const fdb_int *colidx, *rlg, *clg;
    double *vals;
    if (fdb_mat_device_view(mat, &rowptr, &colidx, &vals, &rlg, &clg)) return 1;
    int bs = 1;
    fdb_mat_block_size(mat, &bs);
    const int grid = a->npatch < ctx().sm_count * 8 ? a->npatch : ctx().sm_count * 8;
    k_extract<<<grid, 256, 0, st>>>(a->npatch, a->d_ptr, a->d_mptr, a->d_dofs, rowptr, colidx, vals, bs, a->d_inv);
    FDB_LAUNCH_CHECK();
    FDB_CUDA(cudaMemsetAsync(a->d_info, 0, sizeof(int), st));
    // blocks up to smem_max_n are inverted in shared memory
    const size_t head = invert_head_bytes(a->max_n);
    int smem_max_n = a->max_n;
    const size_t cap = 200 * 1024;
    while (smem_max_n > 0 && head + sizeof(double) * (size_t)smem_max_n * smem_max_n > cap) smem_max_n--;
    const size_t smem = head + sizeof(double) * (size_t)smem_max_n * smem_max_n;
    static size_t configured = 0;
    if (smem > configured) {
        FDB_CUDA(cudaFuncSetAttribute(k_invert, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int ginv = a->npatch < ctx().sm_count * 4 ? a->npatch : ctx().sm_count * 4;
    k_invert<<<ginv, 256, smem, st>>>(a->npatch, a->d_ptr, a->d_mptr, a->d_inv, a->max_n, smem_max_n, a->d_info);
    FDB_LAUNCH_CHECK();
    if (nsingular) {
        FDB_CUDA(cudaMemcpyAsync(nsingular, a->d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

// solve: x[d_p] += inv(P[d_p, d_p]) b[d_p] over all patches (device pointers; x is INCREMENTED)
int fdb_asm_apply(fdb_asm_t a, const double *b, double *x)
{
    if (require_init()) return 1;
    if (a->npatch == 0) return 0;
    const int grid = a->npatch < ctx().sm_count * 8 ? a->npatch : ctx().sm_count * 8;
    k_apply<<<grid, 256, sizeof(double) * (size_t)(a->max_n + 1), ctx().stream>>>(
        a->npatch, a->d_ptr, a->d_mptr, a->d_dofs, a->d_inv, b, x);
    FDB_LAUNCH_CHECK();
    return 0;
}

// the inverted blocks, back to back (row-major), for tests
int fdb_asm_get_blocks(fdb_asm_t a, double *out_host)
{
    if (require_init()) return 1;
    FDB_CUDA(cudaMemcpyAsync(out_host, a->d_inv, sizeof(double) * (size_t)a->total_entries, cudaMemcpyDeviceToHost,
                             ctx().stream));
    FDB_CUDA(cudaStreamSynchronize(ctx().stream));
    return 0;
}

}  // extern "C"
