// Mesh-geometry queries used to choose kernel variants.
#include "common.cuh"

using namespace fdb;

namespace {

// a hex is a parallelepiped iff the xi*eta, eta*zeta, xi*zeta and xi*eta*zeta terms of its
// trilinear coordinate field vanish; EXACT zeros are required so that the affine kernel variant
// reproduces the general one to rounding
__global__ void k_cells_are_affine(const double *__restrict__ coords, const fdb_int *__restrict__ map1,
                                   fdb_int start, fdb_int ncols, int nlay, int o0, int o1, int o2, int o3,
                                   int o4, int o5, int o6, int o7, int *__restrict__ not_affine)
{
    const int off[8] = {o0, o1, o2, o3, o4, o5, o6, o7};
    const long long total = (long long)ncols * nlay;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const fdb_int c = start + (fdb_int)(i / nlay);
        const int l = (int)(i % nlay);
        const double *X[8];
        for (int v = 0; v < 8; v++) X[v] = coords + 3ll * (map1[(long long)c * 8 + v] + off[v] * l);
        bool bad = false;
        for (int a = 0; a < 3; a++) {
            // vertex v = (bx*2 + by)*2 + bz, as in action_hex.cu
            const double X000 = X[0][a], X001 = X[1][a], X010 = X[2][a], X011 = X[3][a], X100 = X[4][a],
                         X101 = X[5][a], X110 = X[6][a], X111 = X[7][a];
            bad |= (X110 - X100 - X010 + X000) != 0.0;
            bad |= (X011 - X010 - X001 + X000) != 0.0;
            bad |= (X101 - X100 - X001 + X000) != 0.0;
            bad |= (X111 - X110 - X101 - X011 + X100 + X010 + X001 - X000) != 0.0;
        }
        if (bad) *not_affine = 1;
    }
}

}  // namespace

extern "C" int fdb_cells_are_affine(const double *coords, const fdb_int *map1, const fdb_int *off1_host,
                                    fdb_int start, fdb_int end, int nlay, int *result)
{
    if (require_init()) return 1;
    if (!coords || !map1 || !result || end < start || nlay < 1) {
        set_error("fdb_cells_are_affine: bad arguments");
        return 1;
    }
    cudaStream_t st = ctx().stream;
    int *d_flag = nullptr;
    FDB_CUDA(cudaMalloc(&d_flag, sizeof(int)));
    FDB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    int o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (off1_host)
        for (int v = 0; v < 8; v++) o[v] = off1_host[v];
    const long long total = (long long)(end - start) * nlay;
    if (total > 0) {
        long long b = (total + 255) / 256;
        const long long cap = (long long)ctx().sm_count * 16;
        if (b > cap) b = cap;
        k_cells_are_affine<<<(int)b, 256, 0, st>>>(coords, map1, start, end - start, nlay, o[0], o[1], o[2], o[3],
                                                   o[4], o[5], o[6], o[7], d_flag);
        FDB_LAUNCH_CHECK();
    }
    int flag = 0;
    FDB_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_flag);
    *result = flag ? 0 : 1;
    return 0;
}
