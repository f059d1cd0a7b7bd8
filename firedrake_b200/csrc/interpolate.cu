// Interpolation parloop: dual evaluation of a
// Q1 (x) P1 field at the nodes of Q_p (x) P_p -- what
// `Function(V_p).interpolate(w1)` / `interpolate(SpatialCoordinate(mesh), VectorFunctionSpace(mesh, "CG", p))`
// runs (reference firedrake/interpolation.py:977-1171: a parloop over cells
// with WRITE access on the target; every cell sharing a node writes the same
// value).  One thread per (cell, target node): gathers the 8 source values,
// evaluates the trilinear basis at the node's reference position (runtime
// table: 1-D node positions in dof numbering), plain store.
#include "common.cuh"

namespace {

struct InterpParams {
    double *out;            // (Nt, cdim)
    const double *src;      // (Ns, cdim)
    const int *map_t;       // (ncols, n^3)
    const int *map_s;       // (ncols, 8)
    const int *off_t;       // device, n^3 (zeros if not extruded)
    const int *off_s;       // device, 8
    int ncols, nlay, n, cdim;
    double nodes[FDB_MAX_1D];   // 1-D node positions on [0,1], dof numbering
};

__global__ void __launch_bounds__(256) interp_q1_kernel(const __grid_constant__ InterpParams P)
{
    const int nd = P.n * P.n * P.n;
    const long long total = (long long)P.ncols * P.nlay * nd;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long cell = i / nd;
        const int loc = (int)(i - cell * nd);
        const int col = (int)(cell / P.nlay), layer = (int)(cell - (long long)col * P.nlay);
        const int az = loc % P.n, ay = (loc / P.n) % P.n, ax = loc / (P.n * P.n);
        const double xi[3] = {P.nodes[ax], P.nodes[ay], P.nodes[az]};
        const long long gt = P.map_t[(long long)col * nd + loc] + (long long)P.off_t[loc] * layer;
        for (int c = 0; c < P.cdim; c++) {
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const int bx = b >> 2, by = (b >> 1) & 1, bz = b & 1;
                const double w = (bx ? xi[0] : 1.0 - xi[0]) * (by ? xi[1] : 1.0 - xi[1]) *
                                 (bz ? xi[2] : 1.0 - xi[2]);
                const long long gs = P.map_s[(long long)col * 8 + b] + (long long)P.off_s[b] * layer;
                v = fma(w, P.src[gs * P.cdim + c], v);
            }
            P.out[gt * P.cdim + c] = v;
        }
    }
}

}  // namespace

extern "C" int fdb_interpolate_q1(double *out, const double *src, const fdb_int *map_t,
                                  const fdb_int *map_s, const fdb_int *off_t_dev,
                                  const fdb_int *off_s_dev, fdb_int ncols, int nlay, int n1d, int cdim,
                                  const double *nodes_host)
{
    if (fdb::require_init()) return 1;
    if (n1d < 2 || n1d > FDB_MAX_1D || cdim < 1 || nlay < 1) {
        fdb::set_error("fdb_interpolate_q1: bad sizes");
        return 1;
    }
    InterpParams P;
    P.out = out;
    P.src = src;
    P.map_t = map_t;
    P.map_s = map_s;
    P.off_t = off_t_dev;
    P.off_s = off_s_dev;
    P.ncols = ncols;
    P.nlay = nlay;
    P.n = n1d;
    P.cdim = cdim;
    for (int i = 0; i < n1d; i++) P.nodes[i] = nodes_host[i];
    long long total = (long long)ncols * nlay * n1d * n1d * n1d;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)fdb::ctx().sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    interp_q1_kernel<<<(int)blocks, 256, 0, fdb::ctx().stream>>>(P);
    FDB_LAUNCH_CHECK();
    return 0;
}
