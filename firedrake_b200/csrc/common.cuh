// Shared internals of libfdb200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/fdb200.h"

namespace fdb {

struct Context {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    double *reduce_scratch = nullptr;   // device scratch for reductions
    double *reduce_host = nullptr;      // pinned
    int *work_counter = nullptr;        // device work-queue counter of the persistent kernels
};

Context &ctx();
void set_error(const char *fmt, ...);
int require_init();

#define FDB_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            fdb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
                           cudaGetErrorString(e_));                                 \
            return 1;                                                               \
        }                                                                           \
    } while (0)

#define FDB_LAUNCH_CHECK()                                                          \
    do {                                                                            \
        fdb::ctx().launches++;                                                      \
        cudaError_t e_ = cudaGetLastError();                                        \
        if (e_ != cudaSuccess) {                                                    \
            fdb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,        \
                           cudaGetErrorString(e_));                                 \
            return 1;                                                               \
        }                                                                           \
    } while (0)

}  // namespace fdb

extern int fdb_opt_matrix_kernel;   // fdb_set_option("matrix_kernel", ...)

struct fdb_jit_s;   // NVRTC-compiled generic wrapper (wrapper_jit.cu)

// kernel object behind fdb_kernel_t
struct fdb_kernel_s {
    fdb_kernel_desc desc;
    int n1d;                 // p + 1
    int arity;               // dofs per cell of the argument map
    fdb_int *d_off0 = nullptr;   // device copies of the layer offsets
    fdb_int *d_off1 = nullptr;
    fdb_int h_off0[512];
    fdb_int h_off1[8];
    double Dt[FDB_MAX_1D * FDB_MAX_1D];   // collocated derivative D * B^{-1}
    // colouring plan for FDB_SCATTER_COLOURED, built lazily per map
    const void *colour_map_key = nullptr;   // plan is valid for (map pointer, generation, end)
    uint64_t colour_map_gen = 0;
    fdb_int colour_end = 0;
    fdb_int *d_colour_cols = nullptr;     // columns sorted by colour
    int ncolours = 0;
    fdb_int colour_start[65];
    // dense B^T D B path (bdb_matrix.cu): tabulated reference gradients / values, built lazily
    double *d_bdb_table = nullptr;
    // non-NULL: this handle is a generated wrapper around an arbitrary local kernel
    fdb_jit_s *jit = nullptr;
};

int fdb_jit_call(fdb_kernel_s *k, const fdb_call_args *a);
void fdb_jit_destroy(fdb_jit_s *j);

extern "C" int fdb_mirror_set_version(const void *host, uint64_t version);
bool fdb_mirror_is_current(const void *host, size_t nbytes, uint64_t version);
void fdb_mirror_new_epoch();   // start of a kernel call: mirrors touched from now on are pinned

typedef struct fdb_mat_s *fdb_mat_t;
int fdb_mat_device_view(fdb_mat_t m, const long long **rowptr, const fdb_int **colidx, double **vals,
                        const fdb_int **row_lg, const fdb_int **col_lg);
int fdb_mat_rank_table(fdb_mat_t m, const unsigned short **rank, int *nvar);
int fdb_mat_block_size(fdb_mat_t m, int *bs);
int fdb_mat_scalar_view_begin(fdb_mat_t blocked, fdb_mat_t *view);
int fdb_mat_scalar_view_end(fdb_mat_t blocked, fdb_mat_t view);
int fdb_launch_helmholtz_matrix(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay,
                                const fdb_int *subset, fdb_mat_t mat, const double *coords,
                                const fdb_int *map0, const fdb_int *map1, double *diag_out);

int fdb_launch_helmholtz_matrix_dmma(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay,
                                     const fdb_int *subset, fdb_mat_t mat, const double *coords,
                                     const fdb_int *map0, const fdb_int *map1);

int fdb_launch_tri_p1(fdb_kernel_s *k, fdb_int start, fdb_int end, const fdb_int *subset, double *y,
                      const double *coords, const double *x, const fdb_int *map, fdb_mat_t mat);
int fdb_launch_dg_advection(fdb_kernel_s *k, fdb_int start, fdb_int end, const fdb_int *subset,
                            double *out, const double *coords, const double *q, const double *u,
                            const double *consts_host, const unsigned *facet, const fdb_int *dgmap,
                            const fdb_int *cgmap, const fdb_int *nbr);

int fdb_launch_q1_action(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay, const fdb_int *subset,
                         double *y, const double *coords, const double *x, const fdb_int *map0,
                         const fdb_int *map1);

int fdb_launch_q2_action(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay, const fdb_int *subset,
                         double *y, const double *coords, const double *x, const fdb_int *map0,
                         const fdb_int *map1);

// launchers implemented in the kernel translation units
int fdb_launch_helmholtz_action(fdb_kernel_s *k, fdb_int start, fdb_int end, int nlay,
                                const fdb_int *subset, double *y, const double *coords,
                                const double *x, const fdb_int *map0, const fdb_int *map1);
