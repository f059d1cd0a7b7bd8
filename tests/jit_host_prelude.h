// TEST INFRASTRUCTURE ONLY.  Host stand-in for the CUDA prelude of the generic
// wrapper builder (firedrake_b200/csrc/wrapper_jit.cu): tests/test_codegen.py
// cuts the generated source at the "prelude end" marker, prepends this file and
// compiles the flavour-independent BODY (local kernel + generated wrapper) with
// g++, so that the generated packing / unpacking / iteration code can be checked
// against the reference's golden arrays without a GPU.  "Threads" run one after
// another, so atomics and warp reductions degenerate to plain updates.  Nothing
// in firedrake_b200/ includes or executes this.
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscInt;
#define restrict __restrict__
#define FDB_DEVICE static inline
#define FDB_CONST static const
#define FDB_ABORT() abort()

struct FdbMatView {
    const long long *rowptr;
    const int *colidx;
    double *vals;
    const int *row_lg;
    const int *col_lg;
    int bs_r, bs_c;
};
struct FdbWrapParams {
    int start, end;
    int layer_lo, layer_hi;
    int bottom;
    int ncl;
    const int *subset;
    const int *col_layers;
    void *arg[16];
    const int *map[8];
    FdbMatView mat[4];
};

template <class T> FDB_DEVICE void fdb_atomic_add(T *p, T v) { *p += v; }
template <class T> FDB_DEVICE void fdb_atomic_min(T *p, T v) { if (v < *p) *p = v; }
template <class T> FDB_DEVICE void fdb_atomic_max(T *p, T v) { if (v > *p) *p = v; }
template <class T> FDB_DEVICE void fdb_reduce_add(T *g, T v, bool active) { if (active) *g += v; }
template <class T> FDB_DEVICE void fdb_reduce_min(T *g, T v, bool active) { if (active && v < *g) *g = v; }
template <class T> FDB_DEVICE void fdb_reduce_max(T *g, T v, bool active) { if (active && v > *g) *g = v; }

FDB_DEVICE void fdb_mat_set(const FdbMatView &m, int rnode, int a, int cnode, int b, double v, int insert)
{
    if (m.row_lg && m.row_lg[(long long)rnode * m.bs_r + a] < 0) return;
    if (m.col_lg && m.col_lg[(long long)cnode * m.bs_c + b] < 0) return;
    long long lo = m.rowptr[rnode], hi = m.rowptr[rnode + 1];
    if (hi <= lo) return;
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (m.colidx[mid] <= cnode) lo = mid; else hi = mid;
    }
    if (m.colidx[lo] != cnode) return;
    double *dst = m.vals + (lo * m.bs_r + a) * m.bs_c + b;
    if (insert) *dst = v; else *dst += v;
}

// the "launch": nthreads is rounded up to a multiple of 128 by the caller to
// exercise the inactive-thread path of the generated code
#define FDB_ENTRY(NAME, BODY)                                                  \
    extern "C" void NAME(const FdbWrapParams *p, long long nthreads)           \
    {                                                                          \
        for (long long t = 0; t < nthreads; ++t) BODY(*p, t);                  \
    }
