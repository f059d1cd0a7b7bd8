/* ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's (Firedrake/PyOP2/TSFC) assembly hot path
 * for the fixed form set of this repository.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (firedrake_b200 + libfdb200.so) never does.
 *
 * Why a restatement: the reference cannot run in this image -- UFL, FIAT/FInAT,
 * GEM, loopy, PETSc and MPI are absent, and the
 * arithmetic of its local kernels is generated at run time by those packages
 * (pyproject.toml:27-36 pins them to git main, i.e. unpinned).  What IS pinned:
 * the PyOP2 wrapper semantics (golden arrays of tests/pyop2/test_matrices.py,
 * test_indirect_loop.py, test_extrusion.py) and global invariants of
 * tests/firedrake/regression/*.  tests/test_oracle_pins.py checks this file
 * against those.  Element-matrix parity for TSFC-generated kernels is UNPINNED
 * in the reference itself (no golden element tensors exist there); here it is
 * anchored by analytic invariants (patch tests, null space, symmetry, exact
 * integrals) and an independent dense NumPy evaluation.
 *
 * Build flags mirror PyOP2's JIT on Linux/GCC: -O3 -march=native -ffast-math
 * -fPIC -shared (reference pyop2/compilation.py:341-363).
 */
#define _GNU_SOURCE
#include <sched.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    const double *B, *D;    /* (Q, N) 1-D trial/test tables, dof numbering   */
    const double *CB, *CD;  /* (Q, 2) 1-D P1 coordinate tables               */
    const double *wq;       /* (Q,)                                          */
} orc_tab;

/* A CSR matrix with sorted column indices per row: the cost model of PETSc's
 * MatSetValues on a preallocated AIJ matrix is a per-row search + add. */
typedef struct {
    int nrows;
    const int64_t *rowptr;
    const int *colidx;
    double *vals;
} orc_csr;

/* MatSetValuesLocal(..., ADD_VALUES): negative row/col indices are dropped
 * (this is how BC-masked LGMaps remove Dirichlet rows/columns; reference
 * firedrake/functionspaceimpl.py:854-926, pyop2/parloop.py:279-314). */
static int orc_mat_add_values(orc_csr *m, int nr, const int *rows, int nc,
                              const int *cols, const double *vals)
{
    for (int i = 0; i < nr; i++) {
        int r = rows[i];
        if (r < 0) continue;
        int64_t lo0 = m->rowptr[r], hi0 = m->rowptr[r + 1];
        for (int j = 0; j < nc; j++) {
            int c = cols[j];
            if (c < 0) continue;
            int64_t lo = lo0, hi = hi0;
            while (hi - lo > 1) {
                int64_t mid = (lo + hi) >> 1;
                if (m->colidx[mid] <= c) lo = mid; else hi = mid;
            }
            if (lo >= hi0 || m->colidx[lo] != c) return 1;  /* not preallocated */
            m->vals[lo] += vals[i * nc + j];
        }
    }
    return 0;
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define N 2
#define Q 2
#define FN(name) CAT(name, _p1)
#include "hex_kernels.inc"
#undef N
#undef Q
#undef FN
#define N 3
#define Q 3
#define FN(name) CAT(name, _p2)
#include "hex_kernels.inc"
#undef N
#undef Q
#undef FN
#define N 4
#define Q 4
#define FN(name) CAT(name, _p3)
#include "hex_kernels.inc"
#undef N
#undef Q
#undef FN
#define N 5
#define Q 5
#define FN(name) CAT(name, _p4)
#include "hex_kernels.inc"
#undef N
#undef Q
#undef FN
#define N 6
#define Q 6
#define FN(name) CAT(name, _p5)
#include "hex_kernels.inc"
#undef N
#undef Q
#undef FN

/* ------------------------------------------------------------------ exported */

int orc_wrap_action_extruded(int degree, int start, int end, const int *layers,
                             double *y, const double *coords, const double *x,
                             const int *map0, const int *off0,
                             const int *map1, const int *off1, int cdim,
                             const double *B, const double *D, const double *CB,
                             const double *CD, const double *wq,
                             double alpha, double beta)
{
    orc_tab t = {B, D, CB, CD, wq};
    switch (degree) {
    case 1: return wrap_action_extruded_p1(start, end, layers, y, coords, x, map0, off0, map1, off1, cdim, &t, alpha, beta);
    case 2: return wrap_action_extruded_p2(start, end, layers, y, coords, x, map0, off0, map1, off1, cdim, &t, alpha, beta);
    case 3: return wrap_action_extruded_p3(start, end, layers, y, coords, x, map0, off0, map1, off1, cdim, &t, alpha, beta);
    case 4: return wrap_action_extruded_p4(start, end, layers, y, coords, x, map0, off0, map1, off1, cdim, &t, alpha, beta);
    case 5: return wrap_action_extruded_p5(start, end, layers, y, coords, x, map0, off0, map1, off1, cdim, &t, alpha, beta);
    }
    return 2;
}

int orc_wrap_matrix_extruded(int degree, int start, int end, const int *layers,
                             int nrows, const int64_t *rowptr, const int *colidx, double *vals,
                             const double *coords,
                             const int *map0, const int *off0,
                             const int *map1, const int *off1,
                             const int *row_lgmap, const int *col_lgmap,
                             const double *B, const double *D, const double *CB,
                             const double *CD, const double *wq,
                             double alpha, double beta)
{
    orc_tab t = {B, D, CB, CD, wq};
    orc_csr m = {nrows, rowptr, colidx, vals};
    switch (degree) {
    case 1: return wrap_matrix_extruded_p1(start, end, layers, &m, coords, map0, off0, map1, off1, row_lgmap, col_lgmap, &t, alpha, beta);
    case 2: return wrap_matrix_extruded_p2(start, end, layers, &m, coords, map0, off0, map1, off1, row_lgmap, col_lgmap, &t, alpha, beta);
    case 3: return wrap_matrix_extruded_p3(start, end, layers, &m, coords, map0, off0, map1, off1, row_lgmap, col_lgmap, &t, alpha, beta);
    case 4: return wrap_matrix_extruded_p4(start, end, layers, &m, coords, map0, off0, map1, off1, row_lgmap, col_lgmap, &t, alpha, beta);
    case 5: return wrap_matrix_extruded_p5(start, end, layers, &m, coords, map0, off0, map1, off1, row_lgmap, col_lgmap, &t, alpha, beta);
    }
    return 2;
}

/* One element tensor, for direct kernel-level checks. */
int orc_cell_action(int degree, double *A, const double *coords, const double *w,
                    const double *B, const double *D, const double *CB,
                    const double *CD, const double *wq, double alpha, double beta)
{
    orc_tab t = {B, D, CB, CD, wq};
    switch (degree) {
    case 1: action_cell_p1(A, coords, w, &t, alpha, beta); return 0;
    case 2: action_cell_p2(A, coords, w, &t, alpha, beta); return 0;
    case 3: action_cell_p3(A, coords, w, &t, alpha, beta); return 0;
    case 4: action_cell_p4(A, coords, w, &t, alpha, beta); return 0;
    case 5: action_cell_p5(A, coords, w, &t, alpha, beta); return 0;
    }
    return 2;
}

int orc_cell_matrix(int degree, double *A, const double *coords,
                    const double *B, const double *D, const double *CB,
                    const double *CD, const double *wq, double alpha, double beta)
{
    orc_tab t = {B, D, CB, CD, wq};
    switch (degree) {
    case 1: matrix_cell_p1(A, coords, &t, alpha, beta); return 0;
    case 2: matrix_cell_p2(A, coords, &t, alpha, beta); return 0;
    case 3: matrix_cell_p3(A, coords, &t, alpha, beta); return 0;
    case 4: matrix_cell_p4(A, coords, &t, alpha, beta); return 0;
    case 5: matrix_cell_p5(A, coords, &t, alpha, beta); return 0;
    }
    return 2;
}

/* ---------------------------------------------------- P1 triangles (config 1)
 * Affine cells: TSFC hoists the constant Jacobian out of the quadrature loop
 * (reference tsfc/fem.py:793-797).  The quadrature/tabulation table is a
 * runtime input so that the PyOP2 golden test (tests/pyop2/test_matrices.py:
 * 166-330, which hard-codes an 8-digit 6-point table with basis order
 * {x, y, 1-x-y}) and the FIAT-ordered case share this code.
 *   tab[i*nq + q] : basis i at point q;  dtab[i*2 + d] : constant gradient
 *   w[q]          : weights on the reference triangle                        */

static void tri_jacobian(const double *c, const double *dtab, double J[2][2])
{
    for (int a = 0; a < 2; a++)
        for (int d = 0; d < 2; d++) {
            double s = 0;
            for (int i = 0; i < 3; i++) s += c[i * 2 + a] * dtab[i * 2 + d];
            J[a][d] = s;
        }
}

static void tri_mass_cell(double *A, const double *c, int nq, const double *tab,
                          const double *dtab, const double *w, int use_abs)
{
    double J[2][2];
    tri_jacobian(c, dtab, J);
    double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    if (use_abs) det = fabs(det);
    for (int q = 0; q < nq; q++)
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                A[i * 3 + j] += tab[i * nq + q] * tab[j * nq + q] * det * w[q];
}

static void tri_rhs_cell(double *b, const double *c, const double *f, int nq,
                         const double *tab, const double *dtab, const double *w,
                         int use_abs)
{
    double J[2][2];
    tri_jacobian(c, dtab, J);
    double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    if (use_abs) det = fabs(det);
    for (int q = 0; q < nq; q++) {
        double fq = 0;
        for (int i = 0; i < 3; i++) fq += f[i] * tab[i * nq + q];
        for (int i = 0; i < 3; i++) b[i] += tab[i * nq + q] * fq * det * w[q];
    }
}

static void tri_laplace_cell(double *A, const double *c, const double *dtab,
                             double area_ref, double beta, int nq,
                             const double *tab, const double *w)
{
    double J[2][2];
    tri_jacobian(c, dtab, J);
    double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double id = 1.0 / det;
    double K[2][2] = {{J[1][1] * id, -J[0][1] * id}, {-J[1][0] * id, J[0][0] * id}};
    double g[3][2];
    for (int i = 0; i < 3; i++)
        for (int a = 0; a < 2; a++)
            g[i][a] = K[0][a] * dtab[i * 2 + 0] + K[1][a] * dtab[i * 2 + 1];
    double ad = fabs(det);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            A[i * 3 + j] += area_ref * ad * (g[i][0] * g[j][0] + g[i][1] * g[j][1]);
    if (beta != 0.0)
        for (int q = 0; q < nq; q++)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                    A[i * 3 + j] += beta * tab[i * nq + q] * tab[j * nq + q] * ad * w[q];
}

/* Non-extruded wrapper, 2-form into CSR (reference builder.py:573-625). */
int orc_wrap_tri_matrix(int kind, int start, int end,
                        int nrows, const int64_t *rowptr, const int *colidx, double *vals,
                        const double *coords, const int *map,
                        const int *row_lgmap, const int *col_lgmap,
                        int nq, const double *tab, const double *dtab, const double *w,
                        int use_abs, double beta)
{
    orc_csr m = {nrows, rowptr, colidx, vals};
    double wsum = 0;
    for (int q = 0; q < nq; q++) wsum += w[q];
    for (int n = start; n < end; n++) {
        double c[6], A[9] = {0};
        int rows[3], cols[3];
        for (int i = 0; i < 3; i++) {
            int v = map[n * 3 + i];
            c[i * 2] = coords[v * 2];
            c[i * 2 + 1] = coords[v * 2 + 1];
            rows[i] = row_lgmap ? row_lgmap[v] : v;
            cols[i] = col_lgmap ? col_lgmap[v] : v;
        }
        if (kind == 0) tri_mass_cell(A, c, nq, tab, dtab, w, use_abs);
        else tri_laplace_cell(A, c, dtab, wsum, beta, nq, tab, w);
        if (orc_mat_add_values(&m, 3, rows, 3, cols, A)) return 1;
    }
    return 0;
}

/* Non-extruded wrapper, 1-form INC (reference builder.py:352-429). */
int orc_wrap_tri_rhs(int start, int end, double *b, const double *coords,
                     const double *f, const int *map, int nq, const double *tab,
                     const double *dtab, const double *w, int use_abs)
{
    for (int n = start; n < end; n++) {
        double c[6], fl[3], t0[3] = {0};
        for (int i = 0; i < 3; i++) {
            int v = map[n * 3 + i];
            c[i * 2] = coords[v * 2];
            c[i * 2 + 1] = coords[v * 2 + 1];
            fl[i] = f[v];
        }
        tri_rhs_cell(t0, c, fl, nq, tab, dtab, w, use_abs);
        for (int i = 0; i < 3; i++) b[map[n * 3 + i]] += t0[i];
    }
    return 0;
}

/* 1-form action of the P1 triangle operator: y += A_e x_e per cell. */
int orc_wrap_tri_action(int start, int end, double *y, const double *coords,
                        const double *x, const int *map, int nq, const double *tab,
                        const double *dtab, const double *w, double alpha, double beta)
{
    double wsum = 0;
    for (int q = 0; q < nq; q++) wsum += w[q];
    for (int n = start; n < end; n++) {
        double c[6], xl[3], A[9] = {0};
        for (int i = 0; i < 3; i++) {
            int v = map[n * 3 + i];
            c[i * 2] = coords[v * 2];
            c[i * 2 + 1] = coords[v * 2 + 1];
            xl[i] = x[v];
        }
        if (alpha != 0.0) tri_laplace_cell(A, c, dtab, wsum * alpha, beta, nq, tab, w);
        else tri_mass_cell(A, c, nq, tab, dtab, w, 1);
        for (int i = 0; i < 3; i++) {
            double s = 0;
            for (int j = 0; j < 3; j++) s += A[i * 3 + j] * xl[j];
            y[map[n * 3 + i]] += s;
        }
    }
    return 0;
}

/* ------------------------------------------------------------- generic loops
 * Indirect INC / READ / global reduction semantics (reference
 * tests/pyop2/test_indirect_loop.py:106-250: "r[map[n]] += x[n]" style
 * kernels), restated for the pin tests. */
int orc_indirect_inc(int start, int end, int arity, const int *map,
                     const double *edge_vals, double *node_vals)
{
    for (int n = start; n < end; n++)
        for (int i = 0; i < arity; i++) node_vals[map[n * arity + i]] += edge_vals[n];
    return 0;
}

/* ------------------------------------------------------------------ sparsity
 * nnz pattern = union over cells (and layers) of rowmap x colmap, diagonal
 * always allocated (reference pyop2/sparsity.pyx:106-160, 198-204; extruded
 * rows rmap[c][i] + off[i]*layer: :347-373).  Two passes: count, then fill. */
static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

int64_t orc_build_sparsity(int nrows, int ncells, int nlayers, int arity,
                           const int *map, const int *off,
                           int64_t *rowptr, int *colidx, int64_t capacity)
{
    /* gather (row, col) pairs, sort, unique.  Small meshes only. */
    int64_t npairs = (int64_t)ncells * nlayers * arity * arity + nrows;
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)npairs);
    if (!keys) return -1;
    int64_t k = 0;
    for (int c = 0; c < ncells; c++)
        for (int l = 0; l < nlayers; l++)
            for (int i = 0; i < arity; i++) {
                int64_t r = map[c * arity + i] + (off ? off[i] * l : 0);
                for (int j = 0; j < arity; j++) {
                    int64_t cc = map[c * arity + j] + (off ? off[j] * l : 0);
                    keys[k++] = r * (int64_t)nrows + cc;
                }
            }
    for (int r = 0; r < nrows; r++) keys[k++] = (int64_t)r * nrows + r;
    qsort(keys, (size_t)k, sizeof(int64_t), cmp_i64);
    int64_t nnz = 0;
    memset(rowptr, 0, sizeof(int64_t) * (size_t)(nrows + 1));
    for (int64_t i = 0; i < k; i++) {
        if (i > 0 && keys[i] == keys[i - 1]) continue;
        if (nnz < capacity) colidx[nnz] = (int)(keys[i] % nrows);
        rowptr[keys[i] / nrows + 1]++;
        nnz++;
    }
    for (int r = 0; r < nrows; r++) rowptr[r + 1] += rowptr[r];
    free(keys);
    return nnz;
}

/* ------------------------------------------------------------ timing baseline
 * The reference's parallel model is one sequential process per core, each
 * with its own ghosted local vectors (SURVEY.md section 8d "CPU baseline
 * timing").  Each worker runs the extruded action wrapper on ITS OWN local
 * problem; the caller times the whole call (max over workers). */
typedef struct {
    int start, end;
    const int *layers;
    double *y;
    const double *coords, *x;
    const int *map0, *off0, *map1, *off1;
} orc_worker;

int orc_action_workers(int degree, int nworkers, const orc_worker *w, int cdim,
                       const double *B, const double *D, const double *CB,
                       const double *CD, const double *wq, double alpha, double beta)
{
    int rc = 0;
#pragma omp parallel for schedule(static, 1) num_threads(nworkers) reduction(| : rc)
    for (int i = 0; i < nworkers; i++)
        rc |= orc_wrap_action_extruded(degree, w[i].start, w[i].end, w[i].layers, w[i].y,
                                       w[i].coords, w[i].x, w[i].map0, w[i].off0,
                                       w[i].map1, w[i].off1, cdim, B, D, CB, CD, wq,
                                       alpha, beta);
    return rc;
}

/* Timing arm of bench.py (--impl reference / cpu_baseline): the reference's
 * parallel model restated with the care an MPI launch gets for free --
 *   * one worker per core, PINNED (mpiexec --bind-to core),
 *   * every worker allocates and first-touches ITS OWN copies of its local
 *     arrays inside its thread (an MPI rank's memory is NUMA-local),
 *   * after the local loops, the contributions a worker accumulated on its
 *     ghost plane are added into the owning neighbour's plane (local_to_global
 *     with INC, reference pyop2/parloop.py:255-260) -- slab decomposition, so a
 *     worker has one ghost plane (x = high face), owned by worker i+1.
 * Every worker gets the same template slab (same sizes as a real slab of the
 * partitioned mesh); times[r] = wall time of pass r (barrier to barrier: zero y,
 * local loop, ghost reduce), i.e. the max over workers.  reps passes after one
 * untimed warm-up pass.  Returns 0, or nonzero on allocation / kernel failure. */
int orc_action_bench(int degree, int nworkers, int reps, const int *cpu_ids,
                     int ncols, const int *layers, int64_t ndof, int64_t nvert,
                     const double *coords, const double *x, const int *map0, int arity0,
                     const int *off0, const int *map1, const int *off1,
                     int nface, const int *face_hi, const int *face_lo,
                     const double *B, const double *D, const double *CB, const double *CD,
                     const double *wq, double alpha, double beta, double *times,
                     double *checksum)
{
    int rc = 0;
    double **ys = (double **)calloc((size_t)nworkers, sizeof(double *));
    if (!ys) return 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nworkers) reduction(| : rc)
    {
        const int w = omp_get_thread_num();
        if (cpu_ids) {
            cpu_set_t set;
            CPU_ZERO(&set);
            CPU_SET(cpu_ids[w], &set);
            sched_setaffinity(0, sizeof(set), &set);
        }
        /* private, first-touched copies */
        double *yl = (double *)malloc(sizeof(double) * (size_t)ndof);
        double *xl = (double *)malloc(sizeof(double) * (size_t)ndof);
        double *cl = (double *)malloc(sizeof(double) * (size_t)nvert * 3);
        int *m0 = (int *)malloc(sizeof(int) * (size_t)ncols * arity0);
        int *m1 = (int *)malloc(sizeof(int) * (size_t)ncols * 8);
        int bad = !yl || !xl || !cl || !m0 || !m1;
        if (!bad) {
            memcpy(xl, x, sizeof(double) * (size_t)ndof);
            memcpy(cl, coords, sizeof(double) * (size_t)nvert * 3);
            memcpy(m0, map0, sizeof(int) * (size_t)ncols * arity0);
            memcpy(m1, map1, sizeof(int) * (size_t)ncols * 8);
            memset(yl, 0, sizeof(double) * (size_t)ndof);
        }
        ys[w] = yl;
        rc |= bad;
#pragma omp barrier
        for (int r = -1; r < reps && !rc; r++) {
#pragma omp barrier
            double t0 = omp_get_wtime();
            memset(yl, 0, sizeof(double) * (size_t)ndof);   /* assembler zeroes the tensor */
            rc |= orc_wrap_action_extruded(degree, 0, ncols, layers, yl, cl, xl, m0, off0, m1,
                                           off1, 1, B, D, CB, CD, wq, alpha, beta);
#pragma omp barrier
            /* ghost plane of worker w-1 -> my owned low plane */
            if (w > 0 && ys[w - 1]) {
                const double *yn = ys[w - 1];
                for (int i = 0; i < nface; i++) yl[face_lo[i]] += yn[face_hi[i]];
            }
#pragma omp barrier
            if (w == 0 && r >= 0) times[r] = omp_get_wtime() - t0;
        }
        if (w == nworkers / 2 && checksum && !bad) {
            double s = 0.0;
            for (int64_t i = 0; i < ndof; i++) s += yl[i];
            *checksum = s;
        }
#pragma omp barrier
        free(yl); free(xl); free(cl); free(m0); free(m1);
    }
#else
    rc = 1;
#endif
    free(ys);
    return rc;
}

/* Full-size parity helper: the same sequential wrapper, run by several threads on
 * disjoint BANDS of base cells.  Cells are grouped by the caller into bands (slabs
 * of base-cell columns at least one cell wide) such that two bands of the same
 * parity share no dof; even bands run concurrently, then odd bands, all adding
 * into the one shared y -- no atomics, and within a band the reference's
 * sequential order.  A band is a list of runs [run_start, run_end) of the cell
 * order (runs of band b: run_first[b] .. run_first[b+1]). */
int orc_action_bands(int degree, int nbands, const int *run_first, const int *run_start,
                     const int *run_end, const int *layers, double *y, const double *coords,
                     const double *x, const int *map0, const int *off0, const int *map1,
                     const int *off1, int cdim, const double *B, const double *D,
                     const double *CB, const double *CD, const double *wq, double alpha,
                     double beta, int nthreads)
{
    int rc = 0;
    for (int parity = 0; parity < 2; parity++) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(| : rc)
        for (int b = parity; b < nbands; b += 2)
            for (int r = run_first[b]; r < run_first[b + 1]; r++)
                rc |= orc_wrap_action_extruded(degree, run_start[r], run_end[r], layers, y, coords,
                                               x, map0, off0, map1, off1, cdim, B, D, CB, CD,
                                               wq, alpha, beta);
    }
    return rc;
}

/* Vector algebra of the host Krylov baseline (bench.py cpu_cg_baseline): what PETSc's Vec
 * operations do across the ranks of the reference's MPI run, here across OpenMP threads. */
void orc_vec_axpy(int64_t n, double a, const double *x, double *y)      /* y += a x */
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) y[i] += a * x[i];
}

void orc_vec_aypx(int64_t n, double a, const double *x, double *y)      /* y = x + a y */
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) y[i] = x[i] + a * y[i];
}

double orc_vec_dot(int64_t n, const double *x, const double *y)
{
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < n; i++) s += x[i] * y[i];
    return s;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
