/* ORACLE -- TEST INFRASTRUCTURE (see oracle.c header).
 *
 * CPU restatement of the three kernels TSFC generates for the right-hand side
 * of the DG advection demo (reference demos/DG_advection/DG_advection.py.rst:182-217)
 *
 *   L1 = dtc*( q*div(phi*u)*dx
 *            - conditional(dot(u,n) < 0, phi*dot(u,n)*q_in, 0)*ds
 *            - conditional(dot(u,n) > 0, phi*dot(u,n)*q,    0)*ds
 *            - (phi('+') - phi('-'))*(un('+')*q('+') - un('-')*q('-'))*dS ),
 *   un = 0.5*(dot(u,n) + abs(dot(u,n)))
 *
 * on quadrilaterals: q, phi in DQ1 (4 dofs per cell), u in vector CG1, Q1
 * coordinates.  Kernel ABI (reference tsfc/kernel_interface/firedrake_loopy.py:
 * 317-381, tsfc/kernel_interface/common.py:518-522, 579-592): facet kernels get
 * the local facet number(s); interior-facet coefficient arrays hold the '+'
 * cell's dofs then the '-' cell's.  Reference-quad facets: 0: x=0, 1: x=1,
 * 2: y=0, 3: y=1 (reference firedrake/cython/dmcommon.pyx:1496-1499).
 * Local dof index of a tensor-product element: ax*2 + ay.
 * Tables (runtime inputs): Bq[nq][2] / Dq[nq][2] DQ1 1-D basis at the 1-D
 * quadrature points, Bq0[2] / Bq1[2] its values at the interval ends,
 * CB/CD the P1 basis, wq, xq.
 */
#include <math.h>

typedef struct {
    int nq;
    const double *Bq, *Dq;     /* (nq, 2) DQ1 */
    const double *Bend;        /* (2, 2): DQ1 basis at x=0 and x=1 */
    const double *CB, *CD;     /* (nq, 2) P1 */
    const double *wq, *xq;
} dg_tab;

static void q1_point(const double *c, double x, double y, double J[2][2])
{
    /* Jacobian of the Q1 map at (x, y); c = coords (4 vertices x 2), local ax*2+ay */
    const double bx[2] = {1 - x, x}, by[2] = {1 - y, y}, dx[2] = {-1, 1}, dy[2] = {-1, 1};
    J[0][0] = J[0][1] = J[1][0] = J[1][1] = 0;
    for (int ax = 0; ax < 2; ax++)
        for (int ay = 0; ay < 2; ay++)
            for (int a = 0; a < 2; a++) {
                double X = c[(ax * 2 + ay) * 2 + a];
                J[a][0] += X * dx[ax] * by[ay];
                J[a][1] += X * bx[ax] * dy[ay];
            }
}

static void p1_eval(const double *v, double x, double y, double out[2])
{
    const double bx[2] = {1 - x, x}, by[2] = {1 - y, y};
    out[0] = out[1] = 0;
    for (int ax = 0; ax < 2; ax++)
        for (int ay = 0; ay < 2; ay++)
            for (int a = 0; a < 2; a++) out[a] += v[(ax * 2 + ay) * 2 + a] * bx[ax] * by[ay];
}

/* dq basis on an arbitrary 1-D point through the two tabulated nodes: the DQ1
 * basis is linear, defined by its end values Bend[0][i], Bend[1][i] */
static double dq1d(const dg_tab *t, int i, double x) { return t->Bend[0 * 2 + i] * (1 - x) + t->Bend[1 * 2 + i] * x; }
static double ddq1d(const dg_tab *t, int i) { return t->Bend[1 * 2 + i] - t->Bend[0 * 2 + i]; }

void dg_cell_kernel(double *A, const double *c, const double *q, const double *u, const dg_tab *t,
                    double dt)
{
    for (int qx = 0; qx < t->nq; qx++)
        for (int qy = 0; qy < t->nq; qy++) {
            double x = t->xq[qx], y = t->xq[qy];
            double J[2][2];
            q1_point(c, x, y, J);
            double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            double K[2][2] = {{J[1][1] / det, -J[0][1] / det}, {-J[1][0] / det, J[0][0] / det}};
            double w = fabs(det) * t->wq[qx] * t->wq[qy];
            /* u and div u */
            double uv[2];
            p1_eval(u, x, y, uv);
            double divu = 0;
            const double bx[2] = {1 - x, x}, by[2] = {1 - y, y}, d1[2] = {-1, 1};
            for (int ax = 0; ax < 2; ax++)
                for (int ay = 0; ay < 2; ay++) {
                    double gr[2] = {d1[ax] * by[ay], bx[ax] * d1[ay]};
                    for (int a = 0; a < 2; a++)   /* d u_a / d x_a = sum_b K[b][a] dref_b */
                        divu += u[(ax * 2 + ay) * 2 + a] * (K[0][a] * gr[0] + K[1][a] * gr[1]);
                }
            double qv = 0;
            for (int ax = 0; ax < 2; ax++)
                for (int ay = 0; ay < 2; ay++) qv += q[ax * 2 + ay] * dq1d(t, ax, x) * dq1d(t, ay, y);
            for (int ax = 0; ax < 2; ax++)
                for (int ay = 0; ay < 2; ay++) {
                    double ph = dq1d(t, ax, x) * dq1d(t, ay, y);
                    double gr[2] = {ddq1d(t, ax) * dq1d(t, ay, y), dq1d(t, ax, x) * ddq1d(t, ay)};
                    double gp[2] = {K[0][0] * gr[0] + K[1][0] * gr[1], K[0][1] * gr[0] + K[1][1] * gr[1]};
                    A[ax * 2 + ay] += dt * w * qv * (gp[0] * uv[0] + gp[1] * uv[1] + ph * divu);
                }
        }
}

static void facet_point(int f, double s, double *x, double *y, double nref[2], double tref[2])
{
    switch (f) {
    case 0: *x = 0; *y = s; nref[0] = -1; nref[1] = 0; tref[0] = 0; tref[1] = 1; break;
    case 1: *x = 1; *y = s; nref[0] = 1; nref[1] = 0; tref[0] = 0; tref[1] = 1; break;
    case 2: *x = s; *y = 0; nref[0] = 0; nref[1] = -1; tref[0] = 1; tref[1] = 0; break;
    default: *x = s; *y = 1; nref[0] = 0; nref[1] = 1; tref[0] = 1; tref[1] = 0; break;
    }
}

/* outward unit normal and surface measure at a facet point */
static void facet_geometry(const double *c, double x, double y, const double nref[2],
                           const double tref[2], double n[2], double *ds)
{
    double J[2][2];
    q1_point(c, x, y, J);
    double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double K[2][2] = {{J[1][1] / det, -J[0][1] / det}, {-J[1][0] / det, J[0][0] / det}};
    double nn[2] = {K[0][0] * nref[0] + K[1][0] * nref[1], K[0][1] * nref[0] + K[1][1] * nref[1]};
    double len = sqrt(nn[0] * nn[0] + nn[1] * nn[1]);
    n[0] = nn[0] / len;
    n[1] = nn[1] / len;
    double tt[2] = {J[0][0] * tref[0] + J[0][1] * tref[1], J[1][0] * tref[0] + J[1][1] * tref[1]};
    *ds = sqrt(tt[0] * tt[0] + tt[1] * tt[1]);
}

void dg_exterior_facet_kernel(double *A, const double *c, const double *q, const double *u,
                              unsigned facet, const dg_tab *t, double dt, double q_in)
{
    for (int k = 0; k < t->nq; k++) {
        double x, y, nref[2], tref[2], n[2], ds, uv[2];
        facet_point((int)facet, t->xq[k], &x, &y, nref, tref);
        facet_geometry(c, x, y, nref, tref, n, &ds);
        p1_eval(u, x, y, uv);
        double udn = uv[0] * n[0] + uv[1] * n[1];
        double qv = 0;
        for (int ax = 0; ax < 2; ax++)
            for (int ay = 0; ay < 2; ay++) qv += q[ax * 2 + ay] * dq1d(t, ax, x) * dq1d(t, ay, y);
        double flux = (udn < 0 ? udn * q_in : 0.0) + (udn > 0 ? udn * qv : 0.0);
        for (int ax = 0; ax < 2; ax++)
            for (int ay = 0; ay < 2; ay++)
                A[ax * 2 + ay] -= dt * ds * t->wq[k] * dq1d(t, ax, x) * dq1d(t, ay, y) * flux;
    }
}

void dg_interior_facet_kernel(double *A, const double *c, const double *q, const double *u,
                              const unsigned facet[2], const dg_tab *t, double dt)
{
    /* c, q, u: '+' cell then '-' cell; A[0..3] = phi('+') rows, A[4..7] = phi('-') rows */
    for (int k = 0; k < t->nq; k++) {
        double xp, yp, xm, ym, nrp[2], trp[2], nrm[2], trm[2], np_[2], nm[2], dsp, dsm, uv[2];
        facet_point((int)facet[0], t->xq[k], &xp, &yp, nrp, trp);
        facet_point((int)facet[1], t->xq[k], &xm, &ym, nrm, trm);
        facet_geometry(c, xp, yp, nrp, trp, np_, &dsp);
        facet_geometry(c + 8, xm, ym, nrm, trm, nm, &dsm);
        p1_eval(u, xp, yp, uv);                       /* u is continuous: '+' restriction */
        double um[2];
        p1_eval(u + 8, xm, ym, um);
        double udnp = uv[0] * np_[0] + uv[1] * np_[1];
        double udnm = um[0] * nm[0] + um[1] * nm[1];
        double unp = 0.5 * (udnp + fabs(udnp)), unm = 0.5 * (udnm + fabs(udnm));
        double qp = 0, qm = 0;
        for (int ax = 0; ax < 2; ax++)
            for (int ay = 0; ay < 2; ay++) {
                qp += q[ax * 2 + ay] * dq1d(t, ax, xp) * dq1d(t, ay, yp);
                qm += q[4 + ax * 2 + ay] * dq1d(t, ax, xm) * dq1d(t, ay, ym);
            }
        double jump = unp * qp - unm * qm;
        for (int ax = 0; ax < 2; ax++)
            for (int ay = 0; ay < 2; ay++) {
                A[ax * 2 + ay] -= dt * dsp * t->wq[k] * dq1d(t, ax, xp) * dq1d(t, ay, yp) * jump;
                A[4 + ax * 2 + ay] += dt * dsp * t->wq[k] * dq1d(t, ax, xm) * dq1d(t, ay, ym) * jump;
            }
    }
}

/* PyOP2 wrappers (non-extruded): cells, exterior facets, interior facets
 * (reference pyop2/codegen/builder.py:352-429; facet maps
 * firedrake/cython/dmcommon.pyx:1636-1677; local facet number Dat
 * firedrake/assemble.py:1886-1905). */
int orc_dg_cells(int start, int end, double *out, const double *coords, const double *q,
                 const double *u, const int *dgmap, const int *cgmap, int nq, const double *Bend,
                 const double *wq, const double *xq, double dt)
{
    dg_tab t = {nq, 0, 0, Bend, 0, 0, wq, xq};
    for (int n = start; n < end; n++) {
        double A[4] = {0}, c[8], ql[4], ul[8];
        for (int i = 0; i < 4; i++) {
            int v = cgmap[n * 4 + i];
            c[i * 2] = coords[v * 2]; c[i * 2 + 1] = coords[v * 2 + 1];
            ul[i * 2] = u[v * 2]; ul[i * 2 + 1] = u[v * 2 + 1];
            ql[i] = q[dgmap[n * 4 + i]];
        }
        dg_cell_kernel(A, c, ql, ul, &t, dt);
        for (int i = 0; i < 4; i++) out[dgmap[n * 4 + i]] += A[i];
    }
    return 0;
}

int orc_dg_exterior_facets(int start, int end, double *out, const double *coords, const double *q,
                           const double *u, const int *facet_cell, const unsigned *facet_local,
                           const int *dgmap, const int *cgmap, int nq, const double *Bend,
                           const double *wq, const double *xq, double dt, double q_in)
{
    dg_tab t = {nq, 0, 0, Bend, 0, 0, wq, xq};
    for (int f = start; f < end; f++) {
        int n = facet_cell[f];
        double A[4] = {0}, c[8], ql[4], ul[8];
        for (int i = 0; i < 4; i++) {
            int v = cgmap[n * 4 + i];
            c[i * 2] = coords[v * 2]; c[i * 2 + 1] = coords[v * 2 + 1];
            ul[i * 2] = u[v * 2]; ul[i * 2 + 1] = u[v * 2 + 1];
            ql[i] = q[dgmap[n * 4 + i]];
        }
        dg_exterior_facet_kernel(A, c, ql, ul, facet_local[f], &t, dt, q_in);
        for (int i = 0; i < 4; i++) out[dgmap[n * 4 + i]] += A[i];
    }
    return 0;
}

int orc_dg_interior_facets(int start, int end, double *out, const double *coords, const double *q,
                           const double *u, const int *facet_cells, const unsigned *facet_local,
                           const int *dgmap, const int *cgmap, int nq, const double *Bend,
                           const double *wq, const double *xq, double dt)
{
    dg_tab t = {nq, 0, 0, Bend, 0, 0, wq, xq};
    for (int f = start; f < end; f++) {
        double A[8] = {0}, c[16], ql[8], ul[16];
        for (int s = 0; s < 2; s++) {
            int n = facet_cells[f * 2 + s];
            for (int i = 0; i < 4; i++) {
                int v = cgmap[n * 4 + i];
                c[s * 8 + i * 2] = coords[v * 2]; c[s * 8 + i * 2 + 1] = coords[v * 2 + 1];
                ul[s * 8 + i * 2] = u[v * 2]; ul[s * 8 + i * 2 + 1] = u[v * 2 + 1];
                ql[s * 4 + i] = q[dgmap[n * 4 + i]];
            }
        }
        dg_interior_facet_kernel(A, c, ql, ul, facet_local + f * 2, &t, dt);
        for (int s = 0; s < 2; s++) {
            int n = facet_cells[f * 2 + s];
            for (int i = 0; i < 4; i++) out[dgmap[n * 4 + i]] += A[s * 4 + i];
        }
    }
    return 0;
}
