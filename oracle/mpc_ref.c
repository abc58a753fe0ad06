/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
 *
 * mpc_ref.c: CPU restatement of the reference DCM model-predictive controller,
 *   WalkingController  modules/Walking_module/src/WalkingDCMModelPredictiveController.cpp:23-543
 *   MPCSolver          modules/Walking_module/src/MPCSolver.cpp:16-322
 * on top of osqp_ref.c (OSQP restatement).  The iDynTree pieces used by that code
 * (Polygon::XYRectangleFromOffsets, ConvexHullProjectionConstraint::{buildConvexHull,
 * computeMargin}) are third-party and absent from /root/reference; they are restated from
 * their published behaviour: rectangle vertices -> transform -> XY projection -> Andrew
 * monotone-chain hull (counter-clockwise) -> one row per hull edge, A_i = outward normal
 * (dy,-dx) of edge i, b_i = A_i . p_i (un-normalised); margin = signed distance to the
 * nearest hull SEGMENT (negative outside).  PARITY UNPINNED (no reference golden vectors).
 */
#include "oracle_internal.h"
#include "wc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

struct wco_mpc {
    wc_mpc_config cfg;
    int N, n, m_eq;
    double a_dyn, b_dyn;        /* exp(w dT), 1-exp(w dT)  (Controller.cpp:235-236) */
    double Q[4], R[4];
    double foot_front, foot_back, foot_left, foot_right; /* Controller.cpp:297-300 */
    int feet_left, feet_right;  /* m_feetStatus */
    int have_solver;            /* m_currentController != nullptr */
    int n_hull;
    double hull_A[WCO_MAX_HULL * 2], hull_b[WCO_MAX_HULL], hull_v[WCO_MAX_HULL * 2];
    wco_osqp* osqp;             /* NULL until initSolver() (lazy, Controller.cpp:494-501) */
    double *gradient, *lower, *upper; /* MPCSolver::m_gradient / m_lowerBound / m_upperBound */
    double output[2];           /* m_output */
    int linsys_kind;
    int hull_variant;           /* sensitivity studies only (tools/sensitivity_report.py): bit 0 = unit-normal rows,
                                   bits 1.. = cyclic rotation of the row order (iDynTree's order / scaling is unknown) */
    int last_status, last_iter;
};

/* ------------------------------------------------------------------ convex hull */
static double cross2(const double* o, const double* a, const double* b)
{
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0]);
}

int wco_hull_build(const double* pts, int npts, double* A, double* b, double* verts)
{
    /* Andrew monotone chain; pts [npts][2]; returns the number of hull vertices (CCW) */
    double p[2 * WCO_MAX_HULL][2];
    double h[4 * WCO_MAX_HULL][2];
    int i, j, k = 0;
    if (npts > 2 * WCO_MAX_HULL) return -1;
    for (i = 0; i < npts; i++) { p[i][0] = pts[2 * i]; p[i][1] = pts[2 * i + 1]; }
    for (i = 1; i < npts; i++) { /* insertion sort, lexicographic */
        double x = p[i][0], y = p[i][1];
        j = i - 1;
        while (j >= 0 && (p[j][0] > x || (p[j][0] == x && p[j][1] > y))) {
            p[j + 1][0] = p[j][0]; p[j + 1][1] = p[j][1]; j--;
        }
        p[j + 1][0] = x; p[j + 1][1] = y;
    }
    for (i = 0; i < npts; i++) {
        while (k >= 2 && cross2(h[k - 2], h[k - 1], p[i]) <= 0) k--;
        h[k][0] = p[i][0]; h[k][1] = p[i][1]; k++;
    }
    for (i = npts - 2, j = k + 1; i >= 0; i--) {
        while (k >= j && cross2(h[k - 2], h[k - 1], p[i]) <= 0) k--;
        h[k][0] = p[i][0]; h[k][1] = p[i][1]; k++;
    }
    k--; /* last point equals the first */
    if (k > WCO_MAX_HULL) return -1;
    for (i = 0; i < k; i++) {
        const double* p0 = h[i];
        const double* p1 = h[(i + 1) % k];
        A[2 * i] = p1[1] - p0[1];
        A[2 * i + 1] = -(p1[0] - p0[0]);
        b[i] = A[2 * i] * p0[0] + A[2 * i + 1] * p0[1];
        verts[2 * i] = p0[0]; verts[2 * i + 1] = p0[1];
    }
    return k;
}

double wco_hull_margin(const double* A, const double* b, const double* verts, int nh, const double* z)
{
    int inside = 1, i;
    double best = INFINITY;
    for (i = 0; i < nh; i++)
        if (!(A[2 * i] * z[0] + A[2 * i + 1] * z[1] <= b[i])) { inside = 0; break; }
    for (i = 0; i < nh; i++) {
        const double* p0 = verts + 2 * i;
        const double* p1 = verts + 2 * ((i + 1) % nh);
        double dx = p1[0] - p0[0], dy = p1[1] - p0[1];
        double len2 = dx * dx + dy * dy;
        double t = len2 > 0 ? ((z[0] - p0[0]) * dx + (z[1] - p0[1]) * dy) / len2 : 0.0;
        t = t < 0 ? 0 : (t > 1 ? 1 : t);
        double ex = z[0] - (p0[0] + t * dx), ey = z[1] - (p0[1] + t * dy);
        double d = sqrt(ex * ex + ey * ey);
        if (d < best) best = d;
    }
    return inside ? best : -best;
}

/* foot rectangle vertices (iDynTree Polygon::XYRectangleFromOffsets(front, back, left, right)),
 * transformed by world_H_foot (row-major 3x4) and projected on the world XY plane */
static void foot_points(const wco_mpc* c, const double* H, double* out /* [4][2] */)
{
    const double vx[4] = {c->foot_front, -c->foot_back, -c->foot_back, c->foot_front};
    const double vy[4] = {c->foot_left, c->foot_left, -c->foot_right, -c->foot_right};
    for (int k = 0; k < 4; k++) {
        out[2 * k] = H[0] * vx[k] + H[1] * vy[k] + H[3];
        out[2 * k + 1] = H[4] * vx[k] + H[5] * vy[k] + H[7];
    }
}

/* ------------------------------------------------------------------ QP data (Controller.cpp:23-168, MPCSolver.cpp:76-123) */
static wco_csc* build_P(const wco_mpc* c)
{
    const int N = c->N;
    int cap = 3 * (N + 1) + 7 * N, nt = 0, k;
    int* ti = (int*)malloc(sizeof(int) * (size_t)cap);
    int* tj = (int*)malloc(sizeof(int) * (size_t)cap);
    double* tx = (double*)malloc(sizeof(double) * (size_t)cap);
#define PUSH(i, j, v) do { if ((v) != 0.0) { ti[nt] = (i); tj[nt] = (j); tx[nt] = (v); nt++; } } while (0)
    for (k = 0; k <= N; k++) { /* Q~ = diag(Q,...,Q), N+1 blocks; upper triangle */
        PUSH(2 * k, 2 * k, c->Q[0]);
        PUSH(2 * k, 2 * k + 1, c->Q[1]);
        PUSH(2 * k + 1, 2 * k + 1, c->Q[3]);
    }
    const int off = 2 * (N + 1);
    for (k = 0; k < N; k++) { /* Theta' R~ Theta: diag 2R (last R), off-diagonal -R */
        const double s = (k < N - 1) ? 2.0 : 1.0;
        PUSH(off + 2 * k, off + 2 * k, s * c->R[0]);
        PUSH(off + 2 * k, off + 2 * k + 1, s * c->R[1]);
        PUSH(off + 2 * k + 1, off + 2 * k + 1, s * c->R[3]);
        if (k < N - 1) {
            PUSH(off + 2 * k, off + 2 * k + 2, -c->R[0]);
            PUSH(off + 2 * k, off + 2 * k + 3, -c->R[1]);
            PUSH(off + 2 * k + 1, off + 2 * k + 2, -c->R[2]);
            PUSH(off + 2 * k + 1, off + 2 * k + 3, -c->R[3]);
        }
    }
    wco_csc* P = wco_csc_from_triplets(c->n, c->n, nt, ti, tj, tx);
    free(ti); free(tj); free(tx);
    return P;
}

static wco_csc* build_A(const wco_mpc* c)
{
    const int N = c->N, nh = c->n_hull;
    int cap = 2 * (N + 1) + 4 * N + 2 * nh, nt = 0, k, a, h;
    int* ti = (int*)malloc(sizeof(int) * (size_t)cap);
    int* tj = (int*)malloc(sizeof(int) * (size_t)cap);
    double* tx = (double*)malloc(sizeof(double) * (size_t)cap);
    const int off = 2 * (N + 1);
    for (k = 0; k < 2 * (N + 1); k++) PUSH(k, k, -1.0);
    for (k = 0; k < N; k++)
        for (a = 0; a < 2; a++) {
            PUSH(2 * (k + 1) + a, 2 * k + a, c->a_dyn);
            PUSH(2 * (k + 1) + a, off + 2 * k + a, c->b_dyn);
        }
    for (h = 0; h < nh; h++) { /* hull rows act on u_0 only (MPCSolver.cpp:83-87) */
        PUSH(off + h, off + 0, c->hull_A[2 * h]);
        PUSH(off + h, off + 1, c->hull_A[2 * h + 1]);
    }
#undef PUSH
    wco_csc* A = wco_csc_from_triplets(c->m_eq + nh, c->n, nt, ti, tj, tx);
    free(ti); free(tj); free(tx);
    return A;
}

/* stage-interleaved KKT ordering (any symmetric permutation of a quasi-definite matrix has
 * an LDL^T; OSQP uses AMD — the choice only changes rounding) */
static int* build_perm(const wco_mpc* c)
{
    const int N = c->N, n = c->n, nh = c->n_hull;
    int* perm = (int*)malloc(sizeof(int) * (size_t)(n + c->m_eq + nh));
    int k, h, w = 0;
    const int off = 2 * (N + 1);
    for (k = 0; k <= N; k++) {
        perm[w++] = n + 2 * k; perm[w++] = n + 2 * k + 1; /* nu_k */
        perm[w++] = 2 * k; perm[w++] = 2 * k + 1;         /* x_k */
        if (k == 0) for (h = 0; h < nh; h++) perm[w++] = n + off + h;
        if (k < N) { perm[w++] = off + 2 * k; perm[w++] = off + 2 * k + 1; }
    }
    return perm;
}

/* ------------------------------------------------------------------ controller */
wco_mpc* wco_mpc_create(const wc_mpc_config* cfg)
{
    if (!cfg || cfg->horizon < 1) return NULL;
    wco_mpc* c = (wco_mpc*)calloc(1, sizeof(wco_mpc));
    c->cfg = *cfg;
    c->N = cfg->horizon;
    c->n = 4 * c->N + 2;
    c->m_eq = 2 * (c->N + 1);
    const double omega = sqrt(cfg->gravity / cfg->com_height);
    c->a_dyn = exp(omega * cfg->sampling_time);
    c->b_dyn = 1.0 - exp(omega * cfg->sampling_time);
    memcpy(c->Q, cfg->Q, sizeof(c->Q));
    memcpy(c->R, cfg->R, sizeof(c->R));
    const double x1 = cfg->foot_x[0], x2 = cfg->foot_x[1], y1 = cfg->foot_y[0], y2 = cfg->foot_y[1];
    c->foot_front = fabs(x1 > x2 ? x1 : x2);
    c->foot_back = fabs(x1 < x2 ? x1 : x2);
    c->foot_left = fabs(y1 > y2 ? y1 : y2);
    c->foot_right = fabs(y1 < y2 ? y1 : y2);
    c->output[0] = cfg->initial_zmp[0];
    c->output[1] = cfg->initial_zmp[1];
    c->gradient = (double*)calloc((size_t)c->n, sizeof(double));
    c->lower = (double*)calloc((size_t)(c->m_eq + WCO_MAX_HULL), sizeof(double));
    c->upper = (double*)calloc((size_t)(c->m_eq + WCO_MAX_HULL), sizeof(double));
    wco_mpc_reset(c);
    return c;
}

void wco_mpc_free(wco_mpc* c)
{
    if (!c) return;
    wco_osqp_free(c->osqp);
    free(c->gradient); free(c->lower); free(c->upper);
    free(c);
}

void wco_mpc_reset(wco_mpc* c) { c->feet_left = 0; c->feet_right = 0; }

void wco_mpc_set_weights(wco_mpc* c, const double q_diag[2], const double r_diag[2])
{
    c->Q[0] = q_diag[0]; c->Q[1] = 0; c->Q[2] = 0; c->Q[3] = q_diag[1];
    c->R[0] = r_diag[0]; c->R[1] = 0; c->R[2] = 0; c->R[3] = r_diag[1];
}

void wco_mpc_set_linsys(wco_mpc* c, int kind) { c->linsys_kind = kind; }
void wco_mpc_set_hull_variant(wco_mpc* c, int variant) { c->hull_variant = variant; }

int wco_mpc_tick(wco_mpc* c, const double* lf_H, const double* rf_H, int l_contact, int r_contact,
                 const double* dcm, const double* ref, int reset_traj,
                 double* zmp, int* osqp_status, int* iters)
{
    const int N = c->N, n = c->n, m_eq = c->m_eq;
    int flags = 0, i, k;
    l_contact = l_contact ? 1 : 0; r_contact = r_contact ? 1 : 0;
    c->last_status = WC_OSQP_UNSOLVED; c->last_iter = 0;
    zmp[0] = c->output[0]; zmp[1] = c->output[1];
    if (osqp_status) *osqp_status = WC_OSQP_UNSOLVED;
    if (iters) *iters = 0;

    /* ---- setConvexHullConstraint (Controller.cpp:364-435) */
    /* "None foot is in contact" (:404-408) returns false and the module stops; it has no defined behaviour for a second
     * tick without contact (m_feetStatus would compare equal and the stale hull be solved).  Restated as: flagged on
     * EVERY tick without contact, nothing solved (wc_capi.h WC_MPC_FLAG_NO_CONTACT). */
    if (!l_contact && !r_contact) {
        c->feet_left = 0; c->feet_right = 0;
        return WC_MPC_FLAG_NO_CONTACT;
    }
    if (!(c->feet_left == l_contact && c->feet_right == r_contact)) {
        c->feet_left = l_contact; c->feet_right = r_contact;
        double pts[16];
        int npts = 0;
        if (l_contact && r_contact) { foot_points(c, lf_H, pts); foot_points(c, rf_H, pts + 8); npts = 8; }
        else if (l_contact) { foot_points(c, lf_H, pts); npts = 4; }
        else { foot_points(c, rf_H, pts); npts = 4; }
        c->n_hull = wco_hull_build(pts, npts, c->hull_A, c->hull_b, c->hull_v);
        if (c->hull_variant) { /* same polygon, rows re-ordered / re-scaled: what another hull builder might hand to OSQP */
            const int nh = c->n_hull, rot = (c->hull_variant >> 1) % (nh > 0 ? nh : 1);
            double A2[WCO_MAX_HULL * 2], b2[WCO_MAX_HULL], v2[WCO_MAX_HULL * 2];
            for (i = 0; i < nh; i++) {
                const int s = (i + rot) % nh;
                double sc = 1.0;
                if (c->hull_variant & 1) sc = 1.0 / sqrt(c->hull_A[2 * s] * c->hull_A[2 * s] + c->hull_A[2 * s + 1] * c->hull_A[2 * s + 1]);
                A2[2 * i] = sc * c->hull_A[2 * s]; A2[2 * i + 1] = sc * c->hull_A[2 * s + 1]; b2[i] = sc * c->hull_b[s];
                v2[2 * i] = c->hull_v[2 * s]; v2[2 * i + 1] = c->hull_v[2 * s + 1];
            }
            memcpy(c->hull_A, A2, sizeof(double) * 2 * (size_t)nh); memcpy(c->hull_b, b2, sizeof(double) * (size_t)nh);
            memcpy(c->hull_v, v2, sizeof(double) * 2 * (size_t)nh);
        }
        /* new MPCSolver (MPCSolver.cpp:16-52): zero gradient/bounds, hull lower bounds = -INFTY */
        wco_osqp_free(c->osqp);
        c->osqp = NULL;
        c->have_solver = 1;
        memset(c->gradient, 0, sizeof(double) * (size_t)n);
        for (i = 0; i < m_eq + WCO_MAX_HULL; i++) { c->lower[i] = 0.0; c->upper[i] = 0.0; }
        for (i = m_eq; i < m_eq + c->n_hull; i++) c->lower[i] = -c->cfg.osqp.infty;
        flags |= WC_MPC_FLAG_NEW_SOLVER;
    }
    if (!c->have_solver) return flags | WC_MPC_FLAG_NO_SOLVER;
    const int m = m_eq + c->n_hull;

    /* ---- setFeedback -> MPCSolver::setBounds (MPCSolver.cpp:125-181) */
    c->lower[0] = -dcm[0]; c->lower[1] = -dcm[1];
    c->upper[0] = -dcm[0]; c->upper[1] = -dcm[1];
    for (i = 0; i < c->n_hull; i++) c->upper[m_eq + i] = c->hull_b[i];
    if (c->osqp) wco_osqp_update_bounds(c->osqp, c->lower, c->upper);

    /* ---- setReferenceSignal -> MPCSolver::setGradient (MPCSolver.cpp:183-266) */
    if (!c->osqp || reset_traj) {
        for (k = 0; k <= N; k++) {
            c->gradient[2 * k] = -(c->Q[0] * ref[2 * k] + c->Q[1] * ref[2 * k + 1]);
            c->gradient[2 * k + 1] = -(c->Q[2] * ref[2 * k] + c->Q[3] * ref[2 * k + 1]);
        }
    } else {
        for (k = 0; k < N; k++) {
            c->gradient[2 * k] = c->gradient[2 * (k + 1)];
            c->gradient[2 * k + 1] = c->gradient[2 * (k + 1) + 1];
        }
        c->gradient[2 * N] = -(c->Q[0] * ref[2 * N] + c->Q[1] * ref[2 * N + 1]);
        c->gradient[2 * N + 1] = -(c->Q[2] * ref[2 * N] + c->Q[3] * ref[2 * N + 1]);
    }
    {
        /* m_gradientSubmatrix * previousControllerOutput, G = -Theta' R~ e1 = [-R; 0; ...] */
        const int off = 2 * (N + 1);
        for (i = 0; i < 2 * N; i++) c->gradient[off + i] = 0.0;
        c->gradient[off] = -(c->R[0] * c->output[0] + c->R[1] * c->output[1]);
        c->gradient[off + 1] = -(c->R[2] * c->output[0] + c->R[3] * c->output[1]);
    }
    if (c->osqp) wco_osqp_update_lin_cost(c->osqp, c->gradient);

    /* ---- solve (Controller.cpp:491-521) */
    if (!c->osqp) {
        wco_csc* P = build_P(c);
        wco_csc* A = build_A(c);
        int* perm = build_perm(c);
        c->osqp = wco_osqp_setup(P, A, c->gradient, c->lower, c->upper, &c->cfg.osqp, perm);
        wco_csc_free(P); wco_csc_free(A); free(perm);
        if (!c->osqp) return flags | WC_MPC_FLAG_NOT_SOLVED;
        if (c->linsys_kind) wco_osqp_set_linsys(c->osqp, c->linsys_kind);
    }
    (void)m;
    const int st = wco_osqp_solve(c->osqp);
    c->last_status = st; c->last_iter = wco_osqp_iter(c->osqp);
    if (osqp_status) *osqp_status = st;
    if (iters) *iters = c->last_iter;
    /* OsqpEigen::Solver::solve() (MPCSolver.cpp:309): osqp-eigen of the reference's time returns osqp_solve() == 0, i.e. true
     * for every status but a non-convex problem; since 0.4 it is false unless the status is OSQP_SOLVED
     * (wc_osqp_settings::osqp_eigen_checks_status).  The controller then reads OSQP's stored solution (Controller.cpp:509-511):
     * the last iterate, or NaN for the infeasible statuses (OSQP store_solution). */
    {
        const int ok = c->cfg.osqp.osqp_eigen_checks_status ? (st == WC_OSQP_SOLVED)
                                                            : (st != WC_OSQP_NON_CVX && st != WC_OSQP_UNSOLVED);
        if (!ok) return flags | WC_MPC_FLAG_NOT_SOLVED;
    }
    const int has_sol = !(st == WC_OSQP_PRIMAL_INFEASIBLE || st == WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                          st == WC_OSQP_DUAL_INFEASIBLE || st == WC_OSQP_DUAL_INFEASIBLE_INACCURATE);
    const double* sol = wco_osqp_solution_x(c->osqp);
    c->output[0] = has_sol ? sol[2 * (N + 1)] : NAN;
    c->output[1] = has_sol ? sol[2 * (N + 1) + 1] : NAN;
    zmp[0] = c->output[0]; zmp[1] = c->output[1];
    if (wco_hull_margin(c->hull_A, c->hull_b, c->hull_v, c->n_hull, c->output) < -c->cfg.hull_tolerance)
        flags |= WC_MPC_FLAG_OUTSIDE_HULL;
    return flags;
}

int wco_mpc_n_hull(const wco_mpc* c) { return c->n_hull; }
void wco_mpc_get_hull(const wco_mpc* c, double* A, double* b)
{
    memcpy(A, c->hull_A, sizeof(double) * 2 * (size_t)c->n_hull);
    memcpy(b, c->hull_b, sizeof(double) * (size_t)c->n_hull);
}
int wco_mpc_get_solution(const wco_mpc* c, double* primal)
{
    if (!c->osqp) return -1;
    memcpy(primal, wco_osqp_solution_x(c->osqp), sizeof(double) * (size_t)c->n);
    return 0;
}
double wco_mpc_rho(const wco_mpc* c) { return c->osqp ? wco_osqp_rho(c->osqp) : 0.0; }
int wco_mpc_rho_updates(const wco_mpc* c) { return c->osqp ? wco_osqp_rho_updates(c->osqp) : 0; }

/* Exposes the assembled QP of the CURRENT solver phase (dense, row-major) for independent
 * cross-checks from Python: P (n x n, full symmetric), A (m x n), q, l, u.  Returns m. */
int wco_mpc_get_qp_dense(const wco_mpc* c, double* Pd, double* Ad, double* q, double* l, double* u)
{
    const int n = c->n, m = c->m_eq + c->n_hull;
    int j, p;
    wco_csc* P = build_P(c);
    wco_csc* A = build_A(c);
    memset(Pd, 0, sizeof(double) * (size_t)n * (size_t)n);
    memset(Ad, 0, sizeof(double) * (size_t)m * (size_t)n);
    for (j = 0; j < n; j++)
        for (p = P->p[j]; p < P->p[j + 1]; p++) {
            Pd[(size_t)P->i[p] * n + j] = P->x[p];
            Pd[(size_t)j * n + P->i[p]] = P->x[p];
        }
    for (j = 0; j < n; j++)
        for (p = A->p[j]; p < A->p[j + 1]; p++) Ad[(size_t)A->i[p] * n + j] = A->x[p];
    memcpy(q, c->gradient, sizeof(double) * (size_t)n);
    memcpy(l, c->lower, sizeof(double) * (size_t)m);
    memcpy(u, c->upper, sizeof(double) * (size_t)m);
    wco_csc_free(P); wco_csc_free(A);
    return m;
}

/* ------------------------------------------------------------------ batched driver (same SoA layout as the C ABI) */
struct wco_mpc_batch {
    int B;
    wco_mpc** inst;
};

wco_mpc_batch* wco_mpc_batch_create(const wc_mpc_config* cfg, int B)
{
    wco_mpc_batch* b = (wco_mpc_batch*)calloc(1, sizeof(wco_mpc_batch));
    b->B = B;
    b->inst = (wco_mpc**)calloc((size_t)B, sizeof(wco_mpc*));
    for (int i = 0; i < B; i++) b->inst[i] = wco_mpc_create(cfg);
    return b;
}
void wco_mpc_batch_free(wco_mpc_batch* b)
{
    if (!b) return;
    for (int i = 0; i < b->B; i++) wco_mpc_free(b->inst[i]);
    free(b->inst); free(b);
}
wco_mpc* wco_mpc_batch_get(wco_mpc_batch* b, int i) { return b->inst[i]; }
void wco_mpc_batch_set_weights(wco_mpc_batch* b, const double* q_diag, const double* r_diag)
{
    for (int i = 0; i < b->B; i++) {
        double q[2] = {q_diag[i], q_diag[b->B + i]}, r[2] = {r_diag[i], r_diag[b->B + i]};
        wco_mpc_set_weights(b->inst[i], q, r);
    }
}
void wco_mpc_batch_reset(wco_mpc_batch* b, const unsigned char* mask)
{
    for (int i = 0; i < b->B; i++) if (!mask || mask[i]) wco_mpc_reset(b->inst[i]);
}

typedef struct {
    wco_mpc_batch* b;
    const double *lf_H, *rf_H, *dcm, *ref;
    const unsigned char *l_contact, *r_contact, *reset;
    double* zmp;
    int *osqp_status, *iters, *flags;
    int begin, end;
} mpc_job;

static void* mpc_worker(void* arg)
{
    mpc_job* j = (mpc_job*)arg;
    const int B = j->b->B;
    for (int i = j->begin; i < j->end; i++) {
        wco_mpc* c = j->b->inst[i];
        const int N = c->N;
        double lf[12], rf[12], d[2], z[2];
        double* r = (double*)malloc(sizeof(double) * 2 * (size_t)(N + 1));
        int st = 0, it = 0, e;
        for (e = 0; e < 12; e++) { lf[e] = j->lf_H[(size_t)e * B + i]; rf[e] = j->rf_H[(size_t)e * B + i]; }
        d[0] = j->dcm[i]; d[1] = j->dcm[(size_t)B + i];
        for (e = 0; e < 2 * (N + 1); e++) r[e] = j->ref[(size_t)e * B + i];
        int f = wco_mpc_tick(c, lf, rf, j->l_contact[i], j->r_contact[i], d, r, j->reset ? j->reset[i] : 0, z, &st, &it);
        j->zmp[i] = z[0]; j->zmp[(size_t)B + i] = z[1];
        if (j->osqp_status) j->osqp_status[i] = st;
        if (j->iters) j->iters[i] = it;
        if (j->flags) j->flags[i] = f;
        free(r);
    }
    return NULL;
}

/* one solver instance per thread, static partition of the instances over n_threads */
void wco_mpc_batch_tick(wco_mpc_batch* b, const double* lf_H, const double* rf_H,
                        const unsigned char* l_contact, const unsigned char* r_contact,
                        const double* dcm, const double* ref, const unsigned char* reset,
                        double* zmp, int* osqp_status, int* iters, int* flags, int n_threads)
{
    const int B = b->B;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > B) n_threads = B > 0 ? B : 1;
    mpc_job* jobs = (mpc_job*)calloc((size_t)n_threads, sizeof(mpc_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        mpc_job j = {b, lf_H, rf_H, dcm, ref, l_contact, r_contact, reset, zmp, osqp_status, iters, flags,
                     (int)((long long)B * t / n_threads), (int)((long long)B * (t + 1) / n_threads)};
        jobs[t] = j;
        if (t > 0) pthread_create(&th[t], NULL, mpc_worker, &jobs[t]);
    }
    mpc_worker(&jobs[0]);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    free(jobs); free(th);
}

void wco_mpc_batch_get_solution(wco_mpc_batch* b, double* primal, double* rho, int* rho_updates)
{
    const int B = b->B;
    for (int i = 0; i < B; i++) {
        wco_mpc* c = b->inst[i];
        if (primal) {
            if (c->osqp) {
                const double* s = wco_osqp_solution_x(c->osqp);
                for (int k = 0; k < c->n; k++) primal[(size_t)k * B + i] = s[k];
            } else
                for (int k = 0; k < c->n; k++) primal[(size_t)k * B + i] = NAN;
        }
        if (rho) rho[i] = wco_mpc_rho(c);
        if (rho_updates) rho_updates[i] = wco_mpc_rho_updates(c);
    }
}
