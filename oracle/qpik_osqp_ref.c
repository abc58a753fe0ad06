/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 * CPU restatement of WalkingQPIK_osqp:
 *   modules/Walking_module/src/WalkingQPInverseKinematics_osqp.cpp:135-408 over osqp-eigen / OSQP (external, not vendored).
 * The QP is the one of the qpOASES variant with the three divergences of SURVEY quirk 10:
 *   (1) the neck term of the gradient carries an extra k_attFoot factor (:183,193);
 *   (2) a foot whose desired twist has twist(0) == twist(1) == 0 is constrained to the bare twist (:286-306);
 *   (3) the nJ "velocity limit" rows are EMPTY rows of A (m_jointRegularizationLinearConstraintTriplets is never filled,
 *       …_osqp.hpp:18, …_osqp.cpp:227-235) with bounds [minVel, maxVel] (:23-51): the joint velocities are NOT limited.
 * Solver life cycle restated from osqp-eigen's Solver (updateHessianMatrix / updateGradient / updateLinearConstraintsMatrix /
 * updateBounds / solve) over OSQP 0.6 (osqp_update_P, osqp_update_lin_cost, osqp_update_A, osqp_update_bounds,
 * osqp_warm_start; default settings except verbose, linear system solver 0 = QDLDL):
 *   first solve()    : osqp_setup (Ruiz scaling, rho_vec, KKT factorisation), cold start;
 *   same P pattern   : osqp_update_P, osqp_update_lin_cost, osqp_update_A, osqp_update_bounds in the order of solve() (:308-334);
 *                      each matrix update unscales, overwrites and RE-EQUILIBRATES from scratch, the scaled iterates and rho
 *                      carry over untouched (warm start);
 *   P pattern changed: m_hessianEigenDense.sparseView() drops exact zeros (:150), so the pattern follows the kinematic
 *                      support of the neck Jacobian and changes when the stance foot changes; osqp-eigen then reads the
 *                      workspace iterates, clears the solver, runs osqp_setup again (rho back to its setting) and warm-starts
 *                      it with those iterates (osqp_warm_start scales them with the new D, E, c and sets z = A x).
 *                      DEVIATION: osqp-eigen does this inside updateHessianMatrix, i.e. with the PREVIOUS tick's A and q still
 *                      in place, and only then updates q, A, bounds; here the new solver is set up directly on this tick's
 *                      data (the intermediate equilibration with stale A / q is not reproduced).
 * A keeps every entry of the task-row block, zeros included (iDynTree::Triplets::addSubMatrix pushes all elements, :206-218),
 * so its pattern never changes.  solve() returns what OsqpEigen::Solver::solve() returns -- by default osqp_solve's exit flag
 * (false only for a non-convex problem: osqp-eigen of the reference's time), with osqp_eigen_checks_status also false for
 * every status but OSQP_SOLVED (osqp-eigen >= 0.4); the same switch and the same rule as the MPC -- and then applies
 * isSolutionFeasible (:395-408, tolerance 1, `||`).
 * PARITY UNPINNED: neither the reference nor this container holds osqp-eigen / OSQP or any golden vector for this class.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_internal.h"
#include "wc_oracle.h"

#define IKO_MAXN 40
#define IKO_MAXM 64

struct wco_ikosqp {
    wc_ik_config cfg;
    wc_osqp_settings st;
    wco_osqp* w;
    int nV, nC, m;
    uint8_t mask[IKO_MAXN * IKO_MAXN]; /* upper-triangular nonzero pattern of H of the live solver */
    int n_setups;
};

wco_ikosqp* wco_ikosqp_create(const wc_ik_config* cfg, const wc_osqp_settings* st)
{
    wco_ikosqp* h = (wco_ikosqp*)calloc(1, sizeof(*h));
    h->cfg = *cfg;
    if (st) h->st = *st; else wc_osqp_default_settings(&h->st);
    h->nV = cfg->n_dofs + 6;
    h->nC = cfg->use_com_as_constraint ? 15 : 12;
    h->m = h->nC + cfg->n_dofs;
    return h;
}

void wco_ikosqp_free(wco_ikosqp* h)
{
    if (!h) return;
    wco_osqp_free(h->w);
    free(h);
}

int wco_ikosqp_setups(const wco_ikosqp* h) { return h->n_setups; }

/* upper triangle of H in CSC order (column j, rows i <= j with H[i][j] != 0) */
static int pack_P(int nV, const double* H, const uint8_t* mask, double* Px)
{
    int k = 0;
    for (int j = 0; j < nV; j++)
        for (int i = 0; i <= j; i++)
            if (mask[i * nV + j]) Px[k++] = H[i * nV + j];
    return k;
}

int wco_ikosqp_solve(wco_ikosqp* h,
                     const double* q_meas, const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                     const double* J_lf, const double* J_rf, const double* J_neck6, const double* J_com,
                     const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                     const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                     double* dq, double* lf_err, double* rf_err, int* osqp_status, int* iters, int* flags)
{
    const int nV = h->nV, nC = h->nC, m = h->m, nJ = h->cfg.n_dofs;
    double H[IKO_MAXN * IKO_MAXN], g[IKO_MAXN], A[15 * IKO_MAXN], b[15], l[IKO_MAXM], u[IKO_MAXM];
    wco_ik_build_qp(&h->cfg, 1, q_meas, lf_H, rf_H, neck_R, com_meas, J_lf, J_rf, J_neck6, J_com, lf_H_des, rf_H_des, lf_twist,
                    rf_twist, com_pos_des, com_vel_des, neck_R_des, H, g, A, b);
    for (int r = 0; r < nC; r++) { l[r] = b[r]; u[r] = b[r]; }                                         /* setBounds */
    for (int j = 0; j < nJ; j++) { l[nC + j] = h->cfg.min_vel[j]; u[nC + j] = h->cfg.max_vel[j]; }     /* setVelocityBounds */
    uint8_t mask[IKO_MAXN * IKO_MAXN];
    int same = h->w != NULL;
    for (int i = 0; i < nV; i++)
        for (int j = i; j < nV; j++) {
            mask[i * nV + j] = H[i * nV + j] != 0.0; /* sparseView(): exact zeros are dropped */
            if (h->w && mask[i * nV + j] != h->mask[i * nV + j]) same = 0;
        }
    double Px[IKO_MAXN * IKO_MAXN], Ax[15 * IKO_MAXN];
    /* A in CSC order: column j, task rows 0..nC-1 (all stored), no entries in the velocity rows */
    for (int j = 0; j < nV; j++) for (int r = 0; r < nC; r++) Ax[j * nC + r] = A[r * nV + j];
    if (same) {
        pack_P(nV, H, mask, Px);
        wco_osqp_update_P(h->w, Px);          /* setHessianMatrix          (:135-167) */
        wco_osqp_update_lin_cost(h->w, g);    /* setGradientVector         (:169-216) */
        wco_osqp_update_A(h->w, Ax);          /* setLinearConstraintMatrix (:218-262) */
        wco_osqp_update_bounds(h->w, l, u);   /* setBounds                 (:264-347) */
    } else {
        double xs[IKO_MAXN], ys[IKO_MAXM];
        const int had = h->w != NULL;
        if (had) wco_osqp_get_iterate(h->w, xs, ys);
        wco_osqp_free(h->w); h->w = NULL;
        int ti[IKO_MAXN * IKO_MAXN], tj[IKO_MAXN * IKO_MAXN]; double tx[IKO_MAXN * IKO_MAXN];
        int nt = 0;
        for (int j = 0; j < nV; j++)
            for (int i = 0; i <= j; i++)
                if (mask[i * nV + j]) { ti[nt] = i; tj[nt] = j; tx[nt] = H[i * nV + j]; nt++; }
        wco_csc* P = wco_csc_from_triplets(nV, nV, nt, ti, tj, tx);
        int ai[15 * IKO_MAXN], aj[15 * IKO_MAXN]; double ax[15 * IKO_MAXN];
        nt = 0;
        for (int j = 0; j < nV; j++) for (int r = 0; r < nC; r++) { ai[nt] = r; aj[nt] = j; ax[nt] = A[r * nV + j]; nt++; }
        wco_csc* Ac = wco_csc_from_triplets(m, nV, nt, ai, aj, ax);
        h->w = wco_osqp_setup(P, Ac, g, l, u, &h->st, NULL);
        wco_csc_free(P); wco_csc_free(Ac);
        memcpy(h->mask, mask, sizeof(mask));
        h->n_setups++;
        if (!h->w) { if (osqp_status) *osqp_status = WC_OSQP_NON_CVX; if (flags) *flags = 1; return 1; }
        if (had) wco_osqp_warm_start(h->w, xs, ys);
    }
    const int st = wco_osqp_solve(h->w);
    if (osqp_status) *osqp_status = st;
    if (iters) *iters = wco_osqp_iter(h->w);
    const double* x = wco_osqp_solution_x(h->w);
    int fl = 0;
    /* m_optimizerSolver->solve() (:377): false for a non-convex problem (osqp_solve's exit flag), and with osqp-eigen >= 0.4
     * (osqp_eigen_checks_status) for every status but OSQP_SOLVED -- the same rule as the MPC (mpc_ref.c) */
    if (st == WC_OSQP_NON_CVX || st == WC_OSQP_UNSOLVED || (h->st.osqp_eigen_checks_status && st != WC_OSQP_SOLVED)) fl |= 1;
    /* isSolutionFeasible (:395-408) */
    {
        double mu = -INFINITY, ml = -INFINITY;
        for (int r = 0; r < m; r++) {
            double ax_r = 0.0;
            if (r < nC) for (int c = 0; c < nV; c++) ax_r += A[r * nV + c] * x[c];
            if (ax_r - u[r] > mu) mu = ax_r - u[r];
            if (ax_r - l[r] > ml) ml = ax_r - l[r];
        }
        if (!((mu < 1.0) || (ml > -1.0))) fl |= 2;
    }
    if (flags) *flags = fl;
    for (int j = 0; j < nJ; j++) dq[j] = x[6 + j];                                                     /* getSolution (:410-428) */
    for (int r = 0; r < 6; r++) {                                                                       /* get*FootError (:430-454) */
        double sl = l[r], sr = l[6 + r];
        for (int c = 0; c < nV; c++) { sl -= J_lf[r * nV + c] * x[c]; sr -= J_rf[r * nV + c] * x[c]; }
        if (lf_err) lf_err[r] = sl;
        if (rf_err) rf_err[r] = sr;
    }
    return fl;
}

/* ---- batched FK + osqp QP-IK tick, SoA layout identical to wc_ikosqp_tick_dev; persistent per-instance solvers ---- */
struct wco_ikosqp_batch { int B; wc_model model; wco_ikosqp** inst; };

wco_ikosqp_batch* wco_ikosqp_batch_create(const wc_ik_config* cfg, const wc_osqp_settings* st, const wc_model* model, int B)
{
    wco_ikosqp_batch* b = (wco_ikosqp_batch*)calloc(1, sizeof(*b));
    b->B = B; b->model = *model;
    b->inst = (wco_ikosqp**)calloc((size_t)B, sizeof(*b->inst));
    for (int i = 0; i < B; i++) b->inst[i] = wco_ikosqp_create(cfg, st);
    return b;
}

void wco_ikosqp_batch_free(wco_ikosqp_batch* b)
{
    if (!b) return;
    for (int i = 0; i < b->B; i++) wco_ikosqp_free(b->inst[i]);
    free(b->inst); free(b);
}

static void gather_col(const double* src, int rows, int B, int i, double* dst)
{
    for (int r = 0; r < rows; r++) dst[r] = src[(size_t)r * B + i];
}

void wco_ikosqp_batch_tick(wco_ikosqp_batch* b, const double* q_meas, const double* q_lin, const double* base_H,
                           const unsigned char* left_is_fixed, const double* lf_H_des, const double* rf_H_des,
                           const double* lf_twist, const double* rf_twist, const double* com_pos_des, const double* com_vel_des,
                           const double* com_meas, const double* neck_R_des,
                           double* dq, double* lf_err, double* rf_err, int* status, int* iters, int* flags, int* setups)
{
    const int B = b->B;
    for (int i = 0; i < B; i++) {
        wco_ikosqp* h = b->inst[i];
        const int nJ = h->cfg.n_dofs;
        double qm[WC_MAX_DOFS], ql[WC_MAX_DOFS], bH[12], lfd[12], rfd[12], ltw[6], rtw[6], cpd[3], cvd[3], cm[3], nRd[9];
        double lfH[12], rfH[12], nR[9], com[3];
        double Jl[6 * IKO_MAXN], Jr[6 * IKO_MAXN], Jn[6 * IKO_MAXN], Jc[3 * IKO_MAXN];
        double d[WC_MAX_DOFS], le[6], re[6];
        int st = 0, it = 0, fl = 0;
        gather_col(q_meas, nJ, B, i, qm); gather_col(q_lin, nJ, B, i, ql); gather_col(base_H, 12, B, i, bH);
        gather_col(lf_H_des, 12, B, i, lfd); gather_col(rf_H_des, 12, B, i, rfd);
        gather_col(lf_twist, 6, B, i, ltw); gather_col(rf_twist, 6, B, i, rtw);
        gather_col(com_pos_des, 3, B, i, cpd); gather_col(com_vel_des, 3, B, i, cvd); gather_col(com_meas, 3, B, i, cm);
        gather_col(neck_R_des, 9, B, i, nRd);
        wco_fk(&b->model, ql, bH, left_is_fixed[i] != 0, lfH, rfH, nR, com, Jl, Jr, Jn, Jc);
        wco_ikosqp_solve(h, qm, lfH, rfH, nR, cm, Jl, Jr, Jn, Jc, lfd, rfd, ltw, rtw, cpd, cvd, nRd, d, le, re, &st, &it, &fl);
        for (int r = 0; r < nJ; r++) dq[(size_t)r * B + i] = d[r];
        if (lf_err) for (int r = 0; r < 6; r++) lf_err[(size_t)r * B + i] = le[r];
        if (rf_err) for (int r = 0; r < 6; r++) rf_err[(size_t)r * B + i] = re[r];
        if (status) status[i] = st;
        if (iters) iters[i] = it;
        if (flags) flags[i] = fl;
        if (setups) setups[i] = h->n_setups;
    }
}
