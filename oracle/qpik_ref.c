/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
 *
 * qpik_ref.c: CPU restatement of the reference QP inverse kinematics,
 *   WalkingQPIK            modules/Walking_module/src/WalkingQPInverseKinematics.cpp:25-256
 *   WalkingQPIK_qpOASES    modules/Walking_module/src/WalkingQPInverseKinematics_qpOASES.cpp:23-401
 * and of the call glue WalkingModule::solveQPIK (modules/Walking_module/src/WalkingModule.cpp:367-425).
 *
 * The QP   min 1/2 v'Hv + g'v   s.t.  A v = b,  lb <= v <= ub   is solved in the reference by
 * qpOASES::SQProblem::{init,hotstart} (…_qpOASES.cpp:312-335), a third-party dependency that is
 * absent from /root/reference and carries no version pin (cmake/FindqpOASES.cmake).  qpOASES is an
 * online active-set method that returns the exact KKT point; because the minimiser is unique
 * whenever H is positive definite on null(A), any exact method yields the same solution.  The
 * stand-in here: Householder QR of A' -> particular solution + orthonormal null-space basis Z,
 * Cholesky of Z'HZ, then the Goldfarb-Idnani dual active-set method (Math. Programming 27, 1983)
 * on the bound constraints in the reduced space; every constraint added to or dropped from the
 * working set counts as one working-set recalculation (nWSR, limit 100 as in the reference).
 * Status codes are the WC_QP_* values of include/wc_capi.h.
 * PARITY UNPINNED; optimality is verified through the KKT conditions in tests/test_oracle_ik.py.
 */
#include "oracle_internal.h"
#include "wc_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define QP_MAXN 40
#define QP_INF 1e20          /* qpOASES INFTY: bounds beyond are "no bound" (…_qpOASES.cpp:39-43 uses DBL_MAX) */
#define QP_VIOL_TOL 1e-10
#define QP_RANK_TOL 1e-11
#define QP_EPS 2.220446049250313e-16

/* ---- exact dense QP ------------------------------------------------------------------ */
int wco_qp_solve(int n, const double* H, const double* g, int mE, const double* A, const double* b,
                 const double* lb, const double* ub, int nwsr_max, double* x, int* nwsr_out)
{
    double T[QP_MAXN][QP_MAXN], V[QP_MAXN][QP_MAXN], beta[QP_MAXN];
    double Z[QP_MAXN][QP_MAXN], xp[QP_MAXN], y[QP_MAXN];
    const int nz = n - mE;
    int i, j, k, c, nwsr = 0;
    if (nwsr_out) *nwsr_out = 0;
    for (i = 0; i < n; i++) x[i] = 0.0;
    if (n > QP_MAXN || mE > n) return WC_QP_INFEASIBLE;
    /* 1. Householder QR of A' (n x mE) */
    double amax = 0.0;
    for (i = 0; i < n; i++) for (k = 0; k < mE; k++) { T[i][k] = A[k * n + i]; if (fabs(T[i][k]) > amax) amax = fabs(T[i][k]); }
    for (k = 0; k < mE; k++) {
        double sig = 0.0;
        for (i = k; i < n; i++) sig += T[i][k] * T[i][k];
        const double nrm = sqrt(sig);
        if (nrm <= QP_RANK_TOL * (amax > 1.0 ? amax : 1.0)) return WC_QP_INFEASIBLE; /* dependent equality rows */
        const double alpha = T[k][k] > 0 ? -nrm : nrm;
        for (i = 0; i < n; i++) V[i][k] = 0.0;
        for (i = k; i < n; i++) V[i][k] = T[i][k];
        V[k][k] = T[k][k] - alpha;
        beta[k] = 1.0 / (sig - T[k][k] * alpha);
        T[k][k] = alpha;
        for (i = k + 1; i < n; i++) T[i][k] = 0.0;
        for (c = k + 1; c < mE; c++) {
            double tau = 0.0;
            for (i = k; i < n; i++) tau += V[i][k] * T[i][c];
            tau *= beta[k];
            for (i = k; i < n; i++) T[i][c] -= tau * V[i][k];
        }
    }
    /* 2. R' y = b */
    for (k = 0; k < mE; k++) {
        double s = b[k];
        for (i = 0; i < k; i++) s -= T[i][k] * y[i];
        y[k] = s / T[k][k];
    }
    /* 3. xp = Q [y;0],  Z = Q [0;I] */
    for (i = 0; i < n; i++) xp[i] = i < mE ? y[i] : 0.0;
    for (c = 0; c < nz; c++) for (i = 0; i < n; i++) Z[i][c] = (i == mE + c) ? 1.0 : 0.0;
    for (k = mE - 1; k >= 0; k--) {
        double tau = 0.0;
        for (i = k; i < n; i++) tau += V[i][k] * xp[i];
        tau *= beta[k];
        for (i = k; i < n; i++) xp[i] -= tau * V[i][k];
        for (c = 0; c < nz; c++) {
            tau = 0.0;
            for (i = k; i < n; i++) tau += V[i][k] * Z[i][c];
            tau *= beta[k];
            for (i = k; i < n; i++) Z[i][c] -= tau * V[i][k];
        }
    }
    if (nz == 0) { for (i = 0; i < n; i++) x[i] = xp[i]; goto check_bounds_only; }
    {
        /* 4. reduced problem */
        double HZ[QP_MAXN][QP_MAXN], M[QP_MAXN][QP_MAXN], hr[QP_MAXN], t[QP_MAXN], w[QP_MAXN];
        double L[QP_MAXN][QP_MAXN], J[QP_MAXN][QP_MAXN], R[QP_MAXN][QP_MAXN];
        for (i = 0; i < n; i++)
            for (c = 0; c < nz; c++) {
                double s = 0.0;
                for (j = 0; j < n; j++) s += H[i * n + j] * Z[j][c];
                HZ[i][c] = s;
            }
        for (i = 0; i < n; i++) {
            double s = g[i];
            for (j = 0; j < n; j++) s += H[i * n + j] * xp[j];
            t[i] = s;
        }
        for (k = 0; k < nz; k++) {
            for (c = 0; c < nz; c++) {
                double s = 0.0;
                for (i = 0; i < n; i++) s += Z[i][k] * HZ[i][c];
                M[k][c] = s;
            }
            double s = 0.0;
            for (i = 0; i < n; i++) s += Z[i][k] * t[i];
            hr[k] = s;
        }
        /* Cholesky M = L L' */
        double mmax = 0.0;
        for (k = 0; k < nz; k++) if (fabs(M[k][k]) > mmax) mmax = fabs(M[k][k]);
        for (j = 0; j < nz; j++) {
            double d = M[j][j];
            for (k = 0; k < j; k++) d -= L[j][k] * L[j][k];
            if (!(d > 1e-13 * (mmax > 1.0 ? mmax : 1.0))) return WC_QP_NOT_CONVEX;
            d = sqrt(d);
            L[j][j] = d;
            for (i = j + 1; i < nz; i++) {
                double s = M[i][j];
                for (k = 0; k < j; k++) s -= L[i][k] * L[j][k];
                L[i][j] = s / d;
            }
        }
        /* w = -M^-1 hr */
        for (i = 0; i < nz; i++) { double s = -hr[i]; for (k = 0; k < i; k++) s -= L[i][k] * w[k]; w[i] = s / L[i][i]; }
        for (i = nz - 1; i >= 0; i--) { double s = w[i]; for (k = i + 1; k < nz; k++) s -= L[k][i] * w[k]; w[i] = s / L[i][i]; }
        /* J = L^-T (upper triangular): column c solves L' J[:,c] = e_c */
        for (c = 0; c < nz; c++) {
            for (i = nz - 1; i >= 0; i--) {
                double s = (i == c) ? 1.0 : 0.0;
                for (k = i + 1; k < nz; k++) s -= L[k][i] * J[k][c];
                J[i][c] = s / L[i][i];
            }
        }
        /* 5. Goldfarb-Idnani on the bounds.  constraint id: 2*j (lower of var j), 2*j+1 (upper) */
        int act[QP_MAXN], q = 0;
        double u[QP_MAXN], np[QP_MAXN], d[QP_MAXN], z[QP_MAXN], r[QP_MAXN];
        unsigned char is_act[2 * QP_MAXN];
        memset(is_act, 0, sizeof(is_act));
        for (;;) {
            /* most violated inactive bound */
            int p = -1;
            double sp = -QP_VIOL_TOL;
            for (j = 0; j < n; j++) {
                double xj = xp[j];
                for (c = 0; c < nz; c++) xj += Z[j][c] * w[c];
                x[j] = xj;
                if (lb[j] > -QP_INF && !is_act[2 * j]) { const double s = xj - lb[j]; if (s < sp) { sp = s; p = 2 * j; } }
                if (ub[j] < QP_INF && !is_act[2 * j + 1]) { const double s = ub[j] - xj; if (s < sp) { sp = s; p = 2 * j + 1; } }
            }
            if (p < 0) break; /* optimal */
            const int pv = p >> 1;
            const double sgn = (p & 1) ? -1.0 : 1.0;
            for (c = 0; c < nz; c++) np[c] = sgn * Z[pv][c];
            double up_ = 0.0;
            int added = 0;
            while (!added) {
                /* d = J' np ; z = J2 d2 ; r = R^-1 d1 */
                for (c = 0; c < nz; c++) { double s = 0.0; for (i = 0; i < nz; i++) s += J[i][c] * np[i]; d[c] = s; }
                for (i = 0; i < nz; i++) { double s = 0.0; for (c = q; c < nz; c++) s += J[i][c] * d[c]; z[i] = s; }
                for (i = q - 1; i >= 0; i--) { double s = d[i]; for (k = i + 1; k < q; k++) s -= R[i][k] * r[k]; r[i] = s / R[i][i]; }
                /* z'np = |d2|^2; the step direction vanishes when np is in the span of the active normals */
                double zn = 0.0, dd = 0.0;
                for (i = 0; i < nz; i++) zn += z[i] * np[i];
                for (c = 0; c < nz; c++) dd += d[c] * d[c];
                const int have_z = zn > 1e-12 * dd && zn > 0.0;
                const double t2 = have_z ? -sp / zn : INFINITY;
                double t1 = INFINITY; int l = -1;
                for (k = 0; k < q; k++) if (r[k] > 0.0 && u[k] / r[k] < t1) { t1 = u[k] / r[k]; l = k; }
                const double tt = t1 < t2 ? t1 : t2;
                if (!(tt < INFINITY)) { if (nwsr_out) *nwsr_out = nwsr; for (i = 0; i < n; i++) x[i] = 0.0; return WC_QP_INFEASIBLE; }
                if (have_z) for (i = 0; i < nz; i++) w[i] += tt * z[i];
                for (k = 0; k < q; k++) u[k] -= tt * r[k];
                up_ += tt;
                if (have_z && t2 <= t1) {
                    /* full step: add p (Givens on d[q..], applied to the columns of J) */
                    for (j = nz - 1; j > q; j--) {
                        const double a0 = d[j - 1], a1 = d[j];
                        const double hh = hypot(a0, a1);
                        if (hh == 0.0) continue;
                        const double cs = a0 / hh, sn = a1 / hh;
                        d[j - 1] = hh; d[j] = 0.0;
                        for (i = 0; i < nz; i++) {
                            const double j0 = J[i][j - 1], j1 = J[i][j];
                            J[i][j - 1] = cs * j0 + sn * j1;
                            J[i][j] = -sn * j0 + cs * j1;
                        }
                    }
                    for (i = 0; i <= q; i++) R[i][q] = d[i];
                    act[q] = p; u[q] = up_; is_act[p] = 1; q++;
                    added = 1;
                } else {
                    /* drop constraint l: delete column l of R, re-triangularise with Givens on rows */
                    is_act[act[l]] = 0;
                    for (k = l; k < q - 1; k++) {
                        act[k] = act[k + 1]; u[k] = u[k + 1];
                        for (i = 0; i <= k + 1; i++) R[i][k] = R[i][k + 1];
                    }
                    q--;
                    for (k = l; k < q; k++) {
                        const double a0 = R[k][k], a1 = R[k + 1][k];
                        const double hh = hypot(a0, a1);
                        if (hh == 0.0) continue;
                        const double cs = a0 / hh, sn = a1 / hh;
                        R[k][k] = hh; R[k + 1][k] = 0.0;
                        for (c = k + 1; c < q; c++) {
                            const double r0 = R[k][c], r1 = R[k + 1][c];
                            R[k][c] = cs * r0 + sn * r1;
                            R[k + 1][c] = -sn * r0 + cs * r1;
                        }
                        for (i = 0; i < nz; i++) {
                            const double j0 = J[i][k], j1 = J[i][k + 1];
                            J[i][k] = cs * j0 + sn * j1;
                            J[i][k + 1] = -sn * j0 + cs * j1;
                        }
                    }
                    /* slack of p at the new point */
                    double xj = xp[pv];
                    for (c = 0; c < nz; c++) xj += Z[pv][c] * w[c];
                    sp = (p & 1) ? ub[pv] - xj : xj - lb[pv];
                }
                nwsr++;
                if (nwsr > nwsr_max) { if (nwsr_out) *nwsr_out = nwsr; for (i = 0; i < n; i++) x[i] = 0.0; return WC_QP_MAX_NWSR_REACHED; }
            }
        }
        if (nwsr_out) *nwsr_out = nwsr;
        return WC_QP_SUCCESSFUL_RETURN;
    }
check_bounds_only:
    for (j = 0; j < n; j++)
        if ((lb[j] > -QP_INF && x[j] < lb[j] - QP_VIOL_TOL) || (ub[j] < QP_INF && x[j] > ub[j] + QP_VIOL_TOL)) {
            for (i = 0; i < n; i++) x[i] = 0.0;
            return WC_QP_INFEASIBLE;
        }
    return WC_QP_SUCCESSFUL_RETURN;
}

/* ---- WalkingQPIK_qpOASES ---------------------------------------------------------------- */
static void mat3_mul_bt(const double* a, const double* b, double* o) /* o = a * b' */
{
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) o[3 * r + c] = a[3 * r] * b[3 * c] + a[3 * r + 1] * b[3 * c + 1] + a[3 * r + 2] * b[3 * c + 2];
}
static void mat3_mul(const double* a, const double* b, double* o)
{
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) o[3 * r + c] = a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c];
}
/* unskew( 1/2 (M - M') ), iDynTreeHelper::Rotation::skewSymmetric (Utils.cpp:22-27) + iDynTree::unskew */
static void rot_error(const double* R, const double* Rdes, double* e)
{
    double M[9];
    mat3_mul_bt(R, Rdes, M);
    e[0] = 0.5 * (M[7] - M[5]);
    e[1] = 0.5 * (M[2] - M[6]);
    e[2] = 0.5 * (M[3] - M[1]);
}

/* H, g, A (nC x nV, task rows), b of the QP-IK from caller-supplied kinematics (all row-major, single instance).
 * variant 0 = WalkingQPIK_qpOASES; variant 1 = WalkingQPIK_osqp, which differs in two places of the build
 * (SURVEY quirk 10): the neck term of the gradient carries an extra k_attFoot factor (…_osqp.cpp:183,193) and a foot whose
 * desired twist has twist(0) == twist(1) == 0 gets the bare twist as bound, without the pose correction (:286-306). */
void wco_ik_build_qp(const wc_ik_config* cfg, int variant,
                     const double* q_meas, const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                     const double* J_lf, const double* J_rf, const double* J_neck6, const double* J_com,
                     const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                     const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                     double* H, double* g, double* A, double* b)
{
    const int nJ = cfg->n_dofs, nV = nJ + 6;
    const double* Jn = J_neck6 + 3 * nV; /* setNeckJacobian keeps rows 3..5 (QPIK.cpp:214-215) */
    int i, j, r, c;
    /* ---- setHessianMatrix (…_qpOASES.cpp:135-153) */
    double WJ[3 * QP_MAXN];
    for (r = 0; r < 3; r++)
        for (j = 0; j < nV; j++)
            WJ[r * nV + j] = cfg->neck_weight[3 * r] * Jn[j] + cfg->neck_weight[3 * r + 1] * Jn[nV + j] + cfg->neck_weight[3 * r + 2] * Jn[2 * nV + j];
    for (i = 0; i < nV; i++)
        for (j = 0; j < nV; j++) {
            double s = (i == j && i >= 6) ? cfg->joint_reg_weight[i - 6] : 0.0;
            for (r = 0; r < 3; r++) s += Jn[r * nV + i] * WJ[r * nV + j];
            H[i * nV + j] = s;
        }
    double WcJ[3 * QP_MAXN];
    if (!cfg->use_com_as_constraint) {
        for (r = 0; r < 3; r++)
            for (j = 0; j < nV; j++)
                WcJ[r * nV + j] = cfg->com_weight[3 * r] * J_com[j] + cfg->com_weight[3 * r + 1] * J_com[nV + j] + cfg->com_weight[3 * r + 2] * J_com[2 * nV + j];
        for (i = 0; i < nV; i++)
            for (j = 0; j < nV; j++)
                for (r = 0; r < 3; r++) H[i * nV + j] += J_com[r * nV + i] * WcJ[r * nV + j];
    }
    /* ---- setGradientVector (:155-182) */
    double Rdes[9], en[3], We[3];
    mat3_mul(neck_R_des, cfg->additional_rotation, Rdes); /* setDesiredNeckOrientation (QPIK.cpp:143-146) */
    rot_error(neck_R, Rdes, en);
    const double kn = variant == 1 ? cfg->k_att_foot : 1.0;
    for (r = 0; r < 3; r++)
        We[r] = cfg->neck_weight[3 * r] * (kn * (-cfg->k_neck * en[0])) + cfg->neck_weight[3 * r + 1] * (kn * (-cfg->k_neck * en[1])) +
                cfg->neck_weight[3 * r + 2] * (kn * (-cfg->k_neck * en[2]));
    for (i = 0; i < nV; i++) {
        double s = -(Jn[i] * We[0] + Jn[nV + i] * We[1] + Jn[2 * nV + i] * We[2]);
        if (i >= 6) s -= cfg->joint_reg_weight[i - 6] * (cfg->joint_reg_gain[i - 6] * (cfg->joint_reg_rad[i - 6] - q_meas[i - 6]));
        if (!cfg->use_com_as_constraint) {
            double Wv[3];
            for (r = 0; r < 3; r++)
                Wv[r] = cfg->com_weight[3 * r] * com_vel_des[0] + cfg->com_weight[3 * r + 1] * com_vel_des[1] + cfg->com_weight[3 * r + 2] * com_vel_des[2];
            s -= J_com[i] * Wv[0] + J_com[nV + i] * Wv[1] + J_com[2 * nV + i] * Wv[2];
        }
        g[i] = s;
    }
    /* ---- setLinearConstraintMatrix (:184-214) */
    for (r = 0; r < 6; r++) for (c = 0; c < nV; c++) { A[r * nV + c] = J_lf[r * nV + c]; A[(6 + r) * nV + c] = J_rf[r * nV + c]; }
    if (cfg->use_com_as_constraint)
        for (r = 0; r < 3; r++) for (c = 0; c < nV; c++) A[(12 + r) * nV + c] = J_com[r * nV + c];
    /* ---- setBounds (:216-282) */
    {
        const double* Hs[2] = {lf_H, rf_H};
        const double* Hd[2] = {lf_H_des, rf_H_des};
        const double* tw[2] = {lf_twist, rf_twist};
        for (int f = 0; f < 2; f++) {
            double R[9], Rd[9], e[3];
            for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) { R[3 * r + c] = Hs[f][4 * r + c]; Rd[3 * r + c] = Hd[f][4 * r + c]; }
            rot_error(R, Rd, e);
            const int bare = variant == 1 && tw[f][0] == tw[f][1] && tw[f][0] == 0.0;
            for (r = 0; r < 3; r++) {
                b[6 * f + r] = bare ? tw[f][r] : tw[f][r] - cfg->k_pos_foot * (Hs[f][4 * r + 3] - Hd[f][4 * r + 3]);
                b[6 * f + 3 + r] = bare ? tw[f][3 + r] : tw[f][3 + r] - cfg->k_att_foot * e[r];
            }
        }
        if (cfg->use_com_as_constraint)
            for (r = 0; r < 3; r++) b[12 + r] = com_vel_des[r] - cfg->k_pos_com * (com_meas[r] - com_pos_des[r]);
    }
}

/* One QP-IK solve with caller-supplied kinematics (all row-major, single instance). */
int wco_ik_solve(const wc_ik_config* cfg,
                 const double* q_meas, const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                 const double* J_lf, const double* J_rf, const double* J_neck6, const double* J_com,
                 const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                 const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                 double* dq, double* lf_err, double* rf_err, int* nwsr,
                 double* H_out, double* g_out, double* A_out, double* b_out, double* v_out)
{
    const int nJ = cfg->n_dofs, nV = nJ + 6;
    const int nC = cfg->use_com_as_constraint ? 15 : 12;
    double H[QP_MAXN * QP_MAXN], g[QP_MAXN], A[15 * QP_MAXN], b[15], lb[QP_MAXN], ub[QP_MAXN], v[QP_MAXN];
    int i, r, c;
    wco_ik_build_qp(cfg, 0, q_meas, lf_H, rf_H, neck_R, com_meas, J_lf, J_rf, J_neck6, J_com, lf_H_des, rf_H_des, lf_twist, rf_twist,
                    com_pos_des, com_vel_des, neck_R_des, H, g, A, b);
    /* ---- setVelocityBounds (:23-51) */
    for (i = 0; i < 6; i++) { lb[i] = -DBL_MAX; ub[i] = DBL_MAX; }
    for (i = 0; i < nJ; i++) { lb[6 + i] = cfg->min_vel[i]; ub[6 + i] = cfg->max_vel[i]; }
    /* ---- solve (:284-339): qpOASES init/hotstart -> exact QP */
    int n_wsr = 0;
    const int st = wco_qp_solve(nV, H, g, nC, A, b, lb, ub, cfg->nwsr, v, &n_wsr);
    if (nwsr) *nwsr = n_wsr;
    for (i = 0; i < nJ; i++) dq[i] = (st == WC_QP_SUCCESSFUL_RETURN) ? v[6 + i] : 0.0;
    /* getLeftFootError / getRightFootError (:364-401) */
    for (r = 0; r < 6; r++) {
        double sl = b[r], sr = b[6 + r];
        for (c = 0; c < nV; c++) { sl -= J_lf[r * nV + c] * v[c]; sr -= J_rf[r * nV + c] * v[c]; }
        if (lf_err) lf_err[r] = (st == WC_QP_SUCCESSFUL_RETURN) ? sl : 0.0;
        if (rf_err) rf_err[r] = (st == WC_QP_SUCCESSFUL_RETURN) ? sr : 0.0;
    }
    if (H_out) memcpy(H_out, H, sizeof(double) * (size_t)nV * (size_t)nV);
    if (g_out) memcpy(g_out, g, sizeof(double) * (size_t)nV);
    if (A_out) memcpy(A_out, A, sizeof(double) * (size_t)nC * (size_t)nV);
    if (b_out) memcpy(b_out, b, sizeof(double) * (size_t)nC);
    if (v_out) memcpy(v_out, v, sizeof(double) * (size_t)nV);
    return st;
}

/* ---- batched FK + QP-IK tick, SoA layout identical to wc_ik_tick_host ------------------- */
typedef struct {
    const wc_ik_config* cfg; const wc_model* model; int B;
    const double *q_meas, *q_lin, *base_H, *lf_H_des, *rf_H_des, *lf_twist, *rf_twist;
    const double *com_pos_des, *com_vel_des, *com_meas, *neck_R_des, *kpos, *katt;
    const unsigned char* left_is_fixed;
    double *dq, *lf_err, *rf_err; int *status, *nwsr;
    int begin, end;
} ik_job;

static void gather(const double* src, int rows, int B, int i, double* dst)
{
    for (int r = 0; r < rows; r++) dst[r] = src[(size_t)r * B + i];
}

static void* ik_worker(void* arg)
{
    ik_job* j = (ik_job*)arg;
    const int B = j->B, nJ = j->cfg->n_dofs, nV = nJ + 6;
    for (int i = j->begin; i < j->end; i++) {
        double qm[WC_MAX_DOFS], ql[WC_MAX_DOFS], bH[12], lfd[12], rfd[12], ltw[6], rtw[6], cpd[3], cvd[3], cm[3], nRd[9];
        double lfH[12], rfH[12], nR[9], com[3];
        double Jl[6 * QP_MAXN], Jr[6 * QP_MAXN], Jn[6 * QP_MAXN], Jc[3 * QP_MAXN];
        double dq[WC_MAX_DOFS], le[6], re[6];
        int nw = 0;
        gather(j->q_meas, nJ, B, i, qm); gather(j->q_lin, nJ, B, i, ql); gather(j->base_H, 12, B, i, bH);
        gather(j->lf_H_des, 12, B, i, lfd); gather(j->rf_H_des, 12, B, i, rfd);
        gather(j->lf_twist, 6, B, i, ltw); gather(j->rf_twist, 6, B, i, rtw);
        gather(j->com_pos_des, 3, B, i, cpd); gather(j->com_vel_des, 3, B, i, cvd); gather(j->com_meas, 3, B, i, cm);
        gather(j->neck_R_des, 9, B, i, nRd);
        wco_fk(j->model, ql, bH, j->left_is_fixed[i] != 0, lfH, rfH, nR, com, Jl, Jr, Jn, Jc);
        (void)nV;
        wc_ik_config cfg = *j->cfg;
        if (j->kpos) cfg.k_pos_foot = j->kpos[i];
        if (j->katt) cfg.k_att_foot = j->katt[i];
        const int st = wco_ik_solve(&cfg, qm, lfH, rfH, nR, cm, Jl, Jr, Jn, Jc, lfd, rfd, ltw, rtw, cpd, cvd, nRd,
                                    dq, le, re, &nw, NULL, NULL, NULL, NULL, NULL);
        for (int r = 0; r < nJ; r++) j->dq[(size_t)r * B + i] = dq[r];
        if (j->lf_err) for (int r = 0; r < 6; r++) j->lf_err[(size_t)r * B + i] = le[r];
        if (j->rf_err) for (int r = 0; r < 6; r++) j->rf_err[(size_t)r * B + i] = re[r];
        if (j->status) j->status[i] = st;
        if (j->nwsr) j->nwsr[i] = nw;
    }
    return NULL;
}

void wco_ik_batch_tick(const wc_ik_config* cfg, const wc_model* model, int B,
                       const double* q_meas, const double* q_lin, const double* base_H, const unsigned char* left_is_fixed,
                       const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                       const double* com_pos_des, const double* com_vel_des, const double* com_meas, const double* neck_R_des,
                       const double* kpos, const double* katt,
                       double* dq, double* lf_err, double* rf_err, int* status, int* nwsr, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > B) n_threads = B > 0 ? B : 1;
    ik_job* jobs = (ik_job*)calloc((size_t)n_threads, sizeof(ik_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        ik_job jb = {cfg, model, B, q_meas, q_lin, base_H, lf_H_des, rf_H_des, lf_twist, rf_twist,
                     com_pos_des, com_vel_des, com_meas, neck_R_des, kpos, katt, left_is_fixed,
                     dq, lf_err, rf_err, status, nwsr,
                     (int)((long long)B * t / n_threads), (int)((long long)B * (t + 1) / n_threads)};
        jobs[t] = jb;
        if (t > 0) pthread_create(&th[t], NULL, ik_worker, &jobs[t]);
    }
    ik_worker(&jobs[0]);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    free(jobs); free(th);
}

/* batched FK only (SoA in/out), for tests of wc_fk_dev */
void wco_fk_batch(const wc_model* model, int B, const double* q, const double* base_H, const unsigned char* left_is_fixed,
                  double* lf_H, double* rf_H, double* neck_R, double* com,
                  double* J_lf, double* J_rf, double* J_neck, double* J_com)
{
    const int nJ = model->n_dofs, nV = nJ + 6;
    for (int i = 0; i < B; i++) {
        double ql[WC_MAX_DOFS], bH[12], a[12], b2[12], nR[9], c3[3];
        double Jl[6 * QP_MAXN], Jr[6 * QP_MAXN], Jn[6 * QP_MAXN], Jc[3 * QP_MAXN];
        gather(q, nJ, B, i, ql); gather(base_H, 12, B, i, bH);
        wco_fk(model, ql, bH, left_is_fixed[i] != 0, a, b2, nR, c3, Jl, Jr, Jn, Jc);
        if (lf_H) for (int r = 0; r < 12; r++) lf_H[(size_t)r * B + i] = a[r];
        if (rf_H) for (int r = 0; r < 12; r++) rf_H[(size_t)r * B + i] = b2[r];
        if (neck_R) for (int r = 0; r < 9; r++) neck_R[(size_t)r * B + i] = nR[r];
        if (com) for (int r = 0; r < 3; r++) com[(size_t)r * B + i] = c3[r];
        if (J_lf) for (int r = 0; r < 6 * nV; r++) J_lf[(size_t)r * B + i] = Jl[r];
        if (J_rf) for (int r = 0; r < 6 * nV; r++) J_rf[(size_t)r * B + i] = Jr[r];
        if (J_neck) for (int r = 0; r < 6 * nV; r++) J_neck[(size_t)r * B + i] = Jn[r];
        if (J_com) for (int r = 0; r < 3 * nV; r++) J_com[(size_t)r * B + i] = Jc[r];
    }
}
