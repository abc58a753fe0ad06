/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
 *
 * osqp_ref.c: CPU restatement of the OSQP ADMM solver (Stellato et al., "OSQP: an operator
 * splitting solver for quadratic programs", Math. Prog. Comp. 2020) as the reference uses
 * it through osqp-eigen: OsqpEigen::Solver::{initSolver, solve, updateGradient, updateBounds,
 * getSolution}, call sites modules/Walking_module/src/MPCSolver.cpp:105,114,157,166-173,249,
 * 258,297,309,314.  OSQP/osqp-eigen are third-party dependencies that are NOT vendored under
 * /root/reference and carry no version pin (find_package(OsqpEigen REQUIRED),
 * modules/Walking_module/CMakeLists.txt:20).  This file restates the published algorithm with
 * the constants and control flow of the osqp 0.6.0 release (defaults untouched by the
 * reference except verbose=0, MPCSolver.cpp:51):
 *   setup  : Ruiz equilibration (scaling passes) with cost normalisation, rho vector with
 *            1e3*rho on equality rows, LDL^T of the quasi-definite KKT matrix, cold start;
 *   solve  : x/z/y ADMM updates with relaxation alpha, termination test every
 *            check_termination iterations on UNSCALED residuals, primal/dual infeasibility
 *            certificates, adaptive rho with re-factorisation;
 *   update : lin-cost and bounds updates re-scale with the stored D, E, c.
 * Deliberate deviation (documented in DESIGN.md): adaptive_rho_interval = 0 means "derive
 * from measured setup/solve time" in a profiling build of OSQP, which is not deterministic;
 * here 0 selects OSQP's own non-profiling fallback ADAPTIVE_RHO_MULTIPLE_TERMINATION(4) *
 * check_termination = 100.
 *
 * PARITY UNPINNED: the reference holds no golden vectors / tests for this path and OSQP
 * itself cannot be built here; the restatement is cross-checked by independent means
 * (exact KKT solutions, scipy) in tests/.
 */
#include "oracle_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_EQ_OVER_RHO_INEQ 1e3
#define RHO_TOL 1e-4
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define ADAPTIVE_RHO_MULTIPLE_TERMINATION 4
#define ADAPTIVE_RHO_FIXED 100

struct wco_osqp {
    int n, m;
    wc_osqp_settings s;
    int adaptive_rho_interval; /* resolved lazily in solve, then sticky (as OSQP does) */
    wco_csc *P, *A;            /* scaled copies */
    double *q, *l, *u;         /* scaled */
    double *D, *E, *Dinv, *Einv, c, cinv;
    double *rho_vec, *rho_inv_vec;
    int* constr_type;
    double *x, *y, *z, *xz_tilde, *x_prev, *z_prev;
    double *Ax, *Px, *Aty, *delta_y, *Atdelta_y, *delta_x, *Pdelta_x, *Adelta_x;
    double *D_temp, *D_temp_A, *E_temp;
    double *sol_x, *sol_y;
    /* KKT */
    int* perm;     /* perm[k] = original KKT index at position k */
    int* iperm;    /* inverse */
    wco_csc* KKT;  /* permuted upper triangle */
    int* rho_pos;  /* position in KKT->x of the -1/rho diagonal of constraint i */
    wco_ldl* ldl;
    double* kkt_rhs;
    int linsys_kind;
    double* Kdense; /* reduced system Cholesky (linsys_kind = 1) */
    /* info */
    int status, iter, rho_updates;
    double pri_res, dua_res, rho_estimate, obj_val;
};

/* ------------------------------------------------------------------ CSC helpers */
wco_csc* wco_csc_from_triplets(int nrow, int ncol, int nt, const int* ti, const int* tj, const double* tx)
{
    wco_csc* a = (wco_csc*)calloc(1, sizeof(wco_csc));
    int k, j;
    a->nrow = nrow; a->ncol = ncol;
    a->p = (int*)calloc((size_t)ncol + 1, sizeof(int));
    a->i = (int*)malloc(sizeof(int) * (size_t)(nt > 0 ? nt : 1));
    a->x = (double*)malloc(sizeof(double) * (size_t)(nt > 0 ? nt : 1));
    /* count, then insertion keeping rows sorted inside a column; duplicates are summed */
    int* cnt = (int*)calloc((size_t)ncol + 1, sizeof(int));
    for (k = 0; k < nt; k++) cnt[tj[k]]++;
    for (j = 0; j < ncol; j++) a->p[j + 1] = a->p[j] + cnt[j];
    int* fill = (int*)calloc((size_t)ncol, sizeof(int));
    for (k = 0; k < nt; k++) {
        int col = tj[k], row = ti[k];
        int base = a->p[col], cntc = fill[col], pos, dup = -1;
        for (pos = 0; pos < cntc; pos++)
            if (a->i[base + pos] == row) { dup = pos; break; }
        if (dup >= 0) { a->x[base + dup] += tx[k]; continue; }
        pos = cntc;
        while (pos > 0 && a->i[base + pos - 1] > row) {
            a->i[base + pos] = a->i[base + pos - 1];
            a->x[base + pos] = a->x[base + pos - 1];
            pos--;
        }
        a->i[base + pos] = row;
        a->x[base + pos] = tx[k];
        fill[col]++;
    }
    /* compact (duplicates leave holes) */
    int w = 0;
    int* newp = (int*)calloc((size_t)ncol + 1, sizeof(int));
    for (j = 0; j < ncol; j++) {
        newp[j] = w;
        for (k = 0; k < fill[j]; k++) {
            a->i[w] = a->i[a->p[j] + k];
            a->x[w] = a->x[a->p[j] + k];
            w++;
        }
    }
    newp[ncol] = w;
    memcpy(a->p, newp, sizeof(int) * ((size_t)ncol + 1));
    a->nnz = w;
    free(newp); free(cnt); free(fill);
    return a;
}

wco_csc* wco_csc_copy(const wco_csc* s)
{
    wco_csc* a = (wco_csc*)calloc(1, sizeof(wco_csc));
    a->nrow = s->nrow; a->ncol = s->ncol; a->nnz = s->nnz;
    a->p = (int*)malloc(sizeof(int) * ((size_t)s->ncol + 1));
    a->i = (int*)malloc(sizeof(int) * (size_t)(s->nnz > 0 ? s->nnz : 1));
    a->x = (double*)malloc(sizeof(double) * (size_t)(s->nnz > 0 ? s->nnz : 1));
    memcpy(a->p, s->p, sizeof(int) * ((size_t)s->ncol + 1));
    memcpy(a->i, s->i, sizeof(int) * (size_t)s->nnz);
    memcpy(a->x, s->x, sizeof(double) * (size_t)s->nnz);
    return a;
}

void wco_csc_free(wco_csc* a)
{
    if (!a) return;
    free(a->p); free(a->i); free(a->x); free(a);
}

/* ------------------------------------------------------------------ linear algebra (osqp lin_alg.c) */
static double vec_norm_inf(const double* v, int n)
{
    double r = 0.0;
    for (int i = 0; i < n; i++) { double a = fabs(v[i]); if (a > r) r = a; }
    return r;
}
static double vec_scaled_norm_inf(const double* S, const double* v, int n)
{
    double r = 0.0;
    for (int i = 0; i < n; i++) { double a = fabs(S[i] * v[i]); if (a > r) r = a; }
    return r;
}
/* y (+)= A x ; plus_eq: 0 overwrite, 1 add, -1 subtract */
static void mat_vec(const wco_csc* A, const double* x, double* y, int plus_eq)
{
    if (!plus_eq) for (int i = 0; i < A->nrow; i++) y[i] = 0.0;
    if (A->p[A->ncol] == 0) return;
    if (plus_eq == -1) {
        for (int j = 0; j < A->ncol; j++)
            for (int p = A->p[j]; p < A->p[j + 1]; p++) y[A->i[p]] -= A->x[p] * x[j];
    } else {
        for (int j = 0; j < A->ncol; j++)
            for (int p = A->p[j]; p < A->p[j + 1]; p++) y[A->i[p]] += A->x[p] * x[j];
    }
}
/* y (+)= A' x ; skip_diag as in osqp (used for the strictly-lower part of symmetric P) */
static void mat_tpose_vec(const wco_csc* A, const double* x, double* y, int plus_eq, int skip_diag)
{
    if (!plus_eq) for (int i = 0; i < A->ncol; i++) y[i] = 0.0;
    if (A->p[A->ncol] == 0) return;
    for (int j = 0; j < A->ncol; j++) {
        for (int p = A->p[j]; p < A->p[j + 1]; p++) {
            int i = A->i[p];
            if (skip_diag && i == j) continue;
            if (plus_eq == -1) y[j] -= A->x[p] * x[i];
            else y[j] += A->x[p] * x[i];
        }
    }
}
static void mat_inf_norm_cols(const wco_csc* M, double* E)
{
    for (int j = 0; j < M->ncol; j++) {
        E[j] = 0.0;
        for (int p = M->p[j]; p < M->p[j + 1]; p++) { double a = fabs(M->x[p]); if (a > E[j]) E[j] = a; }
    }
}
static void mat_inf_norm_rows(const wco_csc* M, double* E)
{
    for (int i = 0; i < M->nrow; i++) E[i] = 0.0;
    for (int j = 0; j < M->ncol; j++)
        for (int p = M->p[j]; p < M->p[j + 1]; p++) {
            double a = fabs(M->x[p]); int i = M->i[p];
            if (a > E[i]) E[i] = a;
        }
}
static void mat_inf_norm_cols_sym_triu(const wco_csc* M, double* E)
{
    for (int j = 0; j < M->ncol; j++) E[j] = 0.0;
    for (int j = 0; j < M->ncol; j++)
        for (int p = M->p[j]; p < M->p[j + 1]; p++) {
            int i = M->i[p]; double a = fabs(M->x[p]);
            if (a > E[j]) E[j] = a;
            if (i != j && a > E[i]) E[i] = a;
        }
}
static void limit_scaling(double* D, int n)
{
    for (int i = 0; i < n; i++) {
        D[i] = D[i] < MIN_SCALING ? 1.0 : D[i];
        D[i] = D[i] > MAX_SCALING ? MAX_SCALING : D[i];
    }
}

/* ------------------------------------------------------------------ scaling (osqp scaling.c) */
static void scale_data(wco_osqp* w)
{
    const int n = w->n, m = w->m;
    int i, j, p, it;
    w->c = 1.0;
    for (i = 0; i < n; i++) { w->D[i] = 1.0; w->Dinv[i] = 1.0; }
    for (i = 0; i < m; i++) { w->E[i] = 1.0; w->Einv[i] = 1.0; }
    for (it = 0; it < w->s.scaling; it++) {
        /* inf-norm of the columns of [[P, A'],[A, 0]] */
        mat_inf_norm_cols_sym_triu(w->P, w->D_temp);
        mat_inf_norm_cols(w->A, w->D_temp_A);
        for (i = 0; i < n; i++) if (w->D_temp_A[i] > w->D_temp[i]) w->D_temp[i] = w->D_temp_A[i];
        mat_inf_norm_rows(w->A, w->E_temp);
        limit_scaling(w->D_temp, n);
        limit_scaling(w->E_temp, m);
        for (i = 0; i < n; i++) w->D_temp[i] = 1.0 / sqrt(w->D_temp[i]);
        for (i = 0; i < m; i++) w->E_temp[i] = 1.0 / sqrt(w->E_temp[i]);
        /* P <- D P D */
        for (j = 0; j < n; j++)
            for (p = w->P->p[j]; p < w->P->p[j + 1]; p++) w->P->x[p] *= w->D_temp[w->P->i[p]];
        for (j = 0; j < n; j++)
            for (p = w->P->p[j]; p < w->P->p[j + 1]; p++) w->P->x[p] *= w->D_temp[j];
        /* A <- E A D */
        for (j = 0; j < n; j++)
            for (p = w->A->p[j]; p < w->A->p[j + 1]; p++) w->A->x[p] *= w->E_temp[w->A->i[p]];
        for (j = 0; j < n; j++)
            for (p = w->A->p[j]; p < w->A->p[j + 1]; p++) w->A->x[p] *= w->D_temp[j];
        /* q <- D q */
        for (i = 0; i < n; i++) w->q[i] *= w->D_temp[i];
        for (i = 0; i < n; i++) w->D[i] *= w->D_temp[i];
        for (i = 0; i < m; i++) w->E[i] *= w->E_temp[i];
        /* cost normalisation */
        mat_inf_norm_cols_sym_triu(w->P, w->D_temp);
        double c_temp = 0.0;
        for (i = 0; i < n; i++) c_temp += w->D_temp[i];
        c_temp /= (double)n;
        double inf_norm_q = vec_norm_inf(w->q, n);
        limit_scaling(&inf_norm_q, 1);
        if (inf_norm_q > c_temp) c_temp = inf_norm_q;
        limit_scaling(&c_temp, 1);
        c_temp = 1.0 / c_temp;
        for (p = 0; p < w->P->nnz; p++) w->P->x[p] *= c_temp;
        for (i = 0; i < n; i++) w->q[i] *= c_temp;
        w->c *= c_temp;
    }
    w->cinv = 1.0 / w->c;
    for (i = 0; i < n; i++) w->Dinv[i] = 1.0 / w->D[i];
    for (i = 0; i < m; i++) w->Einv[i] = 1.0 / w->E[i];
    for (i = 0; i < m; i++) { w->l[i] *= w->E[i]; w->u[i] *= w->E[i]; }
}

/* ------------------------------------------------------------------ rho vector (osqp auxil.c) */
static void set_rho_vec(wco_osqp* w)
{
    const double inf = w->s.infty;
    double rho = w->s.rho;
    rho = rho < RHO_MIN ? RHO_MIN : (rho > RHO_MAX ? RHO_MAX : rho);
    w->s.rho = rho;
    for (int i = 0; i < w->m; i++) {
        if (w->l[i] < -inf * MIN_SCALING && w->u[i] > inf * MIN_SCALING) {
            w->constr_type[i] = -1;
            w->rho_vec[i] = RHO_MIN;
        } else if (w->u[i] - w->l[i] < RHO_TOL) {
            w->constr_type[i] = 1;
            w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * rho;
        } else {
            w->constr_type[i] = 0;
            w->rho_vec[i] = rho;
        }
        w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
    }
}

/* ------------------------------------------------------------------ KKT (osqp kkt.c + qdldl_interface.c) */
static int kkt_build(wco_osqp* w)
{
    const int n = w->n, m = w->m, N = n + m;
    int cap = w->P->nnz + n + w->A->nnz + m;
    int* ti = (int*)malloc(sizeof(int) * (size_t)cap);
    int* tj = (int*)malloc(sizeof(int) * (size_t)cap);
    double* tx = (double*)malloc(sizeof(double) * (size_t)cap);
    int* tag = (int*)malloc(sizeof(int) * (size_t)cap); /* >=0: constraint index of a -1/rho entry */
    int nt = 0, j, p;
    /* P + sigma I (upper), making sure every diagonal entry exists */
    unsigned char* has_diag = (unsigned char*)calloc((size_t)n, 1);
    for (j = 0; j < n; j++)
        for (p = w->P->p[j]; p < w->P->p[j + 1]; p++) {
            int i = w->P->i[p];
            double v = w->P->x[p];
            if (i == j) { v += w->s.sigma; has_diag[j] = 1; }
            ti[nt] = i; tj[nt] = j; tx[nt] = v; tag[nt] = -1; nt++;
        }
    for (j = 0; j < n; j++)
        if (!has_diag[j]) { ti[nt] = j; tj[nt] = j; tx[nt] = w->s.sigma; tag[nt] = -1; nt++; }
    free(has_diag);
    /* A' in the upper-right block */
    for (j = 0; j < n; j++)
        for (p = w->A->p[j]; p < w->A->p[j + 1]; p++) {
            ti[nt] = j; tj[nt] = n + w->A->i[p]; tx[nt] = w->A->x[p]; tag[nt] = -1; nt++;
        }
    for (j = 0; j < m; j++) { ti[nt] = n + j; tj[nt] = n + j; tx[nt] = -w->rho_inv_vec[j]; tag[nt] = j; nt++; }
    /* symmetric permutation, keep upper triangle */
    for (p = 0; p < nt; p++) {
        int a = w->iperm[ti[p]], b = w->iperm[tj[p]];
        if (a > b) { int t = a; a = b; b = t; }
        ti[p] = a; tj[p] = b;
    }
    wco_csc_free(w->KKT);
    w->KKT = wco_csc_from_triplets(N, N, nt, ti, tj, tx);
    /* locate the rho diagonals */
    for (j = 0; j < m; j++) {
        int col = w->iperm[n + j];
        w->rho_pos[j] = -1;
        for (p = w->KKT->p[col]; p < w->KKT->p[col + 1]; p++)
            if (w->KKT->i[p] == col) w->rho_pos[j] = p;
    }
    free(ti); free(tj); free(tx); free(tag);
    return 0;
}

static int reduced_factor(wco_osqp* w)
{
    /* dense Cholesky of P + sigma I + A' diag(rho) A (lower, in place) */
    const int n = w->n;
    int i, j, k, p, pp;
    if (!w->Kdense) w->Kdense = (double*)malloc(sizeof(double) * (size_t)n * (size_t)n);
    double* K = w->Kdense;
    memset(K, 0, sizeof(double) * (size_t)n * (size_t)n);
    for (j = 0; j < n; j++)
        for (p = w->P->p[j]; p < w->P->p[j + 1]; p++) {
            i = w->P->i[p];
            K[(size_t)i * n + j] += w->P->x[p];
            if (i != j) K[(size_t)j * n + i] += w->P->x[p];
        }
    for (j = 0; j < n; j++) K[(size_t)j * n + j] += w->s.sigma;
    for (j = 0; j < n; j++)
        for (p = w->A->p[j]; p < w->A->p[j + 1]; p++)
            for (k = 0; k < n; k++)
                for (pp = w->A->p[k]; pp < w->A->p[k + 1]; pp++)
                    if (w->A->i[pp] == w->A->i[p])
                        K[(size_t)j * n + k] += w->A->x[p] * w->rho_vec[w->A->i[p]] * w->A->x[pp];
    for (j = 0; j < n; j++) {
        double d = K[(size_t)j * n + j];
        for (k = 0; k < j; k++) d -= K[(size_t)j * n + k] * K[(size_t)j * n + k];
        if (d <= 0.0) return -1;
        d = sqrt(d);
        K[(size_t)j * n + j] = d;
        for (i = j + 1; i < n; i++) {
            double s = K[(size_t)i * n + j];
            for (k = 0; k < j; k++) s -= K[(size_t)i * n + k] * K[(size_t)j * n + k];
            K[(size_t)i * n + j] = s / d;
        }
    }
    return 0;
}

static int linsys_factor(wco_osqp* w, int rebuild)
{
    if (w->linsys_kind == 1) return reduced_factor(w);
    if (rebuild || !w->KKT) {
        kkt_build(w);
        if (!w->ldl) w->ldl = wco_ldl_symbolic(w->n + w->m, w->KKT->p, w->KKT->i);
        if (!w->ldl) return -1;
    } else {
        for (int j = 0; j < w->m; j++) w->KKT->x[w->rho_pos[j]] = -w->rho_inv_vec[j];
    }
    int pos = wco_ldl_numeric(w->ldl, w->KKT->p, w->KKT->i, w->KKT->x);
    return (pos == w->n) ? 0 : -1; /* quasi-definite: exactly n positive pivots */
}

/* b = [rhs_x ; rhs_z] -> [x_tilde ; z_tilde]  (osqp qdldl_interface.c solve_linsys_qdldl) */
static void linsys_solve(wco_osqp* w, double* b)
{
    const int n = w->n, m = w->m;
    int i, j;
    if (w->linsys_kind == 1) {
        /* K x = rhs_x + A' (rho .* rhs_z) ; z_tilde = A x */
        double* r = w->kkt_rhs;
        for (i = 0; i < m; i++) w->kkt_rhs[n + i] = w->rho_vec[i] * b[n + i];
        mat_tpose_vec(w->A, w->kkt_rhs + n, r, 0, 0);
        for (i = 0; i < n; i++) r[i] += b[i];
        const double* K = w->Kdense;
        for (i = 0; i < n; i++) {
            double s = r[i];
            for (j = 0; j < i; j++) s -= K[(size_t)i * n + j] * r[j];
            r[i] = s / K[(size_t)i * n + i];
        }
        for (i = n - 1; i >= 0; i--) {
            double s = r[i];
            for (j = i + 1; j < n; j++) s -= K[(size_t)j * n + i] * r[j];
            r[i] = s / K[(size_t)i * n + i];
        }
        for (i = 0; i < n; i++) b[i] = r[i];
        mat_vec(w->A, b, b + n, 0);
        return;
    }
    for (i = 0; i < n + m; i++) w->kkt_rhs[i] = b[w->perm[i]];
    wco_ldl_solve(w->ldl, w->kkt_rhs);
    for (j = 0; j < n; j++) b[j] = w->kkt_rhs[w->iperm[j]];
    for (j = 0; j < m; j++) b[n + j] += w->rho_inv_vec[j] * w->kkt_rhs[w->iperm[n + j]];
}

/* ------------------------------------------------------------------ setup */
static double* dvec(int n) { return (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

wco_osqp* wco_osqp_setup(const wco_csc* P, const wco_csc* A, const double* q, const double* l,
                         const double* u, const wc_osqp_settings* settings, const int* perm)
{
    const int n = P->ncol, m = A->nrow;
    int i;
    for (i = 0; i < m; i++) if (l[i] > u[i]) return NULL;
    wco_osqp* w = (wco_osqp*)calloc(1, sizeof(wco_osqp));
    w->n = n; w->m = m; w->s = *settings;
    w->adaptive_rho_interval = settings->adaptive_rho_interval;
    w->P = wco_csc_copy(P); w->A = wco_csc_copy(A);
    w->q = dvec(n); w->l = dvec(m); w->u = dvec(m);
    memcpy(w->q, q, sizeof(double) * (size_t)n);
    memcpy(w->l, l, sizeof(double) * (size_t)m);
    memcpy(w->u, u, sizeof(double) * (size_t)m);
    w->D = dvec(n); w->Dinv = dvec(n); w->E = dvec(m); w->Einv = dvec(m);
    w->rho_vec = dvec(m); w->rho_inv_vec = dvec(m);
    w->constr_type = (int*)calloc((size_t)(m > 0 ? m : 1), sizeof(int));
    w->x = dvec(n); w->y = dvec(m); w->z = dvec(m); w->xz_tilde = dvec(n + m);
    w->x_prev = dvec(n); w->z_prev = dvec(m);
    w->Ax = dvec(m); w->Px = dvec(n); w->Aty = dvec(n); w->delta_y = dvec(m); w->Atdelta_y = dvec(n);
    w->delta_x = dvec(n); w->Pdelta_x = dvec(n); w->Adelta_x = dvec(m);
    w->D_temp = dvec(n); w->D_temp_A = dvec(n); w->E_temp = dvec(m);
    w->sol_x = dvec(n); w->sol_y = dvec(m);
    w->perm = (int*)malloc(sizeof(int) * (size_t)(n + m));
    w->iperm = (int*)malloc(sizeof(int) * (size_t)(n + m));
    for (i = 0; i < n + m; i++) w->perm[i] = perm ? perm[i] : i;
    for (i = 0; i < n + m; i++) w->iperm[w->perm[i]] = i;
    w->rho_pos = (int*)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    w->kkt_rhs = dvec(n + m);
    w->c = w->cinv = 1.0;
    for (i = 0; i < n; i++) { w->D[i] = w->Dinv[i] = 1.0; }
    for (i = 0; i < m; i++) { w->E[i] = w->Einv[i] = 1.0; }
    if (w->s.scaling) scale_data(w);
    set_rho_vec(w);
    if (linsys_factor(w, 1) != 0) { wco_osqp_free(w); return NULL; }
    w->status = WC_OSQP_UNSOLVED;
    w->rho_updates = 0;
    w->rho_estimate = w->s.rho;
    /* cold start: x = z = y = 0 (calloc) */
    return w;
}

void wco_osqp_set_linsys(wco_osqp* w, int kind)
{
    w->linsys_kind = kind;
    linsys_factor(w, 1);
}

void wco_osqp_free(wco_osqp* w)
{
    if (!w) return;
    wco_csc_free(w->P); wco_csc_free(w->A); wco_csc_free(w->KKT); wco_ldl_free(w->ldl);
    free(w->q); free(w->l); free(w->u); free(w->D); free(w->E); free(w->Dinv); free(w->Einv);
    free(w->rho_vec); free(w->rho_inv_vec); free(w->constr_type);
    free(w->x); free(w->y); free(w->z); free(w->xz_tilde); free(w->x_prev); free(w->z_prev);
    free(w->Ax); free(w->Px); free(w->Aty); free(w->delta_y); free(w->Atdelta_y);
    free(w->delta_x); free(w->Pdelta_x); free(w->Adelta_x);
    free(w->D_temp); free(w->D_temp_A); free(w->E_temp); free(w->sol_x); free(w->sol_y);
    free(w->perm); free(w->iperm); free(w->rho_pos); free(w->kkt_rhs); free(w->Kdense);
    free(w);
}

static void reset_info(wco_osqp* w);
/* ------------------------------------------------------------------ matrix updates (osqp osqp.c: osqp_update_P / osqp_update_A)
 * OSQP 0.6 unscales the WHOLE problem with the current D, E, c, overwrites the matrix values, and runs scale_data again
 * (a fresh Ruiz equilibration from D = E = 1), then refactors and resets the info; the iterates x, z, y are left as they
 * are (scaled with the previous D, E, c) and rho_vec is kept. */
static void unscale_data(wco_osqp* w)
{
    const int n = w->n, m = w->m;
    int i, j, p;
    for (p = 0; p < w->P->nnz; p++) w->P->x[p] *= w->cinv;
    for (j = 0; j < n; j++)
        for (p = w->P->p[j]; p < w->P->p[j + 1]; p++) w->P->x[p] *= w->Dinv[w->P->i[p]];
    for (j = 0; j < n; j++)
        for (p = w->P->p[j]; p < w->P->p[j + 1]; p++) w->P->x[p] *= w->Dinv[j];
    for (i = 0; i < n; i++) w->q[i] *= w->cinv;
    for (i = 0; i < n; i++) w->q[i] *= w->Dinv[i];
    for (j = 0; j < n; j++)
        for (p = w->A->p[j]; p < w->A->p[j + 1]; p++) w->A->x[p] *= w->Einv[w->A->i[p]];
    for (j = 0; j < n; j++)
        for (p = w->A->p[j]; p < w->A->p[j + 1]; p++) w->A->x[p] *= w->Dinv[j];
    for (i = 0; i < m; i++) { w->l[i] *= w->Einv[i]; w->u[i] *= w->Einv[i]; }
}

int wco_osqp_update_P(wco_osqp* w, const double* Px_new)
{
    if (w->s.scaling) unscale_data(w);
    memcpy(w->P->x, Px_new, sizeof(double) * (size_t)w->P->nnz);
    if (w->s.scaling) scale_data(w);
    const int rc = linsys_factor(w, 1);
    reset_info(w);
    return rc;
}

int wco_osqp_update_A(wco_osqp* w, const double* Ax_new)
{
    if (w->s.scaling) unscale_data(w);
    memcpy(w->A->x, Ax_new, sizeof(double) * (size_t)w->A->nnz);
    if (w->s.scaling) scale_data(w);
    const int rc = linsys_factor(w, 1);
    reset_info(w);
    return rc;
}

/* the (scaled) iterates of the workspace, as OsqpEigen::Solver::getPrimalVariable / getDualVariable map them */
void wco_osqp_get_iterate(const wco_osqp* w, double* x, double* y)
{
    memcpy(x, w->x, sizeof(double) * (size_t)w->n);
    memcpy(y, w->y, sizeof(double) * (size_t)w->m);
}

/* osqp_warm_start: copy, scale with the CURRENT D, E, c, z = A x */
void wco_osqp_warm_start(wco_osqp* w, const double* x, const double* y)
{
    memcpy(w->x, x, sizeof(double) * (size_t)w->n);
    memcpy(w->y, y, sizeof(double) * (size_t)w->m);
    if (w->s.scaling) {
        for (int i = 0; i < w->n; i++) w->x[i] *= w->Dinv[i];
        for (int i = 0; i < w->m; i++) w->y[i] *= w->Einv[i];
        for (int i = 0; i < w->m; i++) w->y[i] *= w->c;
    }
    mat_vec(w->A, w->x, w->z, 0);
}

/* ------------------------------------------------------------------ updates (osqp osqp.c) */
static void reset_info(wco_osqp* w)
{
    w->status = WC_OSQP_UNSOLVED;
    w->rho_updates = 0;
}

void wco_osqp_update_lin_cost(wco_osqp* w, const double* q_new)
{
    memcpy(w->q, q_new, sizeof(double) * (size_t)w->n);
    if (w->s.scaling) {
        for (int i = 0; i < w->n; i++) w->q[i] *= w->D[i];
        for (int i = 0; i < w->n; i++) w->q[i] *= w->c;
    }
    reset_info(w);
}

static int update_rho_vec(wco_osqp* w)
{
    /* osqp auxil.c update_rho_vec: re-classify the constraints, refactor if any type changed */
    const double inf = w->s.infty;
    int changed = 0;
    for (int i = 0; i < w->m; i++) {
        if (w->l[i] < -inf * MIN_SCALING && w->u[i] > inf * MIN_SCALING) {
            if (w->constr_type[i] != -1) {
                w->constr_type[i] = -1; w->rho_vec[i] = RHO_MIN; w->rho_inv_vec[i] = 1.0 / RHO_MIN; changed = 1;
            }
        } else if (w->u[i] - w->l[i] < RHO_TOL) {
            if (w->constr_type[i] != 1) {
                w->constr_type[i] = 1; w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->s.rho;
                w->rho_inv_vec[i] = 1.0 / w->rho_vec[i]; changed = 1;
            }
        } else {
            if (w->constr_type[i] != 0) {
                w->constr_type[i] = 0; w->rho_vec[i] = w->s.rho; w->rho_inv_vec[i] = 1.0 / w->s.rho; changed = 1;
            }
        }
    }
    if (changed) return linsys_factor(w, 0);
    return 0;
}

int wco_osqp_update_bounds(wco_osqp* w, const double* l_new, const double* u_new)
{
    for (int i = 0; i < w->m; i++) if (l_new[i] > u_new[i]) return 1;
    memcpy(w->l, l_new, sizeof(double) * (size_t)w->m);
    memcpy(w->u, u_new, sizeof(double) * (size_t)w->m);
    if (w->s.scaling) for (int i = 0; i < w->m; i++) { w->l[i] *= w->E[i]; w->u[i] *= w->E[i]; }
    reset_info(w);
    return update_rho_vec(w);
}

/* ------------------------------------------------------------------ residuals & termination (osqp auxil.c) */
static double compute_pri_res(wco_osqp* w)
{
    mat_vec(w->A, w->x, w->Ax, 0);
    for (int i = 0; i < w->m; i++) w->z_prev[i] = w->Ax[i] - w->z[i];
    if (w->s.scaling) return vec_scaled_norm_inf(w->Einv, w->z_prev, w->m);
    return vec_norm_inf(w->z_prev, w->m);
}
static double compute_pri_tol(wco_osqp* w, double eps_abs, double eps_rel)
{
    double a, b;
    if (w->s.scaling) {
        a = vec_scaled_norm_inf(w->Einv, w->z, w->m);
        b = vec_scaled_norm_inf(w->Einv, w->Ax, w->m);
    } else {
        a = vec_norm_inf(w->z, w->m);
        b = vec_norm_inf(w->Ax, w->m);
    }
    return eps_abs + eps_rel * (a > b ? a : b);
}
static double compute_dua_res(wco_osqp* w)
{
    const int n = w->n;
    memcpy(w->x_prev, w->q, sizeof(double) * (size_t)n);
    mat_vec(w->P, w->x, w->Px, 0);
    mat_tpose_vec(w->P, w->x, w->Px, 1, 1);
    for (int i = 0; i < n; i++) w->x_prev[i] += w->Px[i];
    if (w->m > 0) {
        mat_tpose_vec(w->A, w->y, w->Aty, 0, 0);
        for (int i = 0; i < n; i++) w->x_prev[i] += w->Aty[i];
    }
    if (w->s.scaling) return w->cinv * vec_scaled_norm_inf(w->Dinv, w->x_prev, n);
    return vec_norm_inf(w->x_prev, n);
}
static double compute_dua_tol(wco_osqp* w, double eps_abs, double eps_rel)
{
    double r, t;
    if (w->s.scaling) {
        r = vec_scaled_norm_inf(w->Dinv, w->q, w->n);
        t = vec_scaled_norm_inf(w->Dinv, w->Aty, w->n); if (t > r) r = t;
        t = vec_scaled_norm_inf(w->Dinv, w->Px, w->n); if (t > r) r = t;
        r *= w->cinv;
    } else {
        r = vec_norm_inf(w->q, w->n);
        t = vec_norm_inf(w->Aty, w->n); if (t > r) r = t;
        t = vec_norm_inf(w->Px, w->n); if (t > r) r = t;
    }
    return eps_abs + eps_rel * r;
}
static void update_info(wco_osqp* w, int iter)
{
    w->iter = iter;
    w->pri_res = (w->m == 0) ? 0.0 : compute_pri_res(w);
    w->dua_res = compute_dua_res(w);
}
static int is_primal_infeasible(wco_osqp* w, double eps_prim_inf)
{
    const double inf = w->s.infty;
    const int n = w->n, m = w->m;
    double norm_delta_y, ineq_lhs = 0.0;
    int i;
    for (i = 0; i < m; i++) {
        if (w->u[i] > inf * MIN_SCALING) {
            if (w->l[i] < -inf * MIN_SCALING) w->delta_y[i] = 0.0;
            else w->delta_y[i] = w->delta_y[i] < 0.0 ? w->delta_y[i] : 0.0;
        } else if (w->l[i] < -inf * MIN_SCALING) {
            w->delta_y[i] = w->delta_y[i] > 0.0 ? w->delta_y[i] : 0.0;
        }
    }
    if (w->s.scaling) {
        for (i = 0; i < m; i++) w->Adelta_x[i] = w->E[i] * w->delta_y[i];
        norm_delta_y = vec_norm_inf(w->Adelta_x, m);
    } else norm_delta_y = vec_norm_inf(w->delta_y, m);
    if (norm_delta_y > 1.0 / inf) { /* OSQP_DIVISION_TOL */
        for (i = 0; i < m; i++) {
            double dy = w->delta_y[i];
            ineq_lhs += w->u[i] * (dy > 0.0 ? dy : 0.0) + w->l[i] * (dy < 0.0 ? dy : 0.0);
        }
        if (ineq_lhs < -eps_prim_inf * norm_delta_y) {
            mat_tpose_vec(w->A, w->delta_y, w->Atdelta_y, 0, 0);
            if (w->s.scaling) for (i = 0; i < n; i++) w->Atdelta_y[i] *= w->Dinv[i];
            return vec_norm_inf(w->Atdelta_y, n) < eps_prim_inf * norm_delta_y;
        }
    }
    return 0;
}
static int is_dual_infeasible(wco_osqp* w, double eps_dual_inf)
{
    const double inf = w->s.infty;
    const int n = w->n, m = w->m;
    double norm_delta_x, cost_scaling;
    int i;
    if (w->s.scaling) {
        norm_delta_x = vec_scaled_norm_inf(w->D, w->delta_x, n);
        cost_scaling = w->c;
    } else {
        norm_delta_x = vec_norm_inf(w->delta_x, n);
        cost_scaling = 1.0;
    }
    if (norm_delta_x > 1.0 / inf) {
        double qdx = 0.0;
        for (i = 0; i < n; i++) qdx += w->q[i] * w->delta_x[i];
        if (qdx < -cost_scaling * eps_dual_inf * norm_delta_x) {
            mat_vec(w->P, w->delta_x, w->Pdelta_x, 0);
            mat_tpose_vec(w->P, w->delta_x, w->Pdelta_x, 1, 1);
            if (w->s.scaling) for (i = 0; i < n; i++) w->Pdelta_x[i] *= w->Dinv[i];
            if (vec_norm_inf(w->Pdelta_x, n) < cost_scaling * eps_dual_inf * norm_delta_x) {
                mat_vec(w->A, w->delta_x, w->Adelta_x, 0);
                if (w->s.scaling) for (i = 0; i < m; i++) w->Adelta_x[i] *= w->Einv[i];
                for (i = 0; i < m; i++) {
                    if ((w->u[i] < inf * MIN_SCALING && w->Adelta_x[i] > eps_dual_inf * norm_delta_x) ||
                        (w->l[i] > -inf * MIN_SCALING && w->Adelta_x[i] < -eps_dual_inf * norm_delta_x))
                        return 0;
                }
                return 1;
            }
        }
    }
    return 0;
}
static int check_termination(wco_osqp* w, int approximate)
{
    double eps_abs = w->s.eps_abs, eps_rel = w->s.eps_rel;
    double eps_prim_inf = w->s.eps_prim_inf, eps_dual_inf = w->s.eps_dual_inf;
    int prim_res_check = 0, dual_res_check = 0, prim_inf_check = 0, dual_inf_check = 0;
    if (w->pri_res > w->s.infty || w->dua_res > w->s.infty) {
        w->status = WC_OSQP_NON_CVX;
        w->obj_val = NAN;
        return 1;
    }
    if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_prim_inf *= 10; eps_dual_inf *= 10; }
    if (w->m == 0) prim_res_check = 1;
    else {
        double eps_prim = compute_pri_tol(w, eps_abs, eps_rel);
        if (w->pri_res < eps_prim) prim_res_check = 1;
        else prim_inf_check = is_primal_infeasible(w, eps_prim_inf);
    }
    double eps_dual = compute_dua_tol(w, eps_abs, eps_rel);
    if (w->dua_res < eps_dual) dual_res_check = 1;
    else dual_inf_check = is_dual_infeasible(w, eps_dual_inf);

    if (prim_res_check && dual_res_check) {
        w->status = approximate ? WC_OSQP_SOLVED_INACCURATE : WC_OSQP_SOLVED;
        return 1;
    } else if (prim_inf_check) {
        w->status = approximate ? WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE : WC_OSQP_PRIMAL_INFEASIBLE;
        w->obj_val = w->s.infty;
        return 1;
    } else if (dual_inf_check) {
        w->status = approximate ? WC_OSQP_DUAL_INFEASIBLE_INACCURATE : WC_OSQP_DUAL_INFEASIBLE;
        w->obj_val = -w->s.infty;
        return 1;
    }
    return 0;
}
static double compute_rho_estimate(wco_osqp* w)
{
    const int n = w->n, m = w->m;
    double pri_res = vec_norm_inf(w->z_prev, m); /* scaled residual vectors left by update_info */
    double dua_res = vec_norm_inf(w->x_prev, n);
    double pn = vec_norm_inf(w->z, m), t = vec_norm_inf(w->Ax, m);
    if (t > pn) pn = t;
    pri_res /= (pn + 1e-10);
    double dn = vec_norm_inf(w->q, n);
    t = vec_norm_inf(w->Aty, n); if (t > dn) dn = t;
    t = vec_norm_inf(w->Px, n); if (t > dn) dn = t;
    dua_res /= (dn + 1e-10);
    double r = w->s.rho * sqrt(pri_res / (dua_res + 1e-10));
    r = r < RHO_MIN ? RHO_MIN : r;
    r = r > RHO_MAX ? RHO_MAX : r;
    return r;
}
static int osqp_update_rho(wco_osqp* w, double rho_new)
{
    if (rho_new <= 0) return 1;
    w->s.rho = rho_new < RHO_MIN ? RHO_MIN : (rho_new > RHO_MAX ? RHO_MAX : rho_new);
    for (int i = 0; i < w->m; i++) {
        if (w->constr_type[i] == 0) { w->rho_vec[i] = w->s.rho; w->rho_inv_vec[i] = 1.0 / w->s.rho; }
        else if (w->constr_type[i] == 1) {
            w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->s.rho;
            w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
        }
    }
    return linsys_factor(w, 0);
}
static int has_solution(const wco_osqp* w)
{
    int s = w->status;
    return s != WC_OSQP_PRIMAL_INFEASIBLE && s != WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE &&
           s != WC_OSQP_DUAL_INFEASIBLE && s != WC_OSQP_DUAL_INFEASIBLE_INACCURATE &&
           s != WC_OSQP_UNSOLVED && s != WC_OSQP_NON_CVX;
}

/* ------------------------------------------------------------------ solve (osqp osqp.c osqp_solve) */
int wco_osqp_solve(wco_osqp* w)
{
    const int n = w->n, m = w->m;
    const double alpha = w->s.alpha, sigma = w->s.sigma;
    int iter, i, can_check_termination = 0;
    double* tmp;

    if (!w->s.warm_start) {
        memset(w->x, 0, sizeof(double) * (size_t)n);
        memset(w->z, 0, sizeof(double) * (size_t)m);
        memset(w->y, 0, sizeof(double) * (size_t)m);
    }
    for (iter = 1; iter <= w->s.max_iter; iter++) {
        tmp = w->x; w->x = w->x_prev; w->x_prev = tmp;
        tmp = w->z; w->z = w->z_prev; w->z_prev = tmp;
        /* update_xz_tilde */
        for (i = 0; i < n; i++) w->xz_tilde[i] = sigma * w->x_prev[i] - w->q[i];
        for (i = 0; i < m; i++) w->xz_tilde[i + n] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
        linsys_solve(w, w->xz_tilde);
        /* update_x */
        for (i = 0; i < n; i++) {
            w->x[i] = alpha * w->xz_tilde[i] + (1.0 - alpha) * w->x_prev[i];
            w->delta_x[i] = w->x[i] - w->x_prev[i];
        }
        /* update_z */
        for (i = 0; i < m; i++) {
            double v = alpha * w->xz_tilde[i + n] + (1.0 - alpha) * w->z_prev[i] + w->rho_inv_vec[i] * w->y[i];
            v = v > w->l[i] ? v : w->l[i];
            v = v < w->u[i] ? v : w->u[i];
            w->z[i] = v;
        }
        /* update_y */
        for (i = 0; i < m; i++) {
            w->delta_y[i] = w->rho_vec[i] * (alpha * w->xz_tilde[i + n] + (1.0 - alpha) * w->z_prev[i] - w->z[i]);
            w->y[i] += w->delta_y[i];
        }
        can_check_termination = w->s.check_termination && (iter % w->s.check_termination == 0);
        if (can_check_termination) {
            update_info(w, iter);
            if (check_termination(w, 0)) break;
        }
        if (w->s.adaptive_rho && !w->adaptive_rho_interval) {
            if (w->s.check_termination)
                w->adaptive_rho_interval = ADAPTIVE_RHO_MULTIPLE_TERMINATION * w->s.check_termination;
            else
                w->adaptive_rho_interval = ADAPTIVE_RHO_FIXED;
        }
        if (w->s.adaptive_rho && w->adaptive_rho_interval && (iter % w->adaptive_rho_interval == 0)) {
            if (!can_check_termination) update_info(w, iter);
            double rho_new = compute_rho_estimate(w);
            w->rho_estimate = rho_new;
            if (rho_new > w->s.rho * w->s.adaptive_rho_tolerance ||
                rho_new < w->s.rho / w->s.adaptive_rho_tolerance) {
                osqp_update_rho(w, rho_new);
                w->rho_updates += 1;
            }
        }
    }
    if (!can_check_termination) {
        update_info(w, iter - 1);
        check_termination(w, 0);
    }
    if (w->status == WC_OSQP_UNSOLVED) {
        if (!check_termination(w, 1)) w->status = WC_OSQP_MAX_ITER_REACHED;
    }
    w->rho_estimate = compute_rho_estimate(w);
    /* store_solution */
    if (has_solution(w)) {
        for (i = 0; i < n; i++) w->sol_x[i] = w->x[i];
        for (i = 0; i < m; i++) w->sol_y[i] = w->y[i];
        if (w->s.scaling) {
            for (i = 0; i < n; i++) w->sol_x[i] *= w->D[i];
            for (i = 0; i < m; i++) w->sol_y[i] *= w->E[i];
            for (i = 0; i < m; i++) w->sol_y[i] *= w->cinv;
        }
    } else {
        for (i = 0; i < n; i++) w->sol_x[i] = NAN;
        for (i = 0; i < m; i++) w->sol_y[i] = NAN;
        memset(w->x, 0, sizeof(double) * (size_t)n);
        memset(w->z, 0, sizeof(double) * (size_t)m);
        memset(w->y, 0, sizeof(double) * (size_t)m);
    }
    return w->status;
}

int wco_osqp_status(const wco_osqp* w) { return w->status; }
int wco_osqp_iter(const wco_osqp* w) { return w->iter; }
int wco_osqp_rho_updates(const wco_osqp* w) { return w->rho_updates; }
double wco_osqp_rho(const wco_osqp* w) { return w->s.rho; }
double wco_osqp_pri_res(const wco_osqp* w) { return w->pri_res; }
double wco_osqp_dua_res(const wco_osqp* w) { return w->dua_res; }
const double* wco_osqp_solution_x(const wco_osqp* w) { return w->sol_x; }
const double* wco_osqp_solution_y(const wco_osqp* w) { return w->sol_y; }
int wco_osqp_n(const wco_osqp* w) { return w->n; }
int wco_osqp_m(const wco_osqp* w) { return w->m; }
