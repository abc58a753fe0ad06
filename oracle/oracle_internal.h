/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * Internal declarations shared by the oracle translation units.
 */
#ifndef WCO_ORACLE_INTERNAL_H
#define WCO_ORACLE_INTERNAL_H

#include "../include/wc_capi.h" /* config structs only (single source of truth for layouts) */

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- sparse LDL^T (QDLDL restatement) ---------------- */
typedef struct wco_ldl {
    int n;
    int *etree, *Lnz, *Lp, *Li;
    double *Lx, *D, *Dinv;
    int* iwork;
    unsigned char* bwork;
    double* fwork;
} wco_ldl;

wco_ldl* wco_ldl_symbolic(int n, const int* Ap, const int* Ai);
int wco_ldl_numeric(wco_ldl* f, const int* Ap, const int* Ai, const double* Ax);
void wco_ldl_solve(const wco_ldl* f, double* x);
void wco_ldl_free(wco_ldl* f);

/* ---------------- CSC helper ---------------- */
typedef struct wco_csc {
    int nrow, ncol, nnz;
    int *p, *i;
    double* x;
} wco_csc;

wco_csc* wco_csc_from_triplets(int nrow, int ncol, int nt, const int* ti, const int* tj, const double* tx);
wco_csc* wco_csc_copy(const wco_csc* a);
void wco_csc_free(wco_csc* a);

/* ---------------- OSQP restatement ---------------- */
typedef struct wco_osqp wco_osqp;

/* P: upper triangle CSC (n x n); A: CSC (m x n); perm: optional KKT ordering of size n+m
 * (perm[k] = index of the KKT variable placed k-th; NULL = identity). */
wco_osqp* wco_osqp_setup(const wco_csc* P, const wco_csc* A, const double* q, const double* l,
                         const double* u, const wc_osqp_settings* settings, const int* perm);
void wco_osqp_free(wco_osqp* w);
void wco_osqp_update_lin_cost(wco_osqp* w, const double* q_new);
int wco_osqp_update_bounds(wco_osqp* w, const double* l_new, const double* u_new);
int wco_osqp_solve(wco_osqp* w); /* returns the OSQP status value */
int wco_osqp_status(const wco_osqp* w);
int wco_osqp_iter(const wco_osqp* w);
int wco_osqp_rho_updates(const wco_osqp* w);
double wco_osqp_rho(const wco_osqp* w);
double wco_osqp_pri_res(const wco_osqp* w);
double wco_osqp_dua_res(const wco_osqp* w);
const double* wco_osqp_solution_x(const wco_osqp* w); /* unscaled primal (NaN if no solution) */
const double* wco_osqp_solution_y(const wco_osqp* w);
int wco_osqp_n(const wco_osqp* w);
int wco_osqp_m(const wco_osqp* w);
/* linear-system back end: 0 = KKT LDL^T (QDLDL, what OSQP does), 1 = dense Cholesky of the
 * reduced system P+sigma I+A' rho A (used to quantify the effect of the solve method). */
void wco_osqp_set_linsys(wco_osqp* w, int kind);
/* osqp_update_P / osqp_update_A with all values of the (unchanged) pattern, in CSC order */
int wco_osqp_update_P(wco_osqp* w, const double* Px_new);
int wco_osqp_update_A(wco_osqp* w, const double* Ax_new);
void wco_osqp_get_iterate(const wco_osqp* w, double* x_scaled, double* y_scaled);
void wco_osqp_warm_start(wco_osqp* w, const double* x, const double* y);

#ifdef __cplusplus
}
#endif
#endif
