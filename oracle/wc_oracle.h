/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Public interface of the CPU restatement of the reference hot path (see oracle/README.md).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load liboracle; the product library (walking-controllers_b200/csrc) never does.
 *
 * PARITY UNPINNED: the reference repository contains no tests, golden vectors or fixtures
 * for this path, and none of its third-party solvers (OSQP, qpOASES, iDynTree) can be built
 * or imported in this environment.  The oracle is validated by independent means instead
 * (tests/test_oracle_*.py: exact KKT solves, scipy, finite differences, closed forms).
 */
#ifndef WC_ORACLE_H
#define WC_ORACLE_H

#include "../include/wc_capi.h"

#ifdef __cplusplus
extern "C" {
#endif

#define WCO_MAX_HULL 8

/* ---- 2-D convex hull helpers (iDynTree ConvexHullProjectionConstraint restatement) ---- */
int wco_hull_build(const double* pts, int npts, double* A, double* b, double* verts);
double wco_hull_margin(const double* A, const double* b, const double* verts, int nh, const double* z);

/* ---- single-instance DCM-MPC (WalkingController + MPCSolver + OSQP) ---- */
typedef struct wco_mpc wco_mpc;
wco_mpc* wco_mpc_create(const wc_mpc_config* cfg);
void wco_mpc_free(wco_mpc* c);
void wco_mpc_reset(wco_mpc* c);
void wco_mpc_set_weights(wco_mpc* c, const double q_diag[2], const double r_diag[2]);
void wco_mpc_set_linsys(wco_mpc* c, int kind);
/* lf_H, rf_H: row-major 3x4; ref: (N+1) samples x 2, sample-major; returns WC_MPC_FLAG_* mask */
int wco_mpc_tick(wco_mpc* c, const double* lf_H, const double* rf_H, int l_contact, int r_contact,
                 const double* dcm, const double* ref, int reset_traj,
                 double* zmp, int* osqp_status, int* iters);
int wco_mpc_n_hull(const wco_mpc* c);
void wco_mpc_get_hull(const wco_mpc* c, double* A, double* b);
int wco_mpc_get_solution(const wco_mpc* c, double* primal);
double wco_mpc_rho(const wco_mpc* c);
int wco_mpc_rho_updates(const wco_mpc* c);
int wco_mpc_get_qp_dense(const wco_mpc* c, double* Pd, double* Ad, double* q, double* l, double* u);

/* ---- batched driver, SoA layout identical to wc_mpc_tick_host (one instance per OpenMP thread) ---- */
typedef struct wco_mpc_batch wco_mpc_batch;
wco_mpc_batch* wco_mpc_batch_create(const wc_mpc_config* cfg, int B);
void wco_mpc_batch_free(wco_mpc_batch* b);
wco_mpc* wco_mpc_batch_get(wco_mpc_batch* b, int i);
void wco_mpc_batch_set_weights(wco_mpc_batch* b, const double* q_diag, const double* r_diag);
void wco_mpc_batch_reset(wco_mpc_batch* b, const unsigned char* mask);
void wco_mpc_batch_tick(wco_mpc_batch* b, const double* lf_H, const double* rf_H,
                        const unsigned char* l_contact, const unsigned char* r_contact,
                        const double* dcm, const double* ref, const unsigned char* reset,
                        double* zmp, int* osqp_status, int* iters, int* flags, int n_threads);
void wco_mpc_batch_get_solution(wco_mpc_batch* b, double* primal, double* rho, int* rho_updates);

/* ---- generic OSQP restatement on dense input (for cross-checks from Python) ---- */
/* P: n x n row-major symmetric (upper triangle is used), A: m x n row-major.
 * Returns status; x_out [n], y_out [m], info_out = {iter, rho_updates, rho, pri_res, dua_res}. */
int wco_osqp_solve_dense(int n, int m, const double* P, const double* A, const double* q,
                         const double* l, const double* u, const wc_osqp_settings* settings,
                         int linsys_kind, double* x_out, double* y_out, double* info_out);

#ifdef __cplusplus
}
#endif
#endif
