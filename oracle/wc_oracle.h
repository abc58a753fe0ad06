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
void wco_mpc_set_hull_variant(wco_mpc* c, int variant); /* sensitivity studies: row order / scaling of the hull rows */
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

/* ---- kinematics (iDynTree KinDynComputations restatement), single instance, row-major ---- */
void wco_fk(const wc_model* m, const double* q, const double* base_H12, int left_is_fixed,
            double* lf_H12, double* rf_H12, double* neck_R9, double* com3,
            double* J_lf, double* J_rf, double* J_neck, double* J_com);
void wco_fk_batch(const wc_model* model, int B, const double* q, const double* base_H, const unsigned char* left_is_fixed,
                  double* lf_H, double* rf_H, double* neck_R, double* com,
                  double* J_lf, double* J_rf, double* J_neck, double* J_com);

/* ---- exact dense QP (qpOASES stand-in) and WalkingQPIK_qpOASES restatement ---- */
int wco_qp_solve(int n, const double* H, const double* g, int mE, const double* A, const double* b,
                 const double* lb, const double* ub, int nwsr_max, double* x, int* nwsr_out);
int wco_ik_solve(const wc_ik_config* cfg,
                 const double* q_meas, const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                 const double* J_lf, const double* J_rf, const double* J_neck6, const double* J_com,
                 const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                 const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                 double* dq, double* lf_err, double* rf_err, int* nwsr,
                 double* H_out, double* g_out, double* A_out, double* b_out, double* v_out);
void wco_ik_batch_tick(const wc_ik_config* cfg, const wc_model* model, int B,
                       const double* q_meas, const double* q_lin, const double* base_H, const unsigned char* left_is_fixed,
                       const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                       const double* com_pos_des, const double* com_vel_des, const double* com_meas, const double* neck_R_des,
                       const double* kpos, const double* katt,
                       double* dq, double* lf_err, double* rf_err, int* status, int* nwsr, int n_threads);

/* ---- generic OSQP restatement on dense input (for cross-checks from Python) ---- */
/* P: n x n row-major symmetric (upper triangle is used), A: m x n row-major.
 * Returns status; x_out [n], y_out [m], info_out = {iter, rho_updates, rho, pri_res, dua_res}. */
int wco_osqp_solve_dense(int n, int m, const double* P, const double* A, const double* q,
                         const double* l, const double* u, const wc_osqp_settings* settings,
                         int linsys_kind, double* x_out, double* y_out, double* info_out);

/* ---- closed-loop glue (StableDCMModel, WalkingZMPController, joint integrator); batch SoA ---- */
typedef struct wco_glue wco_glue;
wco_glue* wco_glue_create(const wc_glue_config* cfg, int B);
void wco_glue_free(wco_glue* g);
void wco_glue_reset(wco_glue* g, const double* com_xy, const double* q);
void wco_glue_set_phase(wco_glue* g, const double* dcm_vel_des, const unsigned char* stance);
void wco_glue_com_reference(wco_glue* g, const double* dcm_des, const double* zmp_des, const double* zmp_meas,
                            const double* com_meas, const double* com_z_des, const double* com_z_vel_des,
                            const double* k_zmp, const double* k_com, double* com_pos_des, double* com_vel_des);
void wco_glue_integrate_joints(wco_glue* g, const double* dq, double* q_des);
/* rank 4: WalkingDCMReactiveController, WalkingFK::evaluateDCM, WalkingModule::evaluateZMP, the logger row */
void wco_reactive_dcm(int B, double omega, double k_const, const double* k_dcm, const double* dcm_des,
                      const double* dcm_vel_des, const double* dcm_meas, double* zmp_des);
void wco_evaluate_dcm(int B, double omega, const double* com, const double* com_vel, double* dcm);
void wco_evaluate_zmp(int B, const double* lw, const double* rw, const double* lf_H, const double* rf_H, double* zmp, int* ok);
void wco_log_pack(int B, const wc_log_inputs* in, double* out);
void wco_glue_reset_com_filter(wco_glue* g);
void wco_glue_evaluate_com(wco_glue* g, const double* com, const double* J_com, const double* dq, double cut_frequency,
                           int use_filters, double* com_out, double* vel_out, double* dcm_out);

/* ---- rank 2: WalkingQPIK_osqp (qpik_osqp_ref.c) ---- */
void wco_ik_build_qp(const wc_ik_config* cfg, int variant,
                     const double* q_meas, const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                     const double* J_lf, const double* J_rf, const double* J_neck6, const double* J_com,
                     const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                     const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                     double* H, double* g, double* A, double* b);
typedef struct wco_ikosqp wco_ikosqp;
wco_ikosqp* wco_ikosqp_create(const wc_ik_config* cfg, const wc_osqp_settings* st);
void wco_ikosqp_free(wco_ikosqp* h);
int wco_ikosqp_setups(const wco_ikosqp* h);
int wco_ikosqp_solve(wco_ikosqp* h,
                     const double* q_meas, const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                     const double* J_lf, const double* J_rf, const double* J_neck6, const double* J_com,
                     const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                     const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                     double* dq, double* lf_err, double* rf_err, int* osqp_status, int* iters, int* flags);
typedef struct wco_ikosqp_batch wco_ikosqp_batch;
wco_ikosqp_batch* wco_ikosqp_batch_create(const wc_ik_config* cfg, const wc_osqp_settings* st, const wc_model* model, int B);
void wco_ikosqp_batch_free(wco_ikosqp_batch* b);
void wco_ikosqp_batch_tick(wco_ikosqp_batch* b, const double* q_meas, const double* q_lin, const double* base_H,
                           const unsigned char* left_is_fixed, const double* lf_H_des, const double* rf_H_des,
                           const double* lf_twist, const double* rf_twist, const double* com_pos_des, const double* com_vel_des,
                           const double* com_meas, const double* neck_R_des,
                           double* dq, double* lf_err, double* rf_err, int* status, int* iters, int* flags, int* setups);

/* ---- rank 3: synthetic producer of the path's inputs (CPU twin of wc_traj_*) ---- */
typedef struct wco_traj wco_traj;
wco_traj* wco_traj_create(const wc_traj_config* cfg, int B);
void wco_traj_free(wco_traj* g);
void wco_traj_set_gait(wco_traj* g, const int32_t* phase, const double* step_len, const double* dtheta, const double* theta0,
                       const uint8_t* first_left);
void wco_traj_dcm(const wco_traj* g, int first_sample, int n_samples, double* pos_out, double* vel_out);
void wco_traj_measured_dcm(const wco_traj* g, int tick, double sigma, uint64_t seed, double* out);
void wco_traj_feet(const wco_traj* g, int tick, const wc_traj_feet* out);
void wco_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1);

#ifdef __cplusplus
}
#endif
#endif
