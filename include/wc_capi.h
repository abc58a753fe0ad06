/*
 * wc_capi.h — C ABI of the B200-native batched walking-controllers hot path.
 *
 * The reference (lia2790/walking-controllers) has no FFI layer: its boundary for
 * this path is two C++ class interfaces called once per 10 ms control tick from
 * WalkingModule::updateModule (modules/Walking_module/src/WalkingModule.cpp:601-637
 * for the MPC block, :709-745 and :367-425 for the QP-IK block).  This header is the
 * C-ABI a replacement of that path binds: plain pointers and sizes, no C++/torch
 * types.  The drop-in C++ classes in walking-controllers_b200/host/ are thin views
 * over these entry points.
 *
 * Conventions
 *  - every entry point returns wc_status (0 = ok, <0 = error) and never throws;
 *  - B = number of independent controller instances owned by the handle;
 *  - all batched arrays are fp64 structure-of-arrays: element k of instance i is at
 *    ptr[k*B + i] ("field-major"), so that consecutive instances are consecutive
 *    in memory and warp loads coalesce;
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *    asynchronous on that stream; one stream per handle; handles are not thread-safe;
 *  - *_dev entry points take DEVICE pointers; *_host entry points take HOST pointers
 *    and stage through pinned memory inside the call (these are what the B=1 drop-in
 *    classes and the end-to-end benchmark use);
 *  - per-instance failures never poison the batch: they are reported through the
 *    int32 status arrays (OSQP / qpOASES-style codes below), the call still returns 0.
 */
#ifndef WC_CAPI_H
#define WC_CAPI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t wc_status;

enum {
    WC_OK = 0,
    WC_ERR_INVALID_ARG = -1,
    WC_ERR_CUDA = -2,
    WC_ERR_UNSUPPORTED = -3, /* e.g. non-diagonal Q/R on the fast MPC path */
    WC_ERR_NO_DEVICE = -4,
    WC_ERR_ALLOC = -5,
    WC_ERR_STATE = -6 /* call order violated (e.g. get_output before solve) */
};

/* ---- OSQP status codes (osqp 0.6 constants.h), returned per instance by the MPC ---- */
enum {
    WC_OSQP_DUAL_INFEASIBLE_INACCURATE = 4,
    WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3,
    WC_OSQP_SOLVED_INACCURATE = 2,
    WC_OSQP_SOLVED = 1,
    WC_OSQP_MAX_ITER_REACHED = -2,
    WC_OSQP_PRIMAL_INFEASIBLE = -3,
    WC_OSQP_DUAL_INFEASIBLE = -4,
    WC_OSQP_NON_CVX = -7,
    WC_OSQP_UNSOLVED = -10
};

/* Per-instance controller-level flags of one MPC tick (bit mask, `flags` output).
 * The reference collapses all of these into the bool returned by
 * WalkingController::{setConvexHullConstraint,solve}
 * (WalkingDCMModelPredictiveController.cpp:364-435,491-521). */
enum {
    WC_MPC_FLAG_OK = 0,
    WC_MPC_FLAG_NO_CONTACT = 1,   /* (false,false) contact pair: "None foot is in contact" (:404-408); raised on EVERY
                                     tick without contact (the reference stops at the first), nothing is solved */
    WC_MPC_FLAG_NO_SOLVER = 2,    /* no MPCSolver exists yet (reference would dereference null) */
    WC_MPC_FLAG_NOT_SOLVED = 4,   /* OsqpEigen::Solver::solve() false: a non-convex problem, or -- with
                                     osqp_eigen_checks_status -- any status but OSQP_SOLVED; the output keeps m_output */
    WC_MPC_FLAG_OUTSIDE_HULL = 8, /* computeMargin(u0) < -convex_hull_tolerance (:513-517) */
    WC_MPC_FLAG_NEW_SOLVER = 16   /* informational: support phase changed, cold OSQP setup ran */
};

/* OSQP settings; defaults = osqp 0.6 defaults, which the reference uses untouched
 * except verbose=0 (MPCSolver.cpp:51).  adaptive_rho_interval: the reference leaves 0
 * (= chosen from wall-clock timings inside OSQP, non-deterministic); this
 * implementation and its oracle use OSQP's own non-profiling fallback
 * 4*check_termination = 100 when 0 is given. */
typedef struct wc_osqp_settings {
    double rho;                    /* 0.1 */
    double sigma;                  /* 1e-6 */
    double alpha;                  /* 1.6 */
    double eps_abs;                /* 1e-3 */
    double eps_rel;                /* 1e-3 */
    double eps_prim_inf;           /* 1e-4 */
    double eps_dual_inf;           /* 1e-4 */
    double adaptive_rho_tolerance; /* 5 */
    double infty;                  /* OSQP_INFTY = 1e30 */
    int32_t max_iter;              /* 4000 */
    int32_t scaling;               /* 10 Ruiz passes */
    int32_t check_termination;     /* 25 */
    int32_t adaptive_rho;          /* 1 */
    int32_t adaptive_rho_interval; /* 0 -> 4*check_termination */
    int32_t warm_start;            /* 1 */
    int32_t osqp_eigen_checks_status; /* what OsqpEigen::Solver::solve() returns (MPCSolver.cpp:309, ..._osqp.cpp:377):
                                       0 (default) = osqp-eigen as of the reference (2018): true whenever osqp_solve()
                                           ran, whatever the status -- the caller reads OSQP's stored solution (the last
                                           iterate for SOLVED_INACCURATE / MAX_ITER_REACHED, NaN for the infeasible statuses);
                                       1 = osqp-eigen >= 0.4: false unless the status is OSQP_SOLVED.  DESIGN.md section 9 */
    int32_t reserved1;
} wc_osqp_settings;

/* Configuration of the DCM-MPC (WalkingController::initialize,
 * WalkingDCMModelPredictiveController.cpp:170-362; values from
 * app/robots/<robot>/controllerParams.ini + the GENERAL group). */
typedef struct wc_mpc_config {
    int32_t horizon;         /* N = round(controllerHorizon / sampling_time)  (:185-187) */
    int32_t reserved;
    double sampling_time;    /* dT   (:182, default 0.016) */
    double com_height;       /* (:225) */
    double gravity;          /* (:230, default 9.81) */
    double Q[4];             /* state weight, row-major 2x2, AS SEEN BY THE SOLVER
                                (i.e. after getTripletsFromValues' (col,row) swap, Utils.cpp:93-101) */
    double R[4];             /* input weight, same convention */
    double foot_x[2];        /* foot_size x limits as written in the ini (min max) (:262-263) */
    double foot_y[2];        /* foot_size y limits (:280-281) */
    double initial_zmp[2];   /* initial_zmp_position (:317-344) */
    double hull_tolerance;   /* convex_hull_tolerance (:306, default 0.01) */
    wc_osqp_settings osqp;
} wc_mpc_config;

void wc_osqp_default_settings(wc_osqp_settings* s);
/* iCubGazeboV2_5 shipped parameters with the given horizon (25 = BASELINE configs, 200 = shipped). */
void wc_mpc_default_config(wc_mpc_config* cfg, int32_t horizon);

/* ---------------------------------------------------------------------------------- */
/* Batched DCM-MPC:  replaces WalkingController + MPCSolver + OsqpEigen/OSQP            */
/* (WalkingDCMModelPredictiveController.hpp:190-250, MPCSolver.hpp:27-139).             */
/* ---------------------------------------------------------------------------------- */
typedef struct wc_mpc wc_mpc;

/* WalkingController::initialize (:311-362) for B instances on CUDA device `device`.
 * Optional per-instance weight overrides (gain sweeps, BASELINE config 4) via
 * wc_mpc_set_weights_*. */
wc_status wc_mpc_create(const wc_mpc_config* cfg, int32_t B, int32_t device, wc_mpc** out);
void wc_mpc_destroy(wc_mpc* h);
int32_t wc_mpc_batch(const wc_mpc* h);
int32_t wc_mpc_horizon(const wc_mpc* h);

/* Per-instance diagonal weights q_diag[2*B], r_diag[2*B] (field-major), host pointers.
 * Must be called before the first tick or right after wc_mpc_reset. */
wc_status wc_mpc_set_weights_host(wc_mpc* h, const double* q_diag, const double* r_diag);

/* WalkingController::reset (:537-543): forget the feet status for the masked instances
 * (mask == NULL: all), so that the next tick rebuilds hull + solver. mask is a HOST array. */
wc_status wc_mpc_reset(wc_mpc* h, const uint8_t* mask_host);

/* One control tick of the MPC block (WalkingModule.cpp:601-637), fused:
 *   setConvexHullConstraint(left.front(), right.front(), leftInContact.front(), rightInContact.front())
 *   setFeedback(dcm_meas) ; setReferenceSignal(ref, reset) ; solve() ; getControllerOutput(zmp)
 * Inputs (field-major SoA, B instances):
 *   lf_H, rf_H   [12*B]  world_H_foot, row-major 3x4 ([R|p]) per instance: element e at [e*B+i]
 *   l_contact, r_contact [B] uint8
 *   dcm_meas     [2*B]
 *   ref          [(N+1)*2*B]  DCM reference window: sample k, axis a at [(2*k+a)*B + i]
 *                (a shorter deque is padded with its last sample by the caller, MPCSolver.cpp:201-215)
 *   reset        [B] uint8 or NULL (= all 0)  resetTrajectory flag
 * Outputs:
 *   zmp [2*B] desired ZMP u0 (m_output: the last accepted output when the solve fails);
 *   osqp_status [B], iters [B], flags [B] (WC_MPC_FLAG_* mask); any may be NULL. */
wc_status wc_mpc_tick_dev(wc_mpc* h,
                          const double* lf_H, const double* rf_H,
                          const uint8_t* l_contact, const uint8_t* r_contact,
                          const double* dcm_meas, const double* ref, const uint8_t* reset,
                          double* zmp, int32_t* osqp_status, int32_t* iters, int32_t* flags,
                          void* stream);
/* The same tick with the DCM reference window kept in a device-side ring (what MPCSolver::setGradient's block shift does
 * with the gradient, MPCSolver.cpp:216-239): per tick the caller provides only the NEWEST sample of the window,
 * ref_newest [2*B] = sample N of this tick's window.  ref_window [(N+1)*2*B] (layout of `ref` above) replaces the whole
 * ring instead; it is required at the first ring tick and whenever the trajectory was re-planned (the ticks on which
 * the reference passes resetTrajectory = true) and may be NULL otherwise.  Results are bit-identical to
 * wc_mpc_tick_dev on the same windows (tests/test_host_pipeline_gpu.py). */
wc_status wc_mpc_tick_ring_dev(wc_mpc* h, const double* lf_H, const double* rf_H,
                               const uint8_t* l_contact, const uint8_t* r_contact,
                               const double* dcm_meas, const double* ref_newest, const double* ref_window,
                               const uint8_t* reset, double* zmp, int32_t* osqp_status, int32_t* iters, int32_t* flags,
                               void* stream);

/* Same with HOST pointers (pinned staging + H2D/D2H inside the call, synchronous on return). */
wc_status wc_mpc_tick_host(wc_mpc* h,
                           const double* lf_H, const double* rf_H,
                           const uint8_t* l_contact, const uint8_t* r_contact,
                           const double* dcm_meas, const double* ref, const uint8_t* reset,
                           double* zmp, int32_t* osqp_status, int32_t* iters, int32_t* flags);

/* Split form of wc_mpc_tick_host: _async enqueues the H2D copies, the kernels and the D2H copies on the
 * handle's own upload / compute / download streams and returns.  Staging has three slots: up to THREE ticks may
 * be in flight (a fourth _async call first retires the oldest), so the upload of tick t+1 and the download of
 * tick t-1 overlap the kernels of tick t; another handle's (e.g. the IK's) traffic overlaps as well.
 * wc_mpc_sync_one retires the oldest tick in flight (its outputs are then valid), wc_mpc_sync all of them.
 * Input and output buffers of a tick must stay valid (and should be pinned: cudaHostAlloc / torch pin_memory)
 * until that tick is retired; ticks in flight need distinct output buffers. */
wc_status wc_mpc_tick_host_async(wc_mpc* h,
                           const double* lf_H, const double* rf_H,
                           const uint8_t* l_contact, const uint8_t* r_contact,
                           const double* dcm_meas, const double* ref, const uint8_t* reset,
                           double* zmp, int32_t* osqp_status, int32_t* iters, int32_t* flags);
wc_status wc_mpc_sync(wc_mpc* h);
wc_status wc_mpc_sync_one(wc_mpc* h);
/* NOTE (all handles): wc_*_reset*, wc_*_set_weights_host / wc_*_set_gains_host first retire the host ticks in flight
 * (they touch state those ticks read or write); nothing else orders them against *_dev ticks on caller streams --
 * synchronise those streams before calling them. */

/* Debug / parity taps (device -> host copies, synchronous): full primal solution
 * z = [x_0..x_N, u_0..u_{N-1}] per instance ([n*B], field-major, n = 4N+2), the current
 * rho and number of rho updates of the instance's solver.  Any pointer may be NULL. */
wc_status wc_mpc_debug_get_host(wc_mpc* h, double* primal, double* rho, int32_t* rho_updates);

/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t wc_mpc_launch_count(const wc_mpc* h);

/* Per-kernel device timing: when enabled, every wc_mpc_tick_* brackets K1 (problem build) and K2
 * (ADMM) with CUDA events on the launch stream.  wc_mpc_kernel_times_ms synchronises, returns the
 * accumulated milliseconds and tick count since the previous call and clears them. */
wc_status wc_mpc_enable_profiling(wc_mpc* h, int32_t on);
wc_status wc_mpc_kernel_times_ms(wc_mpc* h, double* prepare_ms, double* admm_ms, int64_t* ticks);
/* Split of admm_ms of the window closed by the last wc_mpc_kernel_times_ms call: OSQP setup of new
 * workspaces / ADMM up to the first termination check (all instances) / the compacted remainder.
 * On the cooperative path (long horizons) everything is reported as phase 1. */
wc_status wc_mpc_kernel_times_detail_ms(wc_mpc* h, double* setup_ms, double* phase1_ms, double* phase2_ms);
/* ADMM kernel family of the handle: 1 = chain-per-thread (mpc_chain.cu, short horizons), 0 = cooperative. */
int32_t wc_mpc_path(const wc_mpc* h);
/* Optional device-side tick source: with tick_dev != NULL the slot of stage 0 in the gradient / reference rings is
 * (*tick_dev) mod (N+1), read on the device, instead of the handle's own host-side tick count (NULL switches back).
 * The caller must pass reset = 1 for every instance on the first tick after a switch (the rings are re-filled). */
wc_status wc_mpc_set_tick_source_dev(wc_mpc* h, const int32_t* tick_dev);

/* ---------------------------------------------------------------------------------- */
/* Kinematic model table (replaces the iDynTree::Model loaded from model.urdf,          */
/* WalkingModule.cpp:104-119; the URDF is not part of the reference repository).        */
/* ---------------------------------------------------------------------------------- */
#define WC_MAX_LINKS 40
#define WC_MAX_DOFS 32

typedef struct wc_model {
    int32_t n_links;                    /* link 0 is the URDF root (root_link) */
    int32_t n_dofs;                     /* actuated joints, reference order (robotControl.ini:3-7) */
    int32_t parent[WC_MAX_LINKS];       /* parent link (-1 for link 0); parent[i] < i */
    int32_t dof[WC_MAX_LINKS];          /* dof index of the joint parent->link, -1 = fixed joint */
    double  joint_R[WC_MAX_LINKS][9];   /* parent_R_jointframe, row-major */
    double  joint_p[WC_MAX_LINKS][3];   /* origin of the joint frame in the parent link frame */
    double  axis[WC_MAX_LINKS][3];      /* revolute axis in the joint (= child link) frame, unit */
    double  mass[WC_MAX_LINKS];
    double  com[WC_MAX_LINKS][3];       /* link CoM in link frame */
    int32_t l_sole_link, r_sole_link, neck_link; /* links the three task frames are attached to */
    int32_t reserved;
    double  l_sole_R[9], l_sole_p[3];   /* link_H_frame of l_sole */
    double  r_sole_R[9], r_sole_p[3];
    double  neck_R[9], neck_p[3];
} wc_model;

/* Built-in iCub-like 23-DoF model (synthetic geometry/inertia; joint order =
 * app/robots/iCubGazeboV2_5/robotControl.ini:3-7). */
void wc_model_icub_like(wc_model* m);

/* Configuration of the QP-IK (WalkingQPIK::initializeMatrices,
 * WalkingQPInverseKinematics.cpp:25-116; WalkingQPIK_qpOASES::initialize,
 * WalkingQPInverseKinematics_qpOASES.cpp:53-133; values from qpInverseKinematics.ini). */
typedef struct wc_ik_config {
    int32_t n_dofs;                     /* actuatedDOFs: 12 .. 26 (23 in every shipped robot) */
    int32_t use_com_as_constraint;      /* useCoMAsConstraint (:66) */
    double com_weight[9];               /* comWeightTriplets as seen by the solver (only if !constraint) */
    double neck_weight[9];              /* neckWeightTriplets as seen by the solver, row-major 3x3 */
    double additional_rotation[9];      /* additional_rotation, row-major (:118) */
    double joint_reg_rad[WC_MAX_DOFS];  /* jointRegularization, already deg->rad (:94-103) */
    double joint_reg_weight[WC_MAX_DOFS]; /* jointRegularizationWeights */
    double joint_reg_gain[WC_MAX_DOFS]; /* jointRegularizationGains */
    double k_pos_foot, k_att_foot, k_neck, k_pos_com;
    double min_vel[WC_MAX_DOFS];        /* minJointsLimit passed to initialize (rad/s) */
    double max_vel[WC_MAX_DOFS];
    int32_t nwsr;                       /* 100 (:312) */
    int32_t reserved;
} wc_ik_config;

void wc_ik_default_config(wc_ik_config* cfg); /* iCubGazeboV2_5 values, v_max = 1 rad/s placeholder */

/* qpOASES-style return codes of the batched QP-IK solve (names follow qpOASES
 * MessageHandling.hpp; numeric values are ours - the reference collapses them to bool,
 * WalkingQPInverseKinematics_qpOASES.cpp:316-335). */
enum {
    WC_QP_SUCCESSFUL_RETURN = 0,
    WC_QP_MAX_NWSR_REACHED = 1,
    WC_QP_INFEASIBLE = 2,     /* RET_INIT_FAILED_INFEASIBILITY / RET_HOTSTART_STOPPED_INFEASIBILITY */
    WC_QP_NOT_CONVEX = 3      /* projected Hessian not positive definite (RET_*_UNBOUNDEDNESS) */
};

/* ---------------------------------------------------------------------------------- */
/* Batched FK + QP-IK: replaces WalkingFK (KinDynComputations) + WalkingQPIK_qpOASES    */
/* (WalkingForwardKinematics.hpp:73-220, WalkingQPInverseKinematics.hpp:25-209).        */
/* ---------------------------------------------------------------------------------- */
typedef struct wc_ik wc_ik;

wc_status wc_ik_create(const wc_ik_config* cfg, const wc_model* model, int32_t B, int32_t device, wc_ik** out);
void wc_ik_destroy(wc_ik* h);
int32_t wc_ik_batch(const wc_ik* h);
int32_t wc_ik_dofs(const wc_ik* h);

/* Per-instance gain overrides k_pos_foot[B], k_att_foot[B] (gain sweeps). Host pointers. */
wc_status wc_ik_set_gains_host(wc_ik* h, const double* k_pos_foot, const double* k_att_foot);

/* One control tick of the FK + QP-IK block (WalkingModule.cpp:715-744 + solveQPIK :367-425):
 *   WalkingFK::evaluateWorldToBaseTransformation / setInternalRobotState(q_lin, .)
 *   setRobotState(q_meas, FK(l_sole), FK(r_sole), FK(neck), com_meas)
 *   setDesiredNeckOrientation(neck_R_des) ; setDesiredFeetTransformation ; setDesiredFeetTwist
 *   setDesiredCoMVelocity ; setDesiredCoMPosition ; set*Jacobian(FK) ; solve() ; getSolution(dq)
 * Inputs (field-major SoA):
 *   q_meas [nJ*B]  measured joint positions (regularisation term, quirk 9 of SURVEY §8)
 *   q_lin  [nJ*B]  joint positions the kinematics is linearised about (integrated desired)
 *   base_H [12*B]  world_H_stancefoot-SOLE frame (desired foot transform on the stance side)
 *   left_is_fixed [B] uint8  stance flag (base link = l_foot / r_foot link)
 *   lf_H_des, rf_H_des [12*B], lf_twist, rf_twist [6*B] (mixed: linear then angular),
 *   com_pos_des, com_vel_des, com_meas [3*B], neck_R_des [9*B] (the matrix passed to
 *   setDesiredNeckOrientation, BEFORE right-multiplication by additional_rotation).
 * Outputs: dq [nJ*B]; lf_err, rf_err [6*B] (getLeftFootError/getRightFootError);
 *          qp_status [B] (WC_QP_*), nwsr [B] working-set changes.  May be NULL. */
wc_status wc_ik_tick_dev(wc_ik* h,
                         const double* q_meas, const double* q_lin,
                         const double* base_H, const uint8_t* left_is_fixed,
                         const double* lf_H_des, const double* rf_H_des,
                         const double* lf_twist, const double* rf_twist,
                         const double* com_pos_des, const double* com_vel_des,
                         const double* com_meas, const double* neck_R_des,
                         double* dq, double* lf_err, double* rf_err,
                         int32_t* qp_status, int32_t* nwsr, void* stream);
/* The same tick as TWO calls, for a caller that wants the kinematics out of the way before the desired CoM exists.  In
 * WalkingModule::updateModule the kinematics of the QP-IK is evaluated at the integrated desired posture
 * (setInternalRobotState(m_qDesired, ...), WalkingModule.cpp:715) and the desired feet / neck come from the trajectory
 * generator: none of it depends on the MPC output of the tick, only setDesiredCoMPosition / setDesiredCoMVelocity do
 * (WalkingModule.cpp:660-740).  So
 *   WC_IK_PART_KINEMATICS  forward kinematics, Jacobian columns of the QP and every task right-hand side except the CoM
 *                          term (reads q_meas, q_lin, base_H, left_is_fixed, the desired feet / twists / neck; com_* may
 *                          be NULL); com_out [3*B] or NULL receives the CoM of that posture (what wc_fk_dev returns);
 *   WC_IK_PART_SOLVE       the CoM term, the QP solves, the outputs (same arguments as wc_ik_tick_dev; the pointers of
 *                          the kinematics call must still hold the same data)
 * issued in this order (any streams, ordered by the caller) give bit for bit the result of wc_ik_tick_dev: the kinematics
 * call can run on another stream next to the MPC block of the same tick.  The split exists for the thread-per-instance
 * fast path (wc_ik_tick_is_split() == 1: CoM as constraint, qpOASES-based solver); any other configuration does the
 * whole tick in the SOLVE call and nothing in the KINEMATICS call. */
enum { WC_IK_PART_KINEMATICS = 1, WC_IK_PART_SOLVE = 2 };
wc_status wc_ik_tick_part_dev(wc_ik* h, int32_t part,
                              const double* q_meas, const double* q_lin,
                              const double* base_H, const uint8_t* left_is_fixed,
                              const double* lf_H_des, const double* rf_H_des,
                              const double* lf_twist, const double* rf_twist,
                              const double* com_pos_des, const double* com_vel_des,
                              const double* com_meas, const double* neck_R_des, double* com_out,
                              double* dq, double* lf_err, double* rf_err,
                              int32_t* qp_status, int32_t* nwsr, void* stream);
int32_t wc_ik_tick_is_split(const wc_ik* h);
wc_status wc_ik_tick_host(wc_ik* h,
                          const double* q_meas, const double* q_lin,
                          const double* base_H, const uint8_t* left_is_fixed,
                          const double* lf_H_des, const double* rf_H_des,
                          const double* lf_twist, const double* rf_twist,
                          const double* com_pos_des, const double* com_vel_des,
                          const double* com_meas, const double* neck_R_des,
                          double* dq, double* lf_err, double* rf_err,
                          int32_t* qp_status, int32_t* nwsr);

/* Split form of wc_ik_tick_host (see wc_mpc_tick_host_async). */
wc_status wc_ik_tick_host_async(wc_ik* h,
                          const double* q_meas, const double* q_lin,
                          const double* base_H, const uint8_t* left_is_fixed,
                          const double* lf_H_des, const double* rf_H_des,
                          const double* lf_twist, const double* rf_twist,
                          const double* com_pos_des, const double* com_vel_des,
                          const double* com_meas, const double* neck_R_des,
                          double* dq, double* lf_err, double* rf_err,
                          int32_t* qp_status, int32_t* nwsr);
wc_status wc_ik_sync(wc_ik* h);
wc_status wc_ik_sync_one(wc_ik* h);

/* Forward kinematics only (WalkingFK getters, WalkingForwardKinematics.cpp:411-454):
 * for each instance the world transforms of l_sole, r_sole [12*B each], neck rotation [9*B],
 * CoM position [3*B] and the four Jacobians in MIXED representation, row-major per
 * instance, field-major across instances: J_lf, J_rf, J_neck [6*(nJ+6)*B], J_com [3*(nJ+6)*B].
 * Device pointers; any output may be NULL. */
wc_status wc_fk_dev(wc_ik* h, const double* q, const double* base_H, const uint8_t* left_is_fixed,
                    double* lf_H, double* rf_H, double* neck_R, double* com,
                    double* J_lf, double* J_rf, double* J_neck, double* J_com, void* stream);

/* QP-IK with CALLER-SUPPLIED kinematics (the WalkingQPIK drop-in at B=1 / arbitrary models):
 * bypasses the FK kernel; same QP build + solve.  Jacobians 6x(nJ+6) (neck: full 6 rows, rows 3-5
 * are used, WalkingQPInverseKinematics.cpp:214-215), J_com 3x(nJ+6); lf_H, rf_H [12*B] and
 * neck_R [9*B] are the CURRENT transforms (setRobotState).  Device pointers. */
wc_status wc_ik_solve_with_jacobians_dev(wc_ik* h,
                         const double* q_meas,
                         const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                         const double* J_lf, const double* J_rf, const double* J_neck, const double* J_com,
                         const double* lf_H_des, const double* rf_H_des,
                         const double* lf_twist, const double* rf_twist,
                         const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                         double* dq, double* lf_err, double* rf_err,
                         int32_t* qp_status, int32_t* nwsr, void* stream);

/* HOST-pointer variants of the two entry points above (copies inside, synchronous on return):
 * what the B=1 drop-in classes WalkingFK / WalkingQPIK_qpOASES in walking-controllers_b200/host use. */
wc_status wc_fk_host(wc_ik* h, const double* q, const double* base_H, const uint8_t* left_is_fixed,
                     double* lf_H, double* rf_H, double* neck_R, double* com,
                     double* J_lf, double* J_rf, double* J_neck, double* J_com);
wc_status wc_ik_solve_with_jacobians_host(wc_ik* h,
                         const double* q_meas,
                         const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                         const double* J_lf, const double* J_rf, const double* J_neck, const double* J_com,
                         const double* lf_H_des, const double* rf_H_des,
                         const double* lf_twist, const double* rf_twist,
                         const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                         double* dq, double* lf_err, double* rf_err,
                         int32_t* qp_status, int32_t* nwsr);

/* ---- WalkingQPIK_osqp: the reference selects between its two WalkingQPIK subclasses at run time
 * (`use_osqp`, WalkingModule.cpp:124-126, 721-740); so does the handle.  With WC_IK_SOLVER_OSQP every tick entry point
 * (wc_ik_tick_*, wc_ik_solve_with_jacobians_*) builds the QP of WalkingQPInverseKinematics_osqp.cpp:135-347 (extra k_attFoot
 * on the neck gradient, bare-twist bounds for a foot with zero desired twist, EMPTY velocity-limit rows) and solves it with
 * the batched OSQP-equivalent ADMM; the OSQP workspace of every instance (scaled iterates, rho, Hessian pattern) persists
 * between ticks as osqp-eigen's solver object does.  Outputs then mean: qp_status = OSQP status (WC_OSQP_*), nwsr = OSQP
 * iterations.  settings NULL = OSQP defaults (what osqp-eigen leaves in place besides verbosity).  Switching solver or
 * calling wc_ik_reset_solver drops the workspaces (a new WalkingQPIK_osqp object). */
enum { WC_IK_SOLVER_QPOASES = 0, WC_IK_SOLVER_OSQP = 1 };
wc_status wc_ik_set_solver(wc_ik* h, int32_t solver, const wc_osqp_settings* settings);
wc_status wc_ik_reset_solver(wc_ik* h);
/* Optional per-instance outputs of the osqp variant, device pointers [B] (NULL = not wanted), written by the next ticks:
 * flags bit 0 = m_optimizerSolver->solve() false (non-convex; with osqp_eigen_checks_status any status but OSQP_SOLVED),
 * bit 1 = isSolutionFeasible false (:395-408; the
 * reference's solve() returns false if any bit is set); setups = osqp_setup calls so far (first solve + every change of
 * the Hessian's sparsity pattern, i.e. of the stance foot). */
wc_status wc_ik_osqp_outputs_dev(wc_ik* h, int32_t* flags, int32_t* setups);
/* the same two arrays of the last tick, copied to host pointers [B] (synchronises the device; either may be NULL) */
wc_status wc_ik_osqp_info_host(wc_ik* h, int32_t* flags, int32_t* setups);

int64_t wc_ik_launch_count(const wc_ik* h);
wc_status wc_ik_enable_profiling(wc_ik* h, int32_t on);
wc_status wc_ik_kernel_times_ms(wc_ik* h, double* ik_ms, int64_t* launches);
/* Split of ik_ms of the window closed by the last wc_ik_kernel_times_ms call: fused-tick fast kernel /
 * general null-space + active-set kernel (the compacted remainder, or everything when the fast path is off). */
wc_status wc_ik_kernel_times_detail_ms(wc_ik* h, double* fast_ms, double* general_ms);
/* Path statistics of the last fused tick (synchronises the device): how many instances took the second pass of the
 * fast kernel (their equality-constrained optimum violated a velocity bound: first working-set change in closed form)
 * and how many went on to the general null-space + Goldfarb-Idnani kernel.  Both 0 when the fast path is off. */
wc_status wc_ik_path_counts_host(wc_ik* h, int32_t* second_pass, int32_t* general);

/* ---------------------------------------------------------------------------------- */
/* Fused host tick of both blocks: what one WalkingModule::updateModule iteration hands  */
/* to WalkingController and WalkingQPIK (WalkingModule.cpp:601-637, 367-425), uploaded    */
/* ONCE.  m_leftTrajectory.front() / m_rightTrajectory.front() serve the MPC's hull, the */
/* IK's desired feet and (base_H == NULL) the stance-foot base transform; the DCM         */
/* reference travels as one new sample per tick (wc_mpc_tick_ring_dev).  836 B instead of */
/* 1 523 B per instance and tick at nJ = 23.  The MPC kernels and the FK + QP-IK kernels  */
/* run on two streams of the MPC handle's pipe; staging is double-buffered as for         */
/* wc_mpc_tick_host_async, and wc_mpc_sync_one / wc_mpc_sync retire fused ticks too.      */
/* All pointers are HOST pointers (pinned for true asynchrony), layouts as in the         */
/* per-block entry points; outputs may be NULL.                                           */
/* Arrays that are contiguous in host memory in the PACKED ORDER are moved by one copy:   */
/*   inputs : lf_H, rf_H, dcm_meas, ref_newest, q_meas, q_lin, lf_twist, rf_twist,        */
/*            com_pos_des, com_vel_des, com_meas, neck_R_des [, base_H] [, ref_window];   */
/*            then l_contact, r_contact, left_is_fixed [, reset], each padded to 8 bytes  */
/*   outputs: zmp, dq, lf_err, rf_err, osqp_status, iters, flags, qp_status, nwsr         */
/*            (int32 arrays padded to 8 bytes)                                            */
/* so a driver that keeps a tick's inputs in one pinned arena pays two uploads and one    */
/* download per tick instead of sixteen and nine (small copies reach half the PCIe rate). */
/* ---------------------------------------------------------------------------------- */
typedef struct wc_tick_host_in {
    const double *lf_H, *rf_H;                    /* [12*B] desired feet transforms (hull, IK targets, stance base) */
    const uint8_t *l_contact, *r_contact;         /* [B] */
    const uint8_t *left_is_fixed;                 /* [B] */
    const uint8_t *reset;                         /* [B] resetTrajectory, or NULL */
    const double *dcm_meas;                       /* [2*B] */
    const double *ref_newest;                     /* [2*B] newest sample of the DCM window (NULL iff ref_window) */
    const double *ref_window;                     /* [(N+1)*2*B] whole window: first tick / after a re-plan, else NULL */
    const double *q_meas, *q_lin;                 /* [nJ*B] */
    const double *base_H;                         /* [12*B] or NULL = stance foot's transform */
    const double *lf_twist, *rf_twist;            /* [6*B] */
    const double *com_pos_des, *com_vel_des, *com_meas; /* [3*B] */
    const double *neck_R_des;                     /* [9*B] */
} wc_tick_host_in;
typedef struct wc_tick_host_out {
    double* zmp; int32_t *osqp_status, *iters, *flags;          /* as wc_mpc_tick_host */
    double *dq, *lf_err, *rf_err; int32_t *qp_status, *nwsr;    /* as wc_ik_tick_host */
} wc_tick_host_out;
wc_status wc_tick_host_async(wc_mpc* mpc, wc_ik* ik, const wc_tick_host_in* in, const wc_tick_host_out* out);

/* ---------------------------------------------------------------------------------- */
/* Closed-loop glue between the two blocks: replaces, batched,     */
/* StableDCMModel (StableDCMModel.cpp:58-90), WalkingZMPController::evaluateControl      */
/* (WalkingZMPController.cpp:126-172) and the joint-velocity integration of updateModule */
/* (WalkingModule.cpp:741-744); iCub::ctrl::Integrator restated as the Tustin rule.      */
/* Gain scheduling (WalkingZMPController::setPhase, :107-126) is wc_glue_set_phase_dev;   */
/* iCub::ctrl::minJerkTrajGen (external, ctrlLib) is restated as the Tustin              */
/* discretisation of its third-order filter.                                             */
/* ---------------------------------------------------------------------------------- */
typedef struct wc_glue wc_glue;
typedef struct wc_glue_config {
    int32_t n_dofs;        /* actuated joints integrated by wc_glue_integrate_joints_dev */
    int32_t use_gain_scheduling; /* useGainScheduling (zmpControllerParams.ini:2; on in the shipped iCubGazeboV2_5 /
                                    icubGazeboSim configurations): the gains start at the stance values and follow
                                    wc_glue_set_phase_dev; off: the walking values, constant (:86-90) */
    double sampling_time;  /* integrators' Ts */
    double com_height;     /* omega = sqrt(gravity / com_height) (StableDCMModel.cpp:31-37) */
    double gravity;
    double k_zmp;          /* kZMP_walking (zmpControllerParams.ini) */
    double k_com;          /* kCoM_walking */
    double smoothing_time; /* smoothingTime: T of the two minJerkTrajGen smoothers (:63-81) */
    double k_zmp_stance;   /* kZMP_stance */
    double k_com_stance;   /* kCoM_stance */
} wc_glue_config;
void wc_glue_default_config(wc_glue_config* cfg); /* iCubGazeboV2_5 values */
wc_status wc_glue_create(const wc_glue_config* cfg, int32_t B, int32_t device, wc_glue** out);
void wc_glue_destroy(wc_glue* h);
/* StableDCMModel::reset + WalkingZMPController::reset (both with com_xy [2*B], WalkingModule.cpp:470-471) and the
 * joint integrator reset (q [nJ*B]); either pointer may be NULL to leave that part untouched.  Device pointers. */
wc_status wc_glue_reset_dev(wc_glue* h, const double* com_xy, const double* q, void* stream);
/* One tick of: m_stableDCMModel->setInput(dcm_des); integrateModel(); ZMP-CoM controller with feedback
 * (zmp_meas [2*B], com_meas [>=2*B], xy rows used) and reference (zmp_des [2*B] = MPC output, LIPM CoM position /
 * velocity) -> com_pos_des, com_vel_des [3*B] (z rows = com_z_des / com_z_vel_des [B], 0 if NULL) as passed to
 * setDesiredCoMPosition / setDesiredCoMVelocity.  k_zmp, k_com [B] or NULL (configured constants). */
wc_status wc_glue_com_reference_dev(wc_glue* h, const double* dcm_des, const double* zmp_des, const double* zmp_meas,
                                    const double* com_meas, const double* com_z_des, const double* com_z_vel_des,
                                    const double* k_zmp, const double* k_com, double* com_pos_des, double* com_vel_des,
                                    void* stream);
/* WalkingZMPController::setPhase (WalkingZMPController.cpp:107-126), called once per tick before
 * wc_glue_com_reference_dev as updateModule does (WalkingModule.cpp:657-661): the stance flag is stance[i] if given, else
 * |dcm_vel_des| < 0.001 (the caller's threshold, WalkingModule.cpp:659-660) from dcm_vel_des [2*B]; both smoothers advance
 * one step towards the stance / walking gains and the per-instance (kZMP, kCoM) they output are what
 * wc_glue_com_reference_dev uses when its k_zmp / k_com arguments are NULL.  A no-op unless use_gain_scheduling.
 * minJerkTrajGen(1, Ts, T) restated: F(s) = -a / (s^3 - c s^2 - b s - a) with a = -150.765868956161 / T^3,
 * b = -84.9812819469538 / T^2, c = -15.9669610709384 / T (90 % of the step at t = T), Tustin-discretised, initialised at
 * steady state (init(y0): all past inputs and outputs y0). */
wc_status wc_glue_set_phase_dev(wc_glue* h, const double* dcm_vel_des, const uint8_t* stance, void* stream);
/* m_qDesired = integrator(m_dqDesired): dq [nJ*B] -> q_des [nJ*B] (may be NULL: state only). */
wc_status wc_glue_integrate_joints_dev(wc_glue* h, const double* dq, double* q_des, void* stream);
int64_t wc_glue_launch_count(const wc_glue* h);

/* ---- the alternative DCM controller, the measured-ZMP / DCM evaluation and the logger row ---- */
/* WalkingDCMReactiveController::evaluateControl (WalkingDCMReactiveController.cpp:63-81), stateless:
 *   zmp_des = dcm_des - dcm_vel_des / omega - kDCM (dcm_des - dcm_meas)      all [2*B]
 * omega = sqrt(gravity / com_height) of the handle (:38-42); k_dcm [B] per-instance gains or NULL (k_dcm_const, `kDCM`
 * of dcmReactiveControllerParams.ini). */
wc_status wc_glue_reactive_dcm_dev(wc_glue* h, double k_dcm_const, const double* k_dcm, const double* dcm_des,
                                   const double* dcm_vel_des, const double* dcm_meas, double* zmp_des, void* stream);
/* same with host pointers (synchronous; what the host class WalkingDCMReactiveController calls) */
wc_status wc_glue_reactive_dcm_host(wc_glue* h, double k_dcm_const, const double* k_dcm, const double* dcm_des,
                                    const double* dcm_vel_des, const double* dcm_meas, double* zmp_des);
/* WalkingFK::evaluateDCM, unfiltered branch (WalkingForwardKinematics.cpp:312-340): com, com_vel [>=2*B] -> dcm [2*B]. */
wc_status wc_glue_evaluate_dcm_dev(wc_glue* h, const double* com, const double* com_vel, double* dcm, void* stream);
/* WalkingFK::evaluateCoM + evaluateDCM with the joint velocities (WalkingForwardKinematics.cpp:278-340; the measured DCM that
 * WalkingController::setFeedback receives, WalkingModule.cpp:559-570): com [3*B] and J_com [3*(nJ+6)*B] from wc_fk_*, dq [nJ*B]
 * measured joint velocities; CoM velocity = J_com [0; dq] (zero base twist, :258-276).  cut_frequency > 0 runs the two
 * first-order low-pass filters of WalkingFK (iCub::ctrl::FirstOrderLowPassFilter, external, restated; state in the handle,
 * wc_glue_reset_com_filter_dev = their init at (0, 0, com_height) / 0, :153-160); use_filters selects the filtered pair for the
 * outputs (`use_filters` of forwardKinematics.ini).  Outputs may be NULL: com_out, com_vel_out [3*B], dcm_out [2*B]. */
wc_status wc_glue_evaluate_com_dev(wc_glue* h, const double* com, const double* J_com, const double* dq, double cut_frequency,
                                   int32_t use_filters, double* com_out, double* com_vel_out, double* dcm_out, void* stream);
wc_status wc_glue_evaluate_com_host(wc_glue* h, const double* com, const double* J_com, const double* dq, double cut_frequency,
                                    int32_t use_filters, double* com_out, double* com_vel_out, double* dcm_out);
wc_status wc_glue_reset_com_filter_dev(wc_glue* h, void* stream);
/* WalkingModule::evaluateZMP (WalkingModule.cpp:824-876): l_wrench, r_wrench [6*B] (force then torque, foot frame),
 * lf_H, rf_H [12*B] -> zmp [2*B]; ok [B] (may be NULL) = 0 where the reference returns false (total fz < 0.1). */
wc_status wc_glue_evaluate_zmp_dev(wc_glue* h, const double* l_wrench, const double* r_wrench, const double* lf_H,
                                   const double* rf_H, double* zmp, int32_t* ok, void* stream);
/* The row WalkingModule::updateModule sends to the WalkingLogger (WalkingModule.cpp:793-803) in the column order of
 * startRecord (:1231-1250): rows [WC_LOG_COLUMNS*B], column c of instance i at [c*B+i].  Poses are [12*B] transforms
 * written as position + Rotation::asRPY; NULL inputs give zero columns.  The text format of the logger module
 * (WalkingLoggerModule.cpp:54-90,142-171) is written by the host class WalkingLogger. */
#define WC_LOG_COLUMNS 53
typedef struct wc_log_inputs {
    const double *dcm, *dcm_des, *dcm_vel_des, *zmp, *zmp_des; /* [2*B] */
    const double *com;                                         /* [3*B] */
    const double *com_des_xy, *com_vel_des_xy;                 /* [>=2*B] */
    const double *lf_H, *rf_H, *lf_H_des, *rf_H_des;           /* [12*B] */
    const double *lf_err, *rf_err;                             /* [6*B] getLeftFootError / getRightFootError */
} wc_log_inputs;
wc_status wc_glue_log_pack_dev(wc_glue* h, const wc_log_inputs* in, double* rows, void* stream);

/* ---------------------------------------------------------------------------------- */
/* On-device producer of the path's per-tick inputs: the signals  */
/* WalkingModule reads from TrajectoryGenerator (TrajectoryGenerator.cpp:501-598) for a */
/* batch of periodic gaits.  The reference's producer is the UnicyclePlanner library    */
/* (external); this is the synthetic workload of SURVEY §8(d) with the shipped planner  */
/* parameters (plannerParams.ini), not a restatement of the unicycle footstep planner.  */
/* ---------------------------------------------------------------------------------- */
typedef struct wc_traj wc_traj;
typedef struct wc_traj_config {
    int32_t step_ticks;     /* nominalDuration / sampling_time (0.9 s / 0.01 s) */
    int32_t ds_ticks;       /* double-support ticks per step: switchOverSwingRatio r -> r / (1 + r) of the step */
    int32_t n_steps;        /* footsteps tabulated per instance (k = -2 .. n_steps-1); queries beyond clamp */
    int32_t reserved;
    double sampling_time;   /* sampling_time */
    double com_height;      /* com_height: omega = sqrt(gravity / com_height), CoM height trajectory */
    double gravity;
    double nominal_width;   /* nominalWidth */
    double step_height;     /* stepHeight */
    double zmp_delta[2];    /* leftZMPDelta = (dx, -dy), rightZMPDelta = (dx, +dy) */
} wc_traj_config;
/* iCubGazeboV2_5 plannerParams.ini; n_steps sized for `n_ticks` ticks + `horizon` look-ahead samples */
void wc_traj_default_config(wc_traj_config* cfg, int32_t n_ticks, int32_t horizon);
wc_status wc_traj_create(const wc_traj_config* cfg, int32_t B, int32_t device, wc_traj** out);
void wc_traj_destroy(wc_traj* h);
/* Per-instance gait (device pointers): phase [B] (tick offset into the step cycle, NULL = 0), step_len [B] (m),
 * dtheta [B] (heading change per step, NULL = 0), theta0 [B] (initial heading, NULL = 0), first_left [B]
 * (1 = even footsteps are left, NULL = 1).  Builds the footstep / ZMP / end-of-step DCM tables. */
wc_status wc_traj_set_gait_dev(wc_traj* h, const int32_t* phase, const double* step_len, const double* dtheta,
                               const double* theta0, const uint8_t* first_left, void* stream);
/* DCM reference position / velocity (either may be NULL) of samples first_sample .. first_sample+n_samples-1:
 * sample s, axis a of instance i at [(2*(s-first_sample)+a)*B+i].  With first_sample = t, n_samples = horizon+1 the
 * position rows are the `dcm_ref` argument of wc_mpc_tick_dev at tick t (getDCMPositionTrajectory, :501-511). */
wc_status wc_traj_dcm_dev(wc_traj* h, int32_t first_sample, int32_t n_samples, double* dcm_pos, double* dcm_vel, void* stream);
/* Reference DCM at `tick` + N(0, sigma^2) per axis, Philox-4x32-10 with key = seed, counter = (instance, tick): dcm [2*B]. */
wc_status wc_traj_measured_dcm_dev(wc_traj* h, int32_t tick, double sigma, uint64_t seed, double* dcm, void* stream);
/* Feet signals at `tick` (any pointer may be NULL): getFeetTrajectories, getFeetTwist (mixed), getFeetStandingPeriods,
 * getWhenUseLeftAsFixed, getCoMHeightTrajectory / getCoMHeightVelocity (:513-598). */
typedef struct wc_traj_feet {
    double *lf_H, *rf_H;             /* [12*B] */
    double *lf_twist, *rf_twist;     /* [6*B] */
    uint8_t *l_contact, *r_contact;  /* [B] */
    uint8_t *left_is_fixed;          /* [B] */
    double *com_height, *com_height_vel; /* [B] */
} wc_traj_feet;
wc_status wc_traj_feet_dev(wc_traj* h, int32_t tick, const wc_traj_feet* out, void* stream);
/* Host pointers, synchronous (drop-in class TrajectoryGenerator): sets the gait and returns n_samples ticks of every
 * signal, element e of sample s of instance i at [(s*E+e)*B+i] (E = 2, 12, 6, 1); outputs may be NULL. */
wc_status wc_traj_generate_host(wc_traj* h, const int32_t* phase, const double* step_len, const double* dtheta, const double* theta0,
                                const uint8_t* first_left, int32_t first_sample, int32_t n_samples, double* dcm_pos, double* dcm_vel,
                                double* lf_H, double* rf_H, double* lf_twist, double* rf_twist, uint8_t* l_contact,
                                uint8_t* r_contact, uint8_t* left_is_fixed, double* com_height, double* com_height_vel);
int64_t wc_traj_launch_count(const wc_traj* h);
/* Optional device-side tick source: the int32 at tick_dev is ADDED, on the device, to every `tick` / `first_sample`
 * argument of the calls above (NULL switches it off).  With it a tick's launch sequence does not depend on the tick index
 * and can be replayed from a CUDA graph (wc_loop_*). */
wc_status wc_traj_set_tick_source_dev(wc_traj* h, const int32_t* tick_dev);

/* ---------------------------------------------------------------------------------- */
/* The closed-loop control tick as ONE call: producer -> measured   */
/* CoM (FK) -> DCM-MPC -> ZMP-CoM controller with gain scheduling -> FK + QP-IK -> joint  */
/* integrator, in the order and with the data dependence of WalkingModule::updateModule   */
/* (WalkingModule.cpp:541-745): the IK of tick t consumes the MPC output of tick t.  The  */
/* plant is synthetic: perfect joint tracking (q_meas = q_des), measured DCM = reference  */
/* + N(0, sigma^2) (+ an impulsive push per instance, BASELINE configs[3]), measured ZMP  */
/* = desired ZMP + offset.  The handle BORROWS the four block handles (same B, same       */
/* device; they must outlive it and must not be ticked directly while it exists) and      */
/* installs a device-side tick counter in the MPC and the producer, so that a tick's      */
/* launch sequence is constant and is replayed from a CUDA graph (use_graph).             */
/* ---------------------------------------------------------------------------------- */
typedef struct wc_loop wc_loop;
typedef struct wc_loop_config {
    double dcm_noise_sigma;  /* std of the Gaussian noise on the measured DCM (0.002 m) */
    uint64_t noise_seed;     /* Philox key of that noise */
    double zmp_meas_offset;  /* measured ZMP = desired ZMP + offset on both axes */
    int32_t use_graph;       /* 1: ticks after the first are replays of one captured CUDA graph */
    int32_t reserved;
} wc_loop_config;
void wc_loop_default_config(wc_loop_config* cfg);
wc_status wc_loop_create(wc_mpc* mpc, wc_ik* ik, wc_glue* glue, wc_traj* traj, const wc_loop_config* cfg, int32_t device,
                         wc_loop** out);
void wc_loop_destroy(wc_loop* h);
/* Start of a rollout: tick counter 0, MPC reset (WalkingController::reset), integrators at the posture q0 [nJ*B] (device
 * pointer) and at its CoM about the stance foot of tick 0 (WalkingModule.cpp:470-471).  The gait must be set on `traj`. */
wc_status wc_loop_reset_dev(wc_loop* h, const double* q0, void* stream);
/* Impulsive DCM disturbance (BASELINE configs[3]): at tick push_tick[i] (< 0: never) the measured DCM of instance i is
 * displaced by push_dxy [2*B].  Device pointers, copied. */
wc_status wc_loop_set_push_dev(wc_loop* h, const int32_t* push_tick, const double* push_dxy, void* stream);
/* n_ticks closed-loop ticks, asynchronous on `stream`. */
wc_status wc_loop_run_dev(wc_loop* h, int32_t n_ticks, void* stream);
int64_t wc_loop_tick_count(const wc_loop* h);
/* Device pointers to the signals of the LAST tick run (valid until wc_loop_destroy; read them on the rollout's stream):
 * layouts as the outputs / inputs of the block entry points. */
typedef struct wc_loop_view {
    const double *zmp, *dq, *q_des, *com_pos_des, *com_vel_des, *com_meas, *dcm_meas, *lf_err, *rf_err; /* [2|nJ|nJ|3|3|3|2|6|6][B] */
    const double *lf_H, *rf_H, *ref_pos;            /* producer outputs of the tick: [12*B], [12*B], [(N+1)*2*B] */
    const int32_t *osqp_status, *iters, *flags;     /* MPC outputs [B] */
    const int32_t *qp_status, *nwsr;                /* QP-IK outputs [B] */
    /* the remaining inputs the two blocks consumed in that tick (what a host that feeds them would have uploaded) */
    const double *lf_twist, *rf_twist, *base_H, *neck_R_des;   /* [6*B], [6*B], [12*B], [9*B] */
    const uint8_t *l_contact, *r_contact, *left_is_fixed;      /* [B] */
} wc_loop_view;
wc_status wc_loop_view_get(wc_loop* h, wc_loop_view* view);

/* ---------------------------------------------------------------------------------- */
/* Result collection across the GPUs of one box (SURVEY §8(b) `wc_gather`, §8(e)).       */
/* Instances are independent, rank g owns a contiguous instance range and there is NO     */
/* collective on the tick path; at its collection points a Monte-Carlo driver all-gathers  */
/* the per-instance results and all-reduces a small statistics vector over NCCL (NVLink 5  */
/* / NVSwitch), on the communicator's own side stream so that the next tick overlaps it.   */
/* NCCL is bound at run time (dlopen of libnccl.so.2); WC_ERR_UNSUPPORTED if it is absent. */
/* ---------------------------------------------------------------------------------- */
typedef struct wc_comm wc_comm;
#define WC_COMM_ID_BYTES 128
/* rank 0 creates the id (ncclGetUniqueId) and hands it to the other ranks by whatever means the driver has (a file, MPI,
 * a torch.distributed store); every rank then joins with it. */
wc_status wc_comm_unique_id(uint8_t* id /*[WC_COMM_ID_BYTES]*/);
wc_status wc_comm_create(const uint8_t* id, int32_t world, int32_t rank, int32_t device, wc_comm** out);
void wc_comm_destroy(wc_comm* c);
int32_t wc_comm_world(const wc_comm* c);
int32_t wc_comm_rank(const wc_comm* c);
/* Device pointers.  Every pair (x, x_all): this rank's block [rows*B] and the gathered array [world][rows*B] (rank-major:
 * rank r's field-major block at offset r*rows*B); a NULL pair is skipped.  stats_sum / stats_max are reduced IN PLACE. */
typedef struct wc_gather_io {
    const double *zmp; double *zmp_all;                      /* rows = 2 */
    const double *dq; double *dq_all;                        /* rows = n_dofs */
    const int32_t *osqp_status; int32_t *osqp_status_all;    /* rows = 1 */
    const int32_t *iters; int32_t *iters_all;
    const int32_t *qp_status; int32_t *qp_status_all;
    const int32_t *nwsr; int32_t *nwsr_all;
    double *stats_sum; int32_t n_stats_sum;                  /* all-reduce (sum) */
    double *stats_max; int32_t n_stats_max;                  /* all-reduce (max) */
} wc_gather_io;
/* One NCCL group on the communicator's side stream, ordered after the work already enqueued on producer_stream. */
wc_status wc_gather_async(wc_comm* c, const wc_gather_io* io, int32_t B, int32_t n_dofs, void* producer_stream);
/* block_host = 0: consumer_stream waits for the last wc_gather_async; block_host = 1: the calling thread does. */
wc_status wc_gather_wait(wc_comm* c, void* consumer_stream, int32_t block_host);

/* Measured FP64 FMA throughput of the device (register-resident FMA chains, all SMs), in GFLOP/s:
 * the denominator of the FP64 roofline fraction that bench.py reports next to the HBM one
 * (MEASURED_PEAKS.json carries no FP64 entry). */
wc_status wc_measure_fp64_peak(int32_t device, double* gflops);

/* Library/ABI version string, e.g. "wc_b200 0.1 sm_100a". */
const char* wc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WC_CAPI_H */
