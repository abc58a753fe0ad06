// ik_kernels.cuh — batched iCub floating-base FK/Jacobians + QP inverse kinematics for sm_100a.
//
// Replaces, per controller instance and per 10 ms tick,
//   WalkingFK (iDynTree::KinDynComputations)   modules/Walking_module/src/WalkingForwardKinematics.cpp:217-276,411-454
//   WalkingQPIK setters                        modules/Walking_module/src/WalkingQPInverseKinematics.cpp:121-256
//   WalkingQPIK_qpOASES::{setHessianMatrix,setGradientVector,setLinearConstraintMatrix,setBounds,solve,
//                         getSolution,getLeftFootError,getRightFootError}
//                                              modules/Walking_module/src/WalkingQPInverseKinematics_qpOASES.cpp:135-401
//   qpOASES::SQProblem::{init,hotstart}        (external)
//
// One warp per instance, lane j <-> QP variable j (6 base + nJ joints <= 32).  The Hessian
// H = Lambda + Jn' Wn Jn (+ Jc' Wc Jc) is never formed: only its diagonal-plus-rank-3 action is used.
// The equality-constrained QP is solved exactly by the null-space method (Householder QR of A' held
// column-per-lane in registers, reflectors broadcast by shuffles; Cholesky of Z'HZ), and instances
// whose velocity bounds are violated continue with the Goldfarb-Idnani dual active-set method in
// the reduced space — the same exact algorithm as the CPU oracle's qpOASES stand-in.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/wc_capi.h"

namespace wcb {

constexpr int WC_MAX_JUMP = 6; // 2^6 >= WC_MAX_LINKS

struct IkModelDev {
    int n_links, n_dofs, max_depth, pad0;
    int parent[WC_MAX_LINKS];
    int dof[WC_MAX_LINKS];
    int depth[WC_MAX_LINKS];
    int link_of_dof[WC_MAX_DOFS];
    double joint_R[WC_MAX_LINKS][9];
    double joint_p[WC_MAX_LINKS][3];
    double axis[WC_MAX_LINKS][3];
    double mass[WC_MAX_LINKS];
    double com[WC_MAX_LINKS][3];
    double sub_mass[WC_MAX_LINKS];
    unsigned long long sub_mask[WC_MAX_LINKS];
    double total_mass;
    int base_link[2];          // [left_is_fixed]
    int frame_link[3];         // l_sole, r_sole, neck
    int pad1;
    double frame_R[3][9];
    double frame_p[3][3];
    // per base choice and dof: path signs of the three task frames, and "base inside the joint's subtree"
    signed char s_frame[2][3][WC_MAX_DOFS];
    signed char base_in_sub[2][WC_MAX_DOFS];
    // pointer-jumping schedule of the transform composition: jump[r][l] = the ancestor link l composes with in
    // round r (its 2^r-th ancestor), -1 = l is already expressed in the root frame; n_jump rounds in all
    int n_jump, pad2;
    signed char jump[WC_MAX_JUMP][WC_MAX_LINKS];
    // thread-per-instance fast path (ik_tpi.cu): schedule of the serial transform composition.  tpi_slot[l] = shared-memory
    // slot root_H_l is parked in (-1: nobody reads it later), tpi_pslot[l] = where link l finds its parent's transform
    // (slot id, kTpiPrev: still in registers because parent == l - 1, -1: link 0); the sub-tree of the link of dof d is the
    // contiguous run of links [link_of_dof[d], tpi_sub_last[d]] (depth-first numbering; checked on the host, tpi_ok = 0 otherwise)
    int tpi_ok, tpi_slots;
    signed char tpi_slot[WC_MAX_LINKS];
    signed char tpi_pslot[WC_MAX_LINKS];
    signed char tpi_sub_last[WC_MAX_DOFS];
};
constexpr int kTpiPrev = -2;
constexpr int kTpiMaxSlots = 8;

struct IkDev {
    int B, nJ;
    wc_ik_config cfg;
    const IkModelDev* model; // device memory
    const double* kpos;      // [B] per-instance gains (device) or nullptr
    const double* katt;
    // fused-tick fast path (ik_fast_kernel): instances handed over to the general kernel
    int32_t* slow_list;      // [B] or nullptr (fast path disabled)
    int32_t* slow_count;     // [2]: [0] = entries of slow_list, [1] = entries of second_list
    double* tpi_scratch;     // thread-per-instance fast path: [tpi_rows][B] per-instance scratch in HBM / L2 (ik_tpi.cu), or nullptr
    int tpi_slots;           // transform slots of the model's schedule (IkModelDev::tpi_slots), for the launch's shared-memory size
    int32_t* second_list;    // [2 B] instances whose equality-constrained optimum violates a bound (second fast launch): [0, B) the
                             // instance, [B, 2 B) the bound Goldfarb-Idnani adds first (2 * variable + upper)
    // WalkingQPIK_osqp (wc_ik_set_solver): OSQP settings and the persistent workspace of every instance, field-major
    int solver;              // WC_IK_SOLVER_*
    wc_osqp_settings osqp;
    double* ox; double* oy; double* oz; // scaled iterates [NV][B], [M][B], [M][B]  (M = task rows + nJ velocity rows)
    double* orho;            // [B]
    uint32_t* omask;         // [NV][B] upper-triangular nonzero pattern of the Hessian of the live solver, column per word
    uint8_t* oinit;          // [B] solver exists
    int32_t* osetups;        // [B] number of osqp_setup calls so far
    int32_t* oflags;         // [B]: bit 0 = solver error (non-convex), bit 1 = isSolutionFeasible failed
};

enum { IK_MODE_TICK = 0, IK_MODE_FK_ONLY = 1, IK_MODE_EXT_KIN = 2,
       // the tick as two calls (wc_ik_tick_part_dev): what depends on the joint state / base / desired feet only, then the rest
       IK_MODE_TICK_KIN = 3, IK_MODE_TICK_SOLVE = 4 };

struct IkTickIO {
    // inputs
    const double* q_meas; const double* q_lin; const double* base_H; const uint8_t* left_is_fixed;
    const double* lf_H_des; const double* rf_H_des; const double* lf_twist; const double* rf_twist;
    const double* com_pos_des; const double* com_vel_des; const double* com_meas; const double* neck_R_des;
    // externally supplied kinematics (IK_MODE_EXT_KIN)
    const double* lf_H; const double* rf_H; const double* neck_R;
    const double* J_lf; const double* J_rf; const double* J_neck; const double* J_com;
    // outputs
    double* dq; double* lf_err; double* rf_err; int32_t* status; int32_t* nwsr;
    // FK outputs (IK_MODE_FK_ONLY)
    double* o_lf_H; double* o_rf_H; double* o_neck_R; double* o_com;
    double* o_J_lf; double* o_J_rf; double* o_J_neck; double* o_J_com;
};

// mid (nullable): event recorded between the fast kernel and the general kernel of the fused tick
cudaError_t launch_ik(const IkDev& d, const IkTickIO& io, int mode, cudaStream_t s, cudaEvent_t mid = nullptr);
bool ik_supported(int n_dofs);
// thread-per-instance fast path of the fused tick (ik_tpi.cu)
int ik_tpi_rows(int n_links, int n_dofs);  // rows of IkDev::tpi_scratch
cudaError_t launch_ik_tpi(const IkDev& d, const IkTickIO& io, cudaStream_t s, int part = 0); // part: 0 whole tick, 1 kinematics, 2 solve
cudaError_t launch_ik_tpi_fk(const IkDev& d, const IkTickIO& io, cudaStream_t s); // FK without Jacobians (IkTickIO::o_lf_H, o_rf_H, o_neck_R, o_com)

} // namespace wcb
