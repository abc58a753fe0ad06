// mpc_kernels.cuh — batched DCM-MPC for sm_100a: problem build (K1) + OSQP-equivalent ADMM (K2).
//
// Replaces, per controller instance and per 10 ms tick, the reference chain
//   WalkingController::{setConvexHullConstraint,setFeedback,setReferenceSignal,solve}
//     (modules/Walking_module/src/WalkingDCMModelPredictiveController.cpp:364-521)
//   MPCSolver::{setConstraintsMatrix,setBounds,setGradient,initialize,solve,getSolution}
//     (modules/Walking_module/src/MPCSolver.cpp:76-322)
//   OsqpEigen::Solver -> OSQP (Ruiz scaling, KKT LDL^T, ADMM, adaptive rho; external).
//
// Formulation: exactly the reference's sparse (non-condensed) QP, z = [x_0..x_N, u_0..u_{N-1}],
// n = 4N+2, m = 2(N+1) + n_hull.  OSQP's KKT solve  [[P+sI, A'],[A, -1/rho]] [x~;nu] = rhs  is done
// on the mathematically identical reduced SPD system  (P + sI + A' rho A) x~ = rhs_x + A' rho rhs_z,
// z~ = A x~.  With diagonal Q, R (all shipped configurations) the two axes decouple except for the
// hull rows on u_0, so with the per-axis "chain" order  x_N, u_{N-1}, x_{N-1}, ..., x_1, u_0, x_0
// the reduced matrix is pentadiagonal per axis plus one dense 4x4 corner (u_0, x_0 of both axes).
//
// Data layout: every per-instance quantity lives in HBM as field-major SoA ([field][B]); one
// instance is handled by G cooperating lanes of a warp with its working set staged in shared
// memory; the two axis chains of the sequential triangular solves run on two lanes.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/wc_capi.h"

namespace wcb {

constexpr int kMaxHull = 8;

constexpr double kRhoMin = 1e-6;
constexpr double kRhoMax = 1e6;
constexpr double kRhoEqOverIneq = 1e3;
constexpr double kMinScaling = 1e-4;
constexpr double kMaxScaling = 1e4;

// Device-side view of one handle's state.  All arrays are [rows][B] (element (r,i) at r*B+i).
struct MpcDev {
    int B, N;
    // constants
    double a_dyn, b_dyn;
    double foot_front, foot_back, foot_left, foot_right;
    double init_zmp[2];
    double hull_tol;
    wc_osqp_settings st;
    int rho_interval; // resolved adaptive_rho_interval
    int ring_head;    // slot of stage 0 in the gradient ring for THIS tick
    // per-instance weights (diagonals) and the handle's off-diagonal weight entries (arbitrary 2x2 triplets,
    // WM/src/WalkingDCMModelPredictiveController.cpp:190-208): Qoff = {Q(0,1), Q(1,0)}, Roff = {R(0,1), R(1,0)}.  The Hessian takes
    // the upper-triangular entries the reference's triplets produce (oracle/mpc_ref.c build_P), the gradient the full matrices.
    // coupled != 0 routes the handle to the cooperative kernel's block (2x2) chain solve.
    double Qoff[2], Roff[2];
    int coupled;
    int par; // cooperative kernel: segment-parallel chain solve (long horizons)
    double* Qd; // [2][B]
    double* Rd; // [2][B]
    // persistent solver state (chain order per axis: index a*nA + j)
    double* x;      // [2*nA][B] scaled primal iterate
    double* yeq;    // [2*nR][B] scaled duals of the equality rows (index a*nR + k)
    double* yh;     // [8][B]
    double* zh;     // [8][B]
    double* z0;     // [2][B]  z of equality row 0 (the only equality z that is not identically its bound)
    double* D;      // [2*nA][B]
    double* Eeq;    // [2*nR][B]
    double* Eh;     // [8][B]
    double* cscale; // [B]
    double* rho;    // [B]
    double* grad;   // [2*nR][B] unscaled gradient of the state block, ring-buffered over stages
    double* hullA;  // [16][B]  (row h: [2h], [2h+1])
    double* hullb;  // [8][B]
    double* hullv;  // [16][B]
    double* uprev;  // [2][B]  m_output
    int32_t* nhull;      // [B]
    uint8_t* feet;       // [B] bit0 left, bit1 right (m_feetStatus)
    uint8_t* have_solver;// [B]
    uint8_t* inited;     // [B] OSQP workspace exists (initSolver done)
    int32_t* rho_updates;// [B]
    int32_t* tick_flags; // [B] K1 -> K2 (WC_MPC_FLAG_*), also the tick's flag output
    // chain-per-thread path only (mpc_chain.cu); NULL on the cooperative path
    double* Dinv;        // [2*nA][B] 1/D  (OSQP keeps Dinv, Einv next to D, E)
    double* Einv;        // [2*nR][B]
    double* Ehinv;       // [8][B]
    double* dy_scratch;  // [2*nR + 8][B] delta_y of the last iteration (infeasibility tests only)
    double* L1;          // [2*nA][B] pentadiagonal LDL' of the reduced KKT matrix: first sub-diagonal,
    double* L2;          // [2*nA][B]   second sub-diagonal,
    double* DI;          // [2*nA][B]   inverse pivots (split index like x)
    double* CF;          // [10][B]   dense 4x4 corner factor
    int32_t* setup_list; // [B] instances whose OSQP workspace must be set up this tick (filled by K1)
    int32_t* setup_count;// [1]
    int setup_overlap;   // split-chain path: osqp_setup of the listed instances and their phase 1 run on a side stream next to the phase 1
                         // of everyone else; K1 marks the listed instances with kMpcFlagSetupPending in tick_flags
    int32_t* cont_list;  // [B] instances that did not stop at the first termination check
    int32_t* cont_count; // [1]
    // device-side ring of the DCM reference window (wc_mpc_tick_ring_*): sample k of this tick's window in slot
    // (ring_head + k) mod (N+1), so a host that feeds the controller uploads one new sample per tick, as
    // MPCSolver::setGradient consumes one (MPCSolver.cpp:216-239)
    double* ref_ring;    // [2*(N+1)][B], allocated on first use
    // optional device-side tick counter (wc_mpc_set_tick_source_dev): the ring head is then read on the device, so a tick's
    // launch sequence no longer depends on the tick index and can be replayed from a CUDA graph (wc_loop.cu)
    const int32_t* tick_dev;
};

__device__ __forceinline__ int mpc_ring_head(const MpcDev& d)
{
    return d.tick_dev ? (int)((unsigned)(*d.tick_dev) % (unsigned)(d.N + 1)) : d.ring_head;
}

// running infinity norm as OSQP's vec_norm_inf writes it (`if (|v| > max) max = |v|`): one compare and one select, where fmax on
// doubles is a seven-instruction sequence (there is no fp64 min/max instruction; the rest is NaN handling).  m is never NaN.
__device__ __forceinline__ double amax(double m, double v) { const double av = fabs(v); return av > m ? av : m; }

// internal bit of tick_flags (never reaches the caller): the instance is on this tick's setup_list
constexpr int kMpcFlagSetupPending = 1 << 30;

struct MpcTickIO {
    const double* lf_H; const double* rf_H;
    const uint8_t* l_contact; const uint8_t* r_contact;
    const double* dcm; const double* ref; const uint8_t* reset;
    double* zmp; int32_t* status; int32_t* iters; int32_t* flags;
    int ref_is_ring; // 0: `ref` is this tick's window [(N+1)*2][B]; 1: the window is read from MpcDev::ref_ring
};

// What OsqpEigen::Solver::solve() returns for an OSQP status (wc_osqp_settings::osqp_eigen_checks_status): the 2018
// osqp-eigen returns osqp_solve() == 0 -- true for every status but a non-convex problem; since 0.4 it is false unless the
// status is OSQP_SOLVED.  The callers (MPCSolver.cpp:309 -> Controller.cpp:503-517, ..._osqp.cpp:377) then read OSQP's
// stored solution: the last iterate where osqp_has_solution, NaN otherwise (OSQP store_solution).
__host__ __device__ inline bool osqp_eigen_solve_ok(int status, int checks_status)
{
    if (checks_status) return status == WC_OSQP_SOLVED;
    return status != WC_OSQP_NON_CVX && status != WC_OSQP_UNSOLVED;
}
__host__ __device__ inline bool osqp_has_solution(int status)
{
    return !(status == WC_OSQP_PRIMAL_INFEASIBLE || status == WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
             status == WC_OSQP_DUAL_INFEASIBLE || status == WC_OSQP_DUAL_INFEASIBLE_INACCURATE ||
             status == WC_OSQP_UNSOLVED || status == WC_OSQP_NON_CVX);
}

__host__ __device__ inline int mpc_nA(int N) { return 2 * N + 1; }
__host__ __device__ inline int mpc_nR(int N) { return N + 1; }
// doubles of shared memory per instance in K2
// Segment-parallel chain solve of the cooperative kernel (long horizons): every lane owns one segment of one axis chain, the
// segments are separated by two-unknown separators, the last separator of an axis is its (u_0, x_0) corner.  Needs at least
// two interior unknowns per segment; P = lanes / 2 segments per axis.
__host__ __device__ inline bool mpc_par_ok(int N, int lanes)
{
    const int P = lanes / 2; // two interior unknowns per segment at least; the 10 P Schur entries per axis are staged in the rhs vector
    return lanes >= 8 && (2 * N + 1) - 2 * P >= 2 * P && 10 * lanes <= 2 * (2 * N + 1);
}
constexpr int kParMinHorizon = 40; // below: the serial two-lane sweeps (fewer barriers)
__host__ __device__ inline int mpc_par_doubles(int N) { return 2 * mpc_nA(N) + 2 * 16 * (4 + 8); } // Z + separator rhs parts + reduced factors
__host__ __device__ inline int mpc_smem_doubles(int N, int coupled = 0, int par = 0)
{
    const int nA = mpc_nA(N), nR = mpc_nR(N);
    return (par ? mpc_par_doubles(N) : 0) +
           2 * nA * (coupled ? 10 : 7)   // X, QB, W, L1, L2, DI, DD (+ the off-diagonal entries of the 2x2 blocks of L1, L2, DI)
         + 2 * nR * 6   // YE, DYE, CM, CA, CB, EE
         + 8 * 8        // HX, HY, UH, LH, ZH, YH, DYH, EH
         + 16           // corner factor + scratch
         + 8;           // reduction scratch / pad (keeps the stride == 8 mod 16 for N == 25)
}

void launch_mpc_prepare(const MpcDev& d, const MpcTickIO& io, cudaStream_t s);
// ring maintenance before K1: the newest sample [2][B] into the slot of stage N, or (window != NULL) all N+1 samples
void launch_mpc_ref_ring_update(const MpcDev& d, const double* newest, const double* window, cudaStream_t s);
cudaError_t launch_mpc_admm(const MpcDev& d, const MpcTickIO& io, int lanes_per_instance, cudaStream_t s);
void launch_mpc_export_primal(const MpcDev& d, double* primal_out, cudaStream_t s);
int mpc_admm_max_smem();
// chain-per-thread path (mpc_chain.cu): setup of new workspaces + phase 1 + phase 2, three launches
bool mpc_chain_supported(int N);
bool mpc_chain_preferred(int N);
size_t mpc_chain_smem_bytes(int N);
// mid (nullable): two events recorded after the setup kernel and after phase 1
cudaError_t launch_mpc_chain(const MpcDev& d, const MpcTickIO& io, cudaStream_t s, cudaEvent_t* mid);
cudaError_t launch_mpc_chain_setup(const MpcDev& d, cudaStream_t s, bool factor);
// split-chain path (mpc_split.cu): four lanes per instance, iterate in registers; compile-time horizons
bool mpc_split_supported(int N);
// side / ev_fork / ev_join: stream and events for the overlapped set-up (d.setup_overlap); nullptr = everything on s
cudaError_t launch_mpc_split(const MpcDev& d, const MpcTickIO& io, cudaStream_t s, cudaEvent_t* mid, cudaStream_t side = nullptr,
                             cudaEvent_t ev_fork = nullptr, cudaEvent_t ev_join = nullptr);

} // namespace wcb
