// wc_loop.cu — the closed-loop control tick as ONE call: what one
// iteration of WalkingModule::updateModule does between reading the robot and writing the joint references
// (modules/Walking_module/src/WalkingModule.cpp:541-745), for a batch of independent instances, device resident:
//
//   producer (TrajectoryGenerator getters, :501-598)      wc_traj_feet_dev, wc_traj_dcm_dev, wc_traj_measured_dcm_dev
//   measured CoM about the current posture (:559-563)      wc_fk_dev
//   DCM-MPC (:601-637)                                     wc_mpc_tick_dev            -> desired ZMP
//   ZMP-CoM controller with gain scheduling (:657-679)     wc_glue_set_phase_dev, wc_glue_com_reference_dev (depends on the MPC)
//   FK + QP-IK about the integrated posture (:693-740)     wc_ik_tick_dev             (depends on the ZMP-CoM law)
//   m_qDesired = integrator(m_dqDesired) (:741-744)        wc_glue_integrate_joints_dev
//
// The plant is synthetic (there is no robot here): perfect joint tracking (q_meas = q_des), measured DCM = reference +
// Gaussian noise (+ an optional impulsive push per instance: BASELINE configs[3]), measured ZMP = desired ZMP + offset.
// Every dependence of the reference's tick is kept: the IK of tick t consumes the MPC output of tick t.
//
// The handle only sequences the public *_dev entry points of the four block handles on one stream.  All per-tick
// variation is on the device (tick counter read by the producer kernels and by the MPC's ring head), so the launch
// sequence of a tick is constant and is replayed from a CUDA graph (use_graph): ~17 launches per tick become one.
#include <cmath>
#include <cstdio>
#include <new>

#include <cuda_runtime.h>

#include "../../include/wc_capi.h"

struct wc_loop {
    wc_mpc* mpc; wc_ik* ik; wc_glue* glue; wc_traj* traj;
    wc_loop_config cfg;
    int B, N, nJ, device;
    int64_t tick;          // ticks run since wc_loop_reset_dev
    int64_t launches;
    double* buf;           // one allocation, carved below
    double *lf_H, *rf_H, *lf_tw, *rf_tw, *com_h, *com_hv, *ref_pos, *ref_vel, *dcm_meas, *base_H, *neck_R, *com_meas;
    double *zmp, *zmp_meas, *com_pos_des, *com_vel_des, *dq, *lf_err, *rf_err, *q_des, *push_dxy;
    uint8_t *l_contact, *r_contact, *lif, *reset_all;
    int32_t *osqp_status, *iters, *flags, *qp_status, *nwsr, *push_tick, *tick_dev;
    cudaGraphExec_t graph;
    cudaStream_t own;      // capture / replay stream when the caller's stream is the legacy default stream (not capturable)
    cudaEvent_t ev_in, ev_out;
    // the measured-CoM kinematics (stance base -> FK) depends on the feet and the joint state only: it runs on a side branch next to
    // the producer's DCM kernels and the latency-bound head of the MPC block (K1, osqp_setup), joined before the ZMP-CoM law
    cudaStream_t side;
    cudaEvent_t ev_fork, ev_join;
    bool first;            // the next tick is the first after reset (the MPC rings are re-filled)
};

namespace {

// stance-foot base transform (WalkingFK::evaluateWorldToBaseTransformation, WalkingForwardKinematics.cpp:217-244) and the
// desired neck orientation RotZ(mean yaw of the feet) (WalkingModule.cpp:697-706, 383: the two inversions cancel)
__global__ void loop_pre_kernel(int B, const double* lf_H, const double* rf_H, const uint8_t* lif, double* base_H, double* neck_R)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    const double* src = lif[i] ? lf_H : rf_H;
#pragma unroll
    for (int k = 0; k < 12; k++) base_H[k * Bz + i] = src[k * Bz + i];
    // Rotation::asRPY()(2) = atan2(R10, R00); rows of the 3x4 transform are 4 apart
    const double yl = atan2(lf_H[4 * Bz + i], lf_H[i]), yr = atan2(rf_H[4 * Bz + i], rf_H[i]);
    const double my = atan2(sin(yl) + sin(yr), cos(yl) + cos(yr));
    const double c = cos(my), s = sin(my);
    const double R[9] = {c, -s, 0.0, s, c, 0.0, 0.0, 0.0, 1.0};
#pragma unroll
    for (int k = 0; k < 9; k++) neck_R[k * Bz + i] = R[k];
}

// synthetic plant: impulsive DCM push at the instance's push tick (BASELINE configs[3]), measured ZMP = desired + offset
__global__ void loop_push_kernel(int B, const int32_t* tick_dev, const int32_t* push_tick, const double* push_dxy, double* dcm_meas)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (push_tick[i] == *tick_dev) {
        dcm_meas[i] += push_dxy[i];
        dcm_meas[(size_t)B + i] += push_dxy[(size_t)B + i];
    }
}
__global__ void loop_zmp_meas_kernel(int B, double offset, const double* zmp, double* zmp_meas)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 2 * B) zmp_meas[k] = zmp[k] + offset;
}
__global__ void loop_advance_kernel(int32_t* tick_dev) { *tick_dev += 1; }
__global__ void loop_fill_kernel(int n, uint8_t* p, uint8_t v)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) p[k] = v;
}

#define TRY(x) do { const wc_status rc_ = (x); if (rc_ != WC_OK) return rc_; } while (0)

// the launch sequence of one tick; identical for every tick (the tick index lives in h->tick_dev)
wc_status enqueue_tick(wc_loop* h, bool first, cudaStream_t s)
{
    const int B = h->B, N = h->N;
    const unsigned g128 = (unsigned)((B + 127) / 128);
    void* sv = static_cast<void*>(s);
    wc_traj_feet f{h->lf_H, h->rf_H, h->lf_tw, h->rf_tw, h->l_contact, h->r_contact, h->lif, h->com_h, h->com_hv};
    TRY(wc_traj_feet_dev(h->traj, 0, &f, sv));
    // side branch (also under graph capture: the side stream joins the capture through the event and is joined back below)
    if (cudaEventRecord(h->ev_fork, s) != cudaSuccess || cudaStreamWaitEvent(h->side, h->ev_fork, 0) != cudaSuccess) return WC_ERR_CUDA;
    loop_pre_kernel<<<g128, 128, 0, h->side>>>(B, h->lf_H, h->rf_H, h->lif, h->base_H, h->neck_R);
    // measured CoM: FK about the current (perfectly tracked) posture.  With the split tick (wc_ik_tick_part_dev) the kinematics part
    // of the QP-IK -- evaluated at the same posture -- delivers it, and one pass over the tree serves both
    const bool split = wc_ik_tick_is_split(h->ik) != 0;
    if (split)
        TRY(wc_ik_tick_part_dev(h->ik, WC_IK_PART_KINEMATICS, h->q_des, h->q_des, h->base_H, h->lif, h->lf_H, h->rf_H, h->lf_tw, h->rf_tw,
                                nullptr, nullptr, nullptr, h->neck_R, h->com_meas, nullptr, nullptr, nullptr, nullptr, nullptr,
                                static_cast<void*>(h->side)));
    else
        TRY(wc_fk_dev(h->ik, h->q_des, h->base_H, h->lif, nullptr, nullptr, nullptr, h->com_meas, nullptr, nullptr, nullptr, nullptr,
                      static_cast<void*>(h->side)));
    if (cudaEventRecord(h->ev_join, h->side) != cudaSuccess) return WC_ERR_CUDA;
    TRY(wc_traj_dcm_dev(h->traj, 0, N + 1, h->ref_pos, h->ref_vel, sv));
    TRY(wc_traj_measured_dcm_dev(h->traj, 0, h->cfg.dcm_noise_sigma, h->cfg.noise_seed, h->dcm_meas, sv));
    loop_push_kernel<<<g128, 128, 0, s>>>(B, h->tick_dev, h->push_tick, h->push_dxy, h->dcm_meas);
    // MPC -> desired ZMP
    TRY(wc_mpc_tick_dev(h->mpc, h->lf_H, h->rf_H, h->l_contact, h->r_contact, h->dcm_meas, h->ref_pos, first ? h->reset_all : nullptr,
                        h->zmp, h->osqp_status, h->iters, h->flags, sv));
    // ZMP-CoM controller (consumes the MPC output of THIS tick)
    loop_zmp_meas_kernel<<<(unsigned)((2 * B + 127) / 128), 128, 0, s>>>(B, h->cfg.zmp_meas_offset, h->zmp, h->zmp_meas);
    TRY(wc_glue_set_phase_dev(h->glue, h->ref_vel, nullptr, sv));
    if (cudaStreamWaitEvent(s, h->ev_join, 0) != cudaSuccess) return WC_ERR_CUDA; // base_H, neck_R, com_meas
    TRY(wc_glue_com_reference_dev(h->glue, h->ref_pos, h->zmp, h->zmp_meas, h->com_meas, h->com_h, h->com_hv, nullptr, nullptr,
                                  h->com_pos_des, h->com_vel_des, sv));
    // FK + QP-IK about the integrated posture (consumes the ZMP-CoM law of THIS tick), then the joint integrator
    if (split)
        TRY(wc_ik_tick_part_dev(h->ik, WC_IK_PART_SOLVE, h->q_des, h->q_des, h->base_H, h->lif, h->lf_H, h->rf_H, h->lf_tw, h->rf_tw,
                                h->com_pos_des, h->com_vel_des, h->com_meas, h->neck_R, nullptr, h->dq, h->lf_err, h->rf_err,
                                h->qp_status, h->nwsr, sv));
    else
        TRY(wc_ik_tick_dev(h->ik, h->q_des, h->q_des, h->base_H, h->lif, h->lf_H, h->rf_H, h->lf_tw, h->rf_tw, h->com_pos_des,
                           h->com_vel_des, h->com_meas, h->neck_R, h->dq, h->lf_err, h->rf_err, h->qp_status, h->nwsr, sv));
    TRY(wc_glue_integrate_joints_dev(h->glue, h->dq, h->q_des, sv));
    loop_advance_kernel<<<1, 1, 0, s>>>(h->tick_dev);
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

} // namespace

extern "C" {

void wc_loop_default_config(wc_loop_config* c)
{
    if (!c) return;
    c->dcm_noise_sigma = 0.002; // measured DCM = reference + N(0, (2 mm)^2)
    c->noise_seed = 0x5EED;
    c->zmp_meas_offset = 0.001;
    c->use_graph = 1;
    c->reserved = 0;
}

wc_status wc_loop_create(wc_mpc* mpc, wc_ik* ik, wc_glue* glue, wc_traj* traj, const wc_loop_config* cfg, int32_t device, wc_loop** out)
{
    if (!mpc || !ik || !glue || !traj || !cfg || !out) return WC_ERR_INVALID_ARG;
    *out = nullptr;
    const int B = wc_mpc_batch(mpc);
    if (wc_ik_batch(ik) != B) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return WC_ERR_CUDA;
    wc_loop* h = new (std::nothrow) wc_loop();
    if (!h) return WC_ERR_ALLOC;
    h->mpc = mpc; h->ik = ik; h->glue = glue; h->traj = traj; h->cfg = *cfg;
    h->B = B; h->N = wc_mpc_horizon(mpc); h->nJ = wc_ik_dofs(ik); h->device = device;
    h->tick = 0; h->launches = 0; h->graph = nullptr; h->first = true;
    h->own = nullptr; h->ev_in = nullptr; h->ev_out = nullptr; h->side = nullptr; h->ev_fork = nullptr; h->ev_join = nullptr;
    if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->own, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming) != cudaSuccess) { delete h; return WC_ERR_CUDA; }
    const size_t Bz = (size_t)B, nJ = (size_t)h->nJ, N = (size_t)h->N;
    // doubles
    const size_t nd = (12 + 12 + 6 + 6 + 1 + 1 + 2 * (N + 1) + 2 * (N + 1) + 2 + 12 + 9 + 3 + 2 + 2 + 3 + 3 + nJ + 6 + 6 + nJ + 2) * Bz;
    const size_t ni = 6 * Bz + 2;          // int32
    const size_t nb = 4 * Bz;              // bytes
    char* base = nullptr;
    if (cudaMalloc(&base, nd * 8 + ni * 4 + nb) != cudaSuccess) { delete h; return WC_ERR_ALLOC; }
    cudaMemset(base, 0, nd * 8 + ni * 4 + nb);
    h->buf = reinterpret_cast<double*>(base);
    double* p = h->buf;
    auto take = [&](size_t rows) { double* q = p; p += rows * Bz; return q; };
    h->lf_H = take(12); h->rf_H = take(12); h->lf_tw = take(6); h->rf_tw = take(6); h->com_h = take(1); h->com_hv = take(1);
    h->ref_pos = take(2 * (N + 1)); h->ref_vel = take(2 * (N + 1)); h->dcm_meas = take(2); h->base_H = take(12); h->neck_R = take(9);
    h->com_meas = take(3); h->zmp = take(2); h->zmp_meas = take(2); h->com_pos_des = take(3); h->com_vel_des = take(3);
    h->dq = take(nJ); h->lf_err = take(6); h->rf_err = take(6); h->q_des = take(nJ); h->push_dxy = take(2);
    int32_t* pi = reinterpret_cast<int32_t*>(p);
    h->osqp_status = pi; h->iters = pi + Bz; h->flags = pi + 2 * Bz; h->qp_status = pi + 3 * Bz; h->nwsr = pi + 4 * Bz;
    h->push_tick = pi + 5 * Bz; h->tick_dev = pi + 6 * Bz;
    uint8_t* pb = reinterpret_cast<uint8_t*>(pi + ni);
    h->l_contact = pb; h->r_contact = pb + Bz; h->lif = pb + 2 * Bz; h->reset_all = pb + 3 * Bz;
    loop_fill_kernel<<<(unsigned)((B + 127) / 128), 128>>>(B, h->reset_all, 1);
    cudaMemset(h->push_tick, 0xFF, Bz * 4); // -1: no push
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(base); delete h; return WC_ERR_CUDA; }
    wc_mpc_set_tick_source_dev(mpc, h->tick_dev);
    wc_traj_set_tick_source_dev(traj, h->tick_dev);
    *out = h;
    return WC_OK;
}

void wc_loop_destroy(wc_loop* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    wc_mpc_set_tick_source_dev(h->mpc, nullptr);
    wc_traj_set_tick_source_dev(h->traj, nullptr);
    if (h->graph) cudaGraphExecDestroy(h->graph);
    if (h->own) cudaStreamDestroy(h->own);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    cudaFree(h->buf);
    delete h;
}

wc_status wc_loop_set_push_dev(wc_loop* h, const int32_t* push_tick, const double* push_dxy, void* stream)
{
    if (!h || !push_tick || !push_dxy) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t Bz = (size_t)h->B;
    if (cudaMemcpyAsync(h->push_tick, push_tick, Bz * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return WC_ERR_CUDA;
    if (cudaMemcpyAsync(h->push_dxy, push_dxy, 2 * Bz * 8, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return WC_ERR_CUDA;
    return WC_OK;
}

// WalkingModule::prepareRobot / startWalking side: tick counter to 0, the integrators at the initial posture q0 [nJ*B] and at
// its CoM (StableDCMModel::reset, WalkingZMPController::reset, WalkingModule.cpp:470-471), MPC forgotten (reset()).
wc_status wc_loop_reset_dev(wc_loop* h, const double* q0, void* stream)
{
    if (!h || !q0) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    void* sv = stream;
    const size_t Bz = (size_t)h->B;
    if (cudaStreamSynchronize(s) != cudaSuccess) return WC_ERR_CUDA;
    TRY(wc_mpc_reset(h->mpc, nullptr));
    if (cudaMemsetAsync(h->tick_dev, 0, 4, s) != cudaSuccess) return WC_ERR_CUDA;
    if (cudaMemcpyAsync(h->q_des, q0, (size_t)h->nJ * Bz * 8, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return WC_ERR_CUDA;
    // CoM of the initial posture about the stance foot of tick 0
    wc_traj_feet f{h->lf_H, h->rf_H, nullptr, nullptr, nullptr, nullptr, h->lif, nullptr, nullptr};
    TRY(wc_traj_feet_dev(h->traj, 0, &f, sv));
    loop_pre_kernel<<<(unsigned)((h->B + 127) / 128), 128, 0, s>>>(h->B, h->lf_H, h->rf_H, h->lif, h->base_H, h->neck_R);
    TRY(wc_fk_dev(h->ik, h->q_des, h->base_H, h->lif, nullptr, nullptr, nullptr, h->com_meas, nullptr, nullptr, nullptr, nullptr, sv));
    TRY(wc_glue_reset_dev(h->glue, h->com_meas, h->q_des, sv));
    h->tick = 0; h->first = true;
    h->launches += 3;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

static wc_status fail_cuda(const char* where, cudaError_t e)
{
    std::fprintf(stderr, "[wc_loop_run_dev] %s: %s\n", where, cudaGetErrorString(e));
    return WC_ERR_CUDA;
}

wc_status wc_loop_run_dev(wc_loop* h, int32_t n_ticks, void* stream)
{
    if (!h || n_ticks < 0) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    // the legacy default stream cannot be captured: graph ticks then run on the handle's own stream, ordered after the
    // work already on the caller's stream and before whatever the caller enqueues next
    const bool legacy = user == nullptr || user == cudaStreamLegacy;
    const bool detour = legacy && h->cfg.use_graph;
    cudaStream_t s = detour ? h->own : user;
    cudaError_t e;
    if (detour) {
        if ((e = cudaEventRecord(h->ev_in, user)) != cudaSuccess) return fail_cuda("event record", e);
        if ((e = cudaStreamWaitEvent(s, h->ev_in, 0)) != cudaSuccess) return fail_cuda("stream wait", e);
    }
    for (int k = 0; k < n_ticks; k++) {
        if (h->first || !h->cfg.use_graph) { // the first tick passes resetTrajectory = 1 (ring re-fill): launched directly
            TRY(enqueue_tick(h, h->first, s));
            h->first = false;
        } else {
            if (!h->graph) {
                cudaGraph_t g = nullptr;
                if ((e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal)) != cudaSuccess) return fail_cuda("begin capture", e);
                const wc_status rc = enqueue_tick(h, false, s);
                e = cudaStreamEndCapture(s, &g);
                if (rc != WC_OK || e != cudaSuccess || !g) {
                    if (g) cudaGraphDestroy(g);
                    return rc != WC_OK ? rc : fail_cuda("end capture", e);
                }
                e = cudaGraphInstantiate(&h->graph, g, 0);
                cudaGraphDestroy(g);
                if (e != cudaSuccess) { h->graph = nullptr; return fail_cuda("graph instantiate", e); }
            }
            if ((e = cudaGraphLaunch(h->graph, s)) != cudaSuccess) return fail_cuda("graph launch", e);
        }
        h->tick += 1;
        h->launches += 1;
    }
    if (detour) {
        if ((e = cudaEventRecord(h->ev_out, s)) != cudaSuccess) return fail_cuda("event record", e);
        if ((e = cudaStreamWaitEvent(user, h->ev_out, 0)) != cudaSuccess) return fail_cuda("stream wait", e);
    }
    return WC_OK;
}

int64_t wc_loop_tick_count(const wc_loop* h) { return h ? h->tick : 0; }

wc_status wc_loop_view_get(wc_loop* h, wc_loop_view* v)
{
    if (!h || !v) return WC_ERR_INVALID_ARG;
    v->zmp = h->zmp; v->dq = h->dq; v->q_des = h->q_des; v->com_pos_des = h->com_pos_des; v->com_vel_des = h->com_vel_des;
    v->com_meas = h->com_meas; v->dcm_meas = h->dcm_meas; v->lf_err = h->lf_err; v->rf_err = h->rf_err;
    v->osqp_status = h->osqp_status; v->iters = h->iters; v->flags = h->flags; v->qp_status = h->qp_status; v->nwsr = h->nwsr;
    v->lf_H = h->lf_H; v->rf_H = h->rf_H; v->ref_pos = h->ref_pos;
    v->lf_twist = h->lf_tw; v->rf_twist = h->rf_tw; v->base_H = h->base_H; v->neck_R_des = h->neck_R;
    v->l_contact = h->l_contact; v->r_contact = h->r_contact; v->left_is_fixed = h->lif;
    return WC_OK;
}

} // extern "C"
