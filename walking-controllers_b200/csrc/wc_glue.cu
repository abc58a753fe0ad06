// wc_glue.cu — the elementwise closed-loop glue between the MPC block and the QP-IK block of one control
// tick, batched, so that multi-tick Monte-Carlo rollouts need no host round trip:
//   StableDCMModel::{setInput,integrateModel}        modules/Walking_module/src/StableDCMModel.cpp:58-90
//   WalkingZMPController::{setFeedback,setReferenceSignal,evaluateControl}
//                                                     modules/Walking_module/src/WalkingZMPController.cpp:126-172
//   joint-velocity integration of updateModule        modules/Walking_module/src/WalkingModule.cpp:741-744
//   iCub::ctrl::Integrator (external, ctrlLib): Tustin rule y_k = y_{k-1} + Ts/2 (x_k + x_{k-1}), reset(y0): x_old = 0
// and the alternative DCM controller and the logger row:
//   WalkingDCMReactiveController::evaluateControl     modules/Walking_module/src/WalkingDCMReactiveController.cpp:63-81
//   WalkingModule::evaluateZMP                        modules/Walking_module/src/WalkingModule.cpp:824-876
//   WalkingFK::evaluateDCM                            modules/Walking_module/src/WalkingForwardKinematics.cpp:312-340
//   the 53 columns of m_walkingLogger->sendData       modules/Walking_module/src/WalkingModule.cpp:793-803, 1231-1250
//   WalkingZMPController::setPhase (gain scheduling) modules/Walking_module/src/WalkingZMPController.cpp:107-126
//   iCub::ctrl::minJerkTrajGen (external, ctrlLib): F(s) = -a / (s^3 - c s^2 - b s - a), Tustin-discretised (wc_capi.h)
// One thread per instance; every array is field-major SoA fp64; the kernels are HBM-streaming.
#include <cmath>
#include <cstdio>
#include <new>

#include <cuda_runtime.h>

#include "../../include/wc_capi.h"

struct wc_glue {
    wc_glue_config cfg;
    int B, device;
    double omega;
    double* lipm_c;   // [2][B] StableDCMModel::m_comPosition (= integrator output)
    double* lipm_v;   // [2][B] integrator x_old
    double* zc_y;     // [2][B] WalkingZMPController::m_velocityIntegral output
    double* zc_v;     // [2][B] its x_old
    double* q_y;      // [nJ][B] joint integrator output (m_qDesired)
    double* q_v;      // [nJ][B] its x_old
    double* cf;       // [12][B] WalkingFK's CoM filters: position y, u_old; velocity y, u_old (3 rows each)
    double* gs;       // [14][B] gain scheduling: (y1..y3, u1..u3) of the kZMP and of the kCoM smoother, then the current kZMP, kCoM
    double fnum[4], fden[4]; // minJerkTrajGen position filter: y a0 = sum b_k u_{n-k} - sum a_k y_{n-k}
    int64_t launches;
};

namespace {

__global__ void glue_reset_kernel(int B, int nJ, const double* com_xy, const double* q, double* lipm_c, double* lipm_v,
                                  double* zc_y, double* zc_v, double* q_y, double* q_v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    for (int a = 0; a < 2; a++) {
        if (com_xy) { lipm_c[a * Bz + i] = com_xy[a * Bz + i]; zc_y[a * Bz + i] = com_xy[a * Bz + i]; }
        if (com_xy) { lipm_v[a * Bz + i] = 0.0; zc_v[a * Bz + i] = 0.0; }
    }
    if (q)
        for (int j = 0; j < nJ; j++) { q_y[j * Bz + i] = q[j * Bz + i]; q_v[j * Bz + i] = 0.0; }
}

struct GsCoef { double b[4], a[4]; };

// WalkingZMPController::setPhase: one step of both smoothers towards the gains of the phase
__global__ void glue_set_phase_kernel(int B, GsCoef f, double kz_stance, double kc_stance, double kz_walk, double kc_walk,
                                      const double* dcm_vel_des, const uint8_t* stance, double* gs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    bool st;
    if (stance) st = stance[i] != 0;
    else { const double vx = dcm_vel_des[i], vy = dcm_vel_des[Bz + i]; st = sqrt(vx * vx + vy * vy) < 0.001; }
#pragma unroll
    for (int g = 0; g < 2; g++) {
        const double u = g == 0 ? (st ? kz_stance : kz_walk) : (st ? kc_stance : kc_walk);
        double* s = gs + (size_t)(6 * g) * Bz + i;
        const double y1 = s[0], y2 = s[Bz], y3 = s[2 * Bz], u1 = s[3 * Bz], u2 = s[4 * Bz], u3 = s[5 * Bz];
        double y = f.b[0] * u;
        y += f.b[1] * u1; y += f.b[2] * u2; y += f.b[3] * u3;
        y -= f.a[1] * y1; y -= f.a[2] * y2; y -= f.a[3] * y3;
        y = (1.0 / f.a[0]) * y;
        s[0] = y; s[Bz] = y1; s[2 * Bz] = y2; s[3 * Bz] = u; s[4 * Bz] = u1; s[5 * Bz] = u2;
        gs[(size_t)(12 + g) * Bz + i] = y;
    }
}

__global__ void glue_gs_init_kernel(int B, double kz, double kc, double* gs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    for (int r = 0; r < 6; r++) { gs[(size_t)r * Bz + i] = kz; gs[(size_t)(6 + r) * Bz + i] = kc; }
    gs[(size_t)12 * Bz + i] = kz; gs[(size_t)13 * Bz + i] = kc;
}

// LIPM reference + ZMP-CoM control law (WalkingModule.cpp:578-597, 657-695)
__global__ void glue_com_reference_kernel(int B, double omega, double half_dt, double k_zmp_c, double k_com_c,
                                          const double* dcm_des, const double* zmp_des, const double* zmp_meas,
                                          const double* com_meas, const double* com_z_des, const double* com_z_vel_des,
                                          const double* k_zmp, const double* k_com, double* lipm_c, double* lipm_v,
                                          double* zc_y, double* zc_v, double* com_pos_des, double* com_vel_des)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    const double kz = k_zmp ? k_zmp[i] : k_zmp_c, kc = k_com ? k_com[i] : k_com_c;
#pragma unroll
    for (int a = 0; a < 2; a++) {
        // StableDCMModel::integrateModel: dc = -omega (c - xi_des), c <- Tustin(dc)
        const double c_old = lipm_c[a * Bz + i];
        const double v = -omega * (c_old - dcm_des[a * Bz + i]);
        const double c_new = c_old + (v + lipm_v[a * Bz + i]) * half_dt;
        lipm_c[a * Bz + i] = c_new; lipm_v[a * Bz + i] = v;
        // WalkingZMPController::evaluateControl: dc* = kCoM (c_des - c) - kZMP (z_des - z) + dc_des, c* <- Tustin(dc*)
        const double vd = kc * (c_new - com_meas[a * Bz + i]) - kz * (zmp_des[a * Bz + i] - zmp_meas[a * Bz + i]) + v;
        const double y_new = zc_y[a * Bz + i] + (vd + zc_v[a * Bz + i]) * half_dt;
        zc_y[a * Bz + i] = y_new; zc_v[a * Bz + i] = vd;
        com_pos_des[a * Bz + i] = y_new;
        com_vel_des[a * Bz + i] = vd;
    }
    com_pos_des[2 * Bz + i] = com_z_des ? com_z_des[i] : 0.0;       // m_comHeightTrajectory.front()
    com_vel_des[2 * Bz + i] = com_z_vel_des ? com_z_vel_des[i] : 0.0; // m_comHeightVelocity.front()
}

__global__ void glue_integrate_joints_kernel(size_t n, double half_dt, const double* dq, double* q_y, double* q_v, double* q_des)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double v = dq[k];
    const double y = q_y[k] + (v + q_v[k]) * half_dt;
    q_y[k] = y; q_v[k] = v;
    if (q_des) q_des[k] = y;
}

// WalkingDCMReactiveController::evaluateControl: u = xi_des - dxi_des / omega - kDCM (xi_des - xi)
__global__ void glue_reactive_dcm_kernel(int B, double omega, double k_c, const double* k_dcm, const double* dcm_des,
                                         const double* dcm_vel_des, const double* dcm_meas, double* zmp_des)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    const double k = k_dcm ? k_dcm[i] : k_c;
#pragma unroll
    for (int a = 0; a < 2; a++) {
        const double xd = dcm_des[a * Bz + i];
        zmp_des[a * Bz + i] = xd - 1 / omega * dcm_vel_des[a * Bz + i] - k * (xd - dcm_meas[a * Bz + i]);
    }
}

// WalkingFK::evaluateDCM (unfiltered branch): the xy projection of com + com_vel / omega
__global__ void glue_evaluate_dcm_kernel(int B, double omega, const double* com, const double* com_vel, double* dcm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    dcm[i] = com[i] + com_vel[i] / omega;
    dcm[Bz + i] = com[Bz + i] + com_vel[Bz + i] / omega;
}

// WalkingModule::evaluateZMP: per-foot centre of pressure from the contact wrench (force, torque in the foot frame),
// moved to the world frame and averaged with the normal forces as weights.  ok = 0 where the reference returns false
// (total normal force < 0.1 N); the output of such an instance is left at 0.
__global__ void glue_evaluate_zmp_kernel(int B, const double* lw, const double* rw, const double* lf_H, const double* rf_H,
                                         double* zmp, int32_t* ok)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    const double fzl = lw[2 * Bz + i], fzr = rw[2 * Bz + i];
    double out[2] = {0.0, 0.0};
    const double totalZ = fzr + fzl;
    const bool good = !(totalZ < 0.1);
    if (good) {
        double w[2][2] = {{0.0, 0.0}, {0.0, 0.0}}; // world xy of the left / right local ZMP (undefined -> local origin)
        double defd[2];
        for (int f = 0; f < 2; f++) {
            const double* W = f == 0 ? lw : rw;
            const double* H = f == 0 ? lf_H : rf_H;
            const double fz = f == 0 ? fzl : fzr;
            double px = 0.0, py = 0.0;
            defd[f] = 0.0;
            if (!(fz < 0.001)) {
                px = -W[4 * Bz + i] / fz;
                py = W[3 * Bz + i] / fz;
                defd[f] = 1.0;
            }
            for (int r = 0; r < 2; r++) // Transform * Position with z = 0
                w[f][r] = H[(4 * r + 0) * Bz + i] * px + H[(4 * r + 1) * Bz + i] * py + H[(4 * r + 3) * Bz + i];
        }
        for (int r = 0; r < 2; r++)
            out[r] = ((fzl * defd[0]) / totalZ) * w[0][r] + ((fzr * defd[1]) / totalZ) * w[1][r];
    }
    zmp[i] = out[0]; zmp[Bz + i] = out[1];
    if (ok) ok[i] = good ? 1 : 0;
}

// iDynTree::Rotation::asRPY (external; restated from its documented convention R = Rz(y) Ry(p) Rx(r))
__device__ __forceinline__ void rot_as_rpy(const double* H, size_t Bz, int i, double* rpy)
{
    const double r20 = H[8 * Bz + i];
    if (r20 < 1.0) {
        if (r20 > -1.0) {
            rpy[0] = atan2(H[9 * Bz + i], H[10 * Bz + i]);
            rpy[1] = asin(-r20);
            rpy[2] = atan2(H[4 * Bz + i], H[0 * Bz + i]);
        } else { // not unique: yaw - roll = atan2(-R12, R11)
            rpy[0] = 0.0; rpy[1] = M_PI / 2.0; rpy[2] = -atan2(-H[6 * Bz + i], H[5 * Bz + i]);
        }
    } else { // not unique: yaw + roll = atan2(-R12, R11)
        rpy[0] = 0.0; rpy[1] = -M_PI / 2.0; rpy[2] = atan2(-H[6 * Bz + i], H[5 * Bz + i]);
    }
}

// one logger row per instance, column-major [WC_LOG_COLUMNS][B] in the order of startRecord (WalkingModule.cpp:1231-1250)
__global__ void glue_log_pack_kernel(int B, wc_log_inputs in, double* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    int c = 0;
    auto put = [&](double v) { out[(size_t)c * Bz + i] = v; c++; };
    auto put_n = [&](const double* p, int n) { for (int k = 0; k < n; k++) put(p ? p[k * Bz + i] : 0.0); };
    auto put_pose = [&](const double* H) {
        if (!H) { for (int k = 0; k < 6; k++) put(0.0); return; }
        put(H[3 * Bz + i]); put(H[7 * Bz + i]); put(H[11 * Bz + i]);
        double rpy[3]; rot_as_rpy(H, Bz, i, rpy);
        put(rpy[0]); put(rpy[1]); put(rpy[2]);
    };
    put_n(in.dcm, 2); put_n(in.dcm_des, 2); put_n(in.dcm_vel_des, 2);
    put_n(in.zmp, 2); put_n(in.zmp_des, 2); put_n(in.com, 3);
    put_n(in.com_des_xy, 2); put_n(in.com_vel_des_xy, 2);
    put_pose(in.lf_H); put_pose(in.rf_H); put_pose(in.lf_H_des); put_pose(in.rf_H_des);
    put_n(in.lf_err, 6); put_n(in.rf_err, 6);
}

// WalkingFK::evaluateCoM + evaluateDCM (WalkingForwardKinematics.cpp:278-340): CoM velocity = J_com [0 ; dq] (the base is the
// fixed stance foot: setInternalRobotState passes a zero base twist, :258-276), both signals through
// iCub::ctrl::FirstOrderLowPassFilter (external ctrlLib, restated: bilinear transform of 1 / (1 + s / (2 pi fc)),
// y_k = (Ts (u_k + u_{k-1}) - (Ts - 2 tau) y_{k-1}) / (2 tau + Ts), tau = 1 / (2 pi fc)), DCM = xy of com + com_vel / omega
// from the filtered or the raw pair (m_useFilters).
__global__ void glue_evaluate_com_kernel(int B, int nJ, double omega, double fa, double fb, int filt, int use_filters,
                                         const double* com, const double* J_com, const double* dq, double* cf,
                                         double* com_out, double* vel_out, double* dcm_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    const int nV = nJ + 6;
    double c[3], v[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 3; r++) c[r] = com[r * Bz + i];
    for (int j = 0; j < nJ; j++) {
        const double d = dq[j * Bz + i];
#pragma unroll
        for (int r = 0; r < 3; r++) v[r] += J_com[(size_t)(r * nV + 6 + j) * Bz + i] * d;
    }
    double cfil[3], vfil[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        cfil[r] = c[r]; vfil[r] = v[r];
        if (filt) {
            double* py = cf + (size_t)r * Bz + i; double* pu = cf + (size_t)(3 + r) * Bz + i;
            double* vy = cf + (size_t)(6 + r) * Bz + i; double* vu = cf + (size_t)(9 + r) * Bz + i;
            cfil[r] = (fb * (c[r] + *pu) - fa * *py); *py = cfil[r]; *pu = c[r];
            vfil[r] = (fb * (v[r] + *vu) - fa * *vy); *vy = vfil[r]; *vu = v[r];
        }
    }
    const double* co = (filt && use_filters) ? cfil : c;
    const double* vo = (filt && use_filters) ? vfil : v;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        if (com_out) com_out[r * Bz + i] = co[r];
        if (vel_out) vel_out[r * Bz + i] = vo[r];
    }
    if (dcm_out) { dcm_out[i] = co[0] + vo[0] / omega; dcm_out[Bz + i] = co[1] + vo[1] / omega; }
}

__global__ void glue_reset_com_filter_kernel(int B, double com_height, double* cf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const size_t Bz = (size_t)B;
    for (int r = 0; r < 12; r++) cf[(size_t)r * Bz + i] = 0.0;
    cf[2 * Bz + i] = com_height; cf[5 * Bz + i] = com_height; // y = u_old = (0, 0, com_height): Filter::init at steady state
}

} // namespace

extern "C" {

wc_status wc_glue_create(const wc_glue_config* cfg, int32_t B, int32_t device, wc_glue** out)
{
    if (!cfg || !out || B < 1 || cfg->n_dofs < 1 || cfg->n_dofs > WC_MAX_DOFS) return WC_ERR_INVALID_ARG;
    if (!(cfg->sampling_time > 0) || !(cfg->com_height > 0) || !(cfg->gravity > 0)) return WC_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return WC_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return WC_ERR_CUDA;
    wc_glue* h = new (std::nothrow) wc_glue();
    if (!h) return WC_ERR_ALLOC;
    h->cfg = *cfg; h->B = B; h->device = device; h->launches = 0;
    h->omega = std::sqrt(cfg->gravity / cfg->com_height);
    const size_t Bz = (size_t)B, nJ = (size_t)cfg->n_dofs;
    double* base = nullptr;
    const size_t total = (8 + 2 * nJ + 12 + 14) * Bz;
    if (cudaMalloc(&base, total * sizeof(double)) != cudaSuccess) { delete h; return WC_ERR_ALLOC; }
    cudaMemset(base, 0, total * sizeof(double));
    h->lipm_c = base; h->lipm_v = base + 2 * Bz; h->zc_y = base + 4 * Bz; h->zc_v = base + 6 * Bz;
    h->q_y = base + 8 * Bz; h->q_v = h->q_y + nJ * Bz; h->cf = h->q_v + nJ * Bz; h->gs = h->cf + 12 * Bz;
    glue_reset_com_filter_kernel<<<(B + 127) / 128, 128>>>(B, cfg->com_height, h->cf); // the filters' init of WalkingFK::initialize
    if (cfg->use_gain_scheduling) {
        if (!(cfg->smoothing_time > 0)) { cudaFree(base); delete h; return WC_ERR_INVALID_ARG; }
        // minJerkTrajGen(1, Ts, T): F(s) = -a / (s^3 - c s^2 - b s - a) with s = (2 / Ts)(z - 1)/(z + 1)
        const double T = cfg->smoothing_time, k = 2.0 / cfg->sampling_time;
        const double a = -150.765868956161 / (T * T * T), b = -84.9812819469538 / (T * T), c = -15.9669610709384 / T;
        const double k2 = k * k, k3 = k2 * k;
        h->fden[0] = k3 - c * k2 - b * k - a;  h->fden[1] = -3.0 * k3 + c * k2 - b * k - 3.0 * a;
        h->fden[2] = 3.0 * k3 + c * k2 + b * k - 3.0 * a;  h->fden[3] = -k3 - c * k2 + b * k - a;
        h->fnum[0] = -a; h->fnum[1] = -3.0 * a; h->fnum[2] = -3.0 * a; h->fnum[3] = -a;
        // m_kZMPSmoother->init(kZMP_stance), m_kCoMSmoother->init(kCoM_stance) (:78-83)
        glue_gs_init_kernel<<<(B + 127) / 128, 128>>>(B, cfg->k_zmp_stance, cfg->k_com_stance, h->gs);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(base); delete h; return WC_ERR_CUDA; }
    *out = h;
    return WC_OK;
}

void wc_glue_destroy(wc_glue* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->lipm_c);
    delete h;
}

int64_t wc_glue_launch_count(const wc_glue* h) { return h ? h->launches : 0; }

wc_status wc_glue_reset_dev(wc_glue* h, const double* com_xy, const double* q, void* stream)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    glue_reset_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        h->B, h->cfg.n_dofs, com_xy, q, h->lipm_c, h->lipm_v, h->zc_y, h->zc_v, h->q_y, h->q_v);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_glue_com_reference_dev(wc_glue* h, const double* dcm_des, const double* zmp_des, const double* zmp_meas,
                                    const double* com_meas, const double* com_z_des, const double* com_z_vel_des,
                                    const double* k_zmp, const double* k_com, double* com_pos_des, double* com_vel_des,
                                    void* stream)
{
    if (!h || !dcm_des || !zmp_des || !zmp_meas || !com_meas || !com_pos_des || !com_vel_des) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    // gains: the caller's per-instance arrays, else the scheduled ones (setPhase), else the walking constants
    const size_t Bz = (size_t)h->B;
    if (!k_zmp && h->cfg.use_gain_scheduling) k_zmp = h->gs + 12 * Bz;
    if (!k_com && h->cfg.use_gain_scheduling) k_com = h->gs + 13 * Bz;
    glue_com_reference_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        h->B, h->omega, 0.5 * h->cfg.sampling_time, h->cfg.k_zmp, h->cfg.k_com, dcm_des, zmp_des, zmp_meas, com_meas,
        com_z_des, com_z_vel_des, k_zmp, k_com, h->lipm_c, h->lipm_v, h->zc_y, h->zc_v, com_pos_des, com_vel_des);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_glue_set_phase_dev(wc_glue* h, const double* dcm_vel_des, const uint8_t* stance, void* stream)
{
    if (!h || (!dcm_vel_des && !stance)) return WC_ERR_INVALID_ARG;
    if (!h->cfg.use_gain_scheduling) return WC_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    GsCoef f;
    for (int k = 0; k < 4; k++) { f.b[k] = h->fnum[k]; f.a[k] = h->fden[k]; }
    glue_set_phase_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        h->B, f, h->cfg.k_zmp_stance, h->cfg.k_com_stance, h->cfg.k_zmp, h->cfg.k_com, dcm_vel_des, stance, h->gs);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_glue_integrate_joints_dev(wc_glue* h, const double* dq, double* q_des, void* stream)
{
    if (!h || !dq) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t n = (size_t)h->cfg.n_dofs * (size_t)h->B;
    glue_integrate_joints_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, 0.5 * h->cfg.sampling_time, dq, h->q_y, h->q_v, q_des);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_glue_reactive_dcm_dev(wc_glue* h, double k_dcm_const, const double* k_dcm, const double* dcm_des,
                                   const double* dcm_vel_des, const double* dcm_meas, double* zmp_des, void* stream)
{
    if (!h || !dcm_des || !dcm_vel_des || !dcm_meas || !zmp_des) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    glue_reactive_dcm_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        h->B, h->omega, k_dcm_const, k_dcm, dcm_des, dcm_vel_des, dcm_meas, zmp_des);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_glue_evaluate_dcm_dev(wc_glue* h, const double* com, const double* com_vel, double* dcm, void* stream)
{
    if (!h || !com || !com_vel || !dcm) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    glue_evaluate_dcm_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->B, h->omega, com, com_vel, dcm);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_glue_evaluate_zmp_dev(wc_glue* h, const double* l_wrench, const double* r_wrench, const double* lf_H,
                                   const double* rf_H, double* zmp, int32_t* ok, void* stream)
{
    if (!h || !l_wrench || !r_wrench || !lf_H || !rf_H || !zmp) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    glue_evaluate_zmp_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        h->B, l_wrench, r_wrench, lf_H, rf_H, zmp, ok);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_glue_log_pack_dev(wc_glue* h, const wc_log_inputs* in, double* rows, void* stream)
{
    if (!h || !in || !rows) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    glue_log_pack_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->B, *in, rows);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

// WalkingDCMReactiveController for the host classes: a batch of one (or a few) instances with host pointers
wc_status wc_glue_reactive_dcm_host(wc_glue* h, double k_dcm_const, const double* k_dcm, const double* dcm_des,
                                    const double* dcm_vel_des, const double* dcm_meas, double* zmp_des)
{
    if (!h || !dcm_des || !dcm_vel_des || !dcm_meas || !zmp_des) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t Bz = (size_t)h->B;
    double* d = nullptr;
    if (cudaMalloc(&d, 9 * Bz * sizeof(double)) != cudaSuccess) return WC_ERR_ALLOC;
    cudaMemcpy(d, dcm_des, 2 * Bz * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d + 2 * Bz, dcm_vel_des, 2 * Bz * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d + 4 * Bz, dcm_meas, 2 * Bz * sizeof(double), cudaMemcpyHostToDevice);
    if (k_dcm) cudaMemcpy(d + 8 * Bz, k_dcm, Bz * sizeof(double), cudaMemcpyHostToDevice);
    wc_status st = wc_glue_reactive_dcm_dev(h, k_dcm_const, k_dcm ? d + 8 * Bz : nullptr, d, d + 2 * Bz, d + 4 * Bz, d + 6 * Bz, nullptr);
    if (st == WC_OK && cudaMemcpy(zmp_des, d + 6 * Bz, 2 * Bz * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) st = WC_ERR_CUDA;
    cudaFree(d);
    return st;
}

wc_status wc_glue_reset_com_filter_dev(wc_glue* h, void* stream)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    glue_reset_com_filter_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->B, h->cfg.com_height, h->cf);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_glue_evaluate_com_dev(wc_glue* h, const double* com, const double* J_com, const double* dq, double cut_frequency,
                                   int32_t use_filters, double* com_out, double* com_vel_out, double* dcm_out, void* stream)
{
    if (!h || !com || !J_com || !dq) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const double Ts = h->cfg.sampling_time;
    const int filt = cut_frequency > 0.0;
    double fa = 0.0, fb = 0.0;
    if (filt) {
        const double tau = 1.0 / (2.0 * 3.14159265358979323846 * cut_frequency);
        fa = (Ts - 2.0 * tau) / (2.0 * tau + Ts); fb = Ts / (2.0 * tau + Ts);
    }
    glue_evaluate_com_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        h->B, h->cfg.n_dofs, h->omega, fa, fb, filt, use_filters, com, J_com, dq, h->cf, com_out, com_vel_out, dcm_out);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_glue_evaluate_com_host(wc_glue* h, const double* com, const double* J_com, const double* dq, double cut_frequency,
                                    int32_t use_filters, double* com_out, double* com_vel_out, double* dcm_out)
{
    if (!h || !com || !J_com || !dq) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t Bz = (size_t)h->B, nJ = (size_t)h->cfg.n_dofs, nV = nJ + 6;
    const size_t n_in = (3 + 3 * nV + nJ) * Bz;
    double* d = nullptr;
    if (cudaMalloc(&d, (n_in + 8 * Bz) * sizeof(double)) != cudaSuccess) return WC_ERR_ALLOC;
    double* d_com = d; double* d_J = d + 3 * Bz; double* d_dq = d_J + 3 * nV * Bz; double* d_o = d + n_in;
    cudaMemcpy(d_com, com, 3 * Bz * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_J, J_com, 3 * nV * Bz * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_dq, dq, nJ * Bz * sizeof(double), cudaMemcpyHostToDevice);
    wc_status st = wc_glue_evaluate_com_dev(h, d_com, d_J, d_dq, cut_frequency, use_filters, d_o, d_o + 3 * Bz, d_o + 6 * Bz, nullptr);
    auto back = [&](double* dst, const double* src, size_t n) { if (dst && st == WC_OK && cudaMemcpy(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) st = WC_ERR_CUDA; };
    back(com_out, d_o, 3 * Bz); back(com_vel_out, d_o + 3 * Bz, 3 * Bz); back(dcm_out, d_o + 6 * Bz, 2 * Bz);
    cudaFree(d);
    return st;
}

} // extern "C"
