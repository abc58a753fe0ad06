// wc_glue.cu — the elementwise closed-loop glue between the MPC block and the QP-IK block of one control
// tick, batched, so that multi-tick Monte-Carlo rollouts need no host round trip:
//   StableDCMModel::{setInput,integrateModel}        modules/Walking_module/src/StableDCMModel.cpp:58-90
//   WalkingZMPController::{setFeedback,setReferenceSignal,evaluateControl}
//                                                     modules/Walking_module/src/WalkingZMPController.cpp:126-172
//   joint-velocity integration of updateModule        modules/Walking_module/src/WalkingModule.cpp:741-744
//   iCub::ctrl::Integrator (external, ctrlLib): Tustin rule y_k = y_{k-1} + Ts/2 (x_k + x_{k-1}), reset(y0): x_old = 0
// The gain scheduling of WalkingZMPController::setPhase (iCub::ctrl::minJerkTrajGen, external) is not restated:
// gains are inputs (constants from the configuration, or per-instance arrays the caller schedules).
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
    const size_t total = (8 + 2 * nJ) * Bz;
    if (cudaMalloc(&base, total * sizeof(double)) != cudaSuccess) { delete h; return WC_ERR_ALLOC; }
    cudaMemset(base, 0, total * sizeof(double));
    h->lipm_c = base; h->lipm_v = base + 2 * Bz; h->zc_y = base + 4 * Bz; h->zc_v = base + 6 * Bz;
    h->q_y = base + 8 * Bz; h->q_v = h->q_y + nJ * Bz;
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
    glue_com_reference_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        h->B, h->omega, 0.5 * h->cfg.sampling_time, h->cfg.k_zmp, h->cfg.k_com, dcm_des, zmp_des, zmp_meas, com_meas,
        com_z_des, com_z_vel_des, k_zmp, k_com, h->lipm_c, h->lipm_v, h->zc_y, h->zc_v, com_pos_des, com_vel_des);
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

} // extern "C"
