// wc_traj.cu — on-device producer of the hot path's per-tick inputs: what
// TrajectoryGenerator hands WalkingModule each tick
//   getDCMPositionTrajectory / getDCMVelocityTrajectory, getFeetTrajectories, getFeetTwist, getWhenUseLeftAsFixed,
//   getFeetStandingPeriods, getCoMHeightTrajectory / getCoMHeightVelocity
//                                                   modules/Walking_module/src/TrajectoryGenerator.cpp:501-598
// for a batch of periodic gaits, so that Monte-Carlo rollouts need no host generator and no H2D traffic.
// The reference's producer is the UnicyclePlanner library (external, not vendored: footstep planning towards a joypad
// goal, cubic-spline feet, DCM by backward recursion over the planned steps); it is NOT restated.  This producer is the
// synthetic workload of SURVEY.md §8(d): periodic walking with the shipped planner parameters
// (app/robots/iCubGazeboV2_5/plannerParams.ini: nominalWidth, nominalDuration, switchOverSwingRatio, stepHeight,
// left/rightZMPDelta, com_height) and per-instance phase, step length, heading and heading rate; the DCM reference is the
// analytic LIPM solution xi(t) = z_k + e^{omega (t - T)} (xi_end,k - z_k) with xi_end by backward recursion over the
// footsteps, the swing foot a cubic blend with a parabolic height profile.  Measured DCM = reference + Gaussian noise from
// Philox-4x32-10 keyed by (seed; instance, tick) — the disturbance of the Monte-Carlo sweeps.
// One thread per instance (tables, feet) or per (sample, instance) (DCM window); field-major SoA fp64; HBM-streaming.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <new>

#include <cuda_runtime.h>

#include "../../include/wc_capi.h"

struct wc_traj {
    wc_traj_config cfg;
    int B, device, S2; // S2 = table rows (footsteps k = -2 .. n_steps-1)
    double omega;
    double* tbl;       // [7][S2][B]: foot_x, foot_y, foot_yaw, zmp_x, zmp_y, xi_end_x, xi_end_y
    int32_t* phase;    // [B]
    uint8_t* first_left; // [B]
    int64_t launches;
    bool gait_set;
    const int32_t* tick_dev; // optional device-side tick offset added to every tick / first_sample argument
};

namespace {

enum { T_FX = 0, T_FY, T_FYAW, T_ZX, T_ZY, T_EX, T_EY, T_COUNT };

struct TrajDev {
    int B, S2, step_ticks, ds_ticks;
    double dt, omega, width, step_height, zmp_dx, zmp_dy, com_height;
    double* tbl; int32_t* phase; uint8_t* first_left;
    const int32_t* tick_dev;
};
__device__ __forceinline__ int tick_of(const TrajDev& d, int t) { return d.tick_dev ? t + *d.tick_dev : t; }

__device__ __forceinline__ double tb(const TrajDev& d, int f, int k, int i)
{
    int j = k + 2;
    j = j < 0 ? 0 : (j > d.S2 - 1 ? d.S2 - 1 : j);
    return d.tbl[((size_t)f * d.S2 + j) * (size_t)d.B + i];
}
__device__ __forceinline__ double sign_of(const TrajDev& d, int k, int i)
{
    const bool even = (k & 1) == 0;
    return (even == (d.first_left[i] != 0)) ? 1.0 : -1.0; // +1 = left foot
}
__device__ __forceinline__ int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

// footstep tables of one instance: midpoint path c_k = c_{k-1} + L (cos th_k, sin th_k), feet at -/+ width/2 across it,
// ZMP of step k = foot + R(yaw) (zmp_dx, -/+ zmp_dy), xi_end by the backward recursion xi_end(k) = xi_start(k+1)
__global__ void traj_tables_kernel(TrajDev d, const int32_t* phase, const double* step_len, const double* dtheta,
                                   const double* theta0, const uint8_t* first_left)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.B) return;
    const size_t B = (size_t)d.B;
    d.phase[i] = phase ? phase[i] : 0;
    d.first_left[i] = first_left ? first_left[i] : 1;
    const double L = step_len[i], dth = dtheta ? dtheta[i] : 0.0, th0 = theta0 ? theta0[i] : 0.0;
    double cx = 0.0, cy = 0.0;
    for (int j = 0; j < d.S2; j++) {
        const int k = j - 2;
        const double th = th0 + (double)k * dth;
        const double c = cos(th), s = sin(th);
        cx += L * c; cy += L * s;
        const double sg = sign_of(d, k, i);
        const double fx = cx - 0.5 * d.width * sg * s, fy = cy + 0.5 * d.width * sg * c;
        const double zl = -d.zmp_dy * sg;
        d.tbl[((size_t)T_FX * d.S2 + j) * B + i] = fx;
        d.tbl[((size_t)T_FY * d.S2 + j) * B + i] = fy;
        d.tbl[((size_t)T_FYAW * d.S2 + j) * B + i] = th;
        d.tbl[((size_t)T_ZX * d.S2 + j) * B + i] = fx + c * d.zmp_dx - s * zl;
        d.tbl[((size_t)T_ZY * d.S2 + j) * B + i] = fy + s * d.zmp_dx + c * zl;
    }
    const double decay = exp(-d.omega * d.step_ticks * d.dt);
    double ex = d.tbl[((size_t)T_ZX * d.S2 + d.S2 - 1) * B + i], ey = d.tbl[((size_t)T_ZY * d.S2 + d.S2 - 1) * B + i];
    d.tbl[((size_t)T_EX * d.S2 + d.S2 - 1) * B + i] = ex;
    d.tbl[((size_t)T_EY * d.S2 + d.S2 - 1) * B + i] = ey;
    for (int j = d.S2 - 2; j >= 0; j--) {
        const double zx = d.tbl[((size_t)T_ZX * d.S2 + j + 1) * B + i], zy = d.tbl[((size_t)T_ZY * d.S2 + j + 1) * B + i];
        ex = zx + decay * (ex - zx); ey = zy + decay * (ey - zy);
        d.tbl[((size_t)T_EX * d.S2 + j) * B + i] = ex;
        d.tbl[((size_t)T_EY * d.S2 + j) * B + i] = ey;
    }
}

__device__ __forceinline__ void dcm_at(const TrajDev& d, int s, int i, double* pos, double* vel)
{
    const int a = s + d.phase[i];
    const int k = floordiv(a, d.step_ticks);
    const double tau = (double)(a - k * d.step_ticks) * d.dt;
    const double e = exp(d.omega * (tau - d.step_ticks * d.dt));
    const double zx = tb(d, T_ZX, k, i), zy = tb(d, T_ZY, k, i);
    const double gx = tb(d, T_EX, k, i) - zx, gy = tb(d, T_EY, k, i) - zy;
    pos[0] = zx + e * gx; pos[1] = zy + e * gy;
    if (vel) { vel[0] = d.omega * (e * gx); vel[1] = d.omega * (e * gy); }
}

// samples s0 .. s0+n-1 of the DCM reference (and velocity): sample s, axis a of instance i at [(2 (s - s0) + a) B + i] —
// rows [0, 2 (N+1)) with s0 = t are the deque WalkingController::setReferenceSignal reads at tick t
__global__ void traj_dcm_kernel(TrajDev d, int s0_arg, int n, double* pos_out, double* vel_out)
{
    const int s0 = tick_of(d, s0_arg);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= d.B || r >= n) return;
    double p[2], v[2];
    dcm_at(d, s0 + r, i, p, vel_out ? v : nullptr);
    const size_t B = (size_t)d.B;
    if (pos_out) { pos_out[(size_t)(2 * r) * B + i] = p[0]; pos_out[(size_t)(2 * r + 1) * B + i] = p[1]; }
    if (vel_out) { vel_out[(size_t)(2 * r) * B + i] = v[0]; vel_out[(size_t)(2 * r + 1) * B + i] = v[1]; }
}

// Philox-4x32-10 (Salmon et al., SC'11), counter (instance, tick, 0, 0), key (seed lo, seed hi)
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__global__ void traj_measured_dcm_kernel(TrajDev d, int t_arg, double sigma, uint64_t seed, double* out)
{
    const int t = tick_of(d, t_arg);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.B) return;
    double p[2];
    dcm_at(d, t, i, p, nullptr);
    uint32_t c[4] = {(uint32_t)i, (uint32_t)t, 0u, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    // two 53-bit uniforms in (0, 1] and [0, 1) -> Box-Muller
    const double u1 = ((double)((((uint64_t)c[0] << 32) | c[1]) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
    const double u2 = (double)((((uint64_t)c[2] << 32) | c[3]) >> 11) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1)), ang = 6.283185307179586476925 * u2;
    out[i] = p[0] + sigma * (rad * cos(ang));
    out[(size_t)d.B + i] = p[1] + sigma * (rad * sin(ang));
}

__device__ __forceinline__ void put_H(double* H, size_t B, int i, double x, double y, double z, double yaw)
{
    const double c = cos(yaw), s = sin(yaw);
    const double v[12] = {c, -s, 0.0, x, s, c, 0.0, y, 0.0, 0.0, 1.0, z};
#pragma unroll
    for (int e = 0; e < 12; e++) H[(size_t)e * B + i] = v[e];
}

// feet poses, twists (mixed: linear then angular), contact flags and the fixed-foot flag at tick t
__global__ void traj_feet_kernel(TrajDev d, int t_arg, wc_traj_feet o)
{
    const int t = tick_of(d, t_arg);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.B) return;
    const size_t B = (size_t)d.B;
    const int a = t + d.phase[i];
    const int k = floordiv(a, d.step_ticks);
    const int ii = a - k * d.step_ticks;
    const bool in_ds = ii < d.ds_ticks;
    const double sx = tb(d, T_FX, k, i), sy = tb(d, T_FY, k, i), syaw = tb(d, T_FYAW, k, i);
    const double px = tb(d, T_FX, k - 1, i), py = tb(d, T_FY, k - 1, i), pyaw = tb(d, T_FYAW, k - 1, i);
    const double nx = tb(d, T_FX, k + 1, i), ny = tb(d, T_FY, k + 1, i), nyaw = tb(d, T_FYAW, k + 1, i);
    const double ss_len = (double)(d.step_ticks - d.ds_ticks);
    double u = (double)(ii - d.ds_ticks) / ss_len;
    u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
    const double blend = 3.0 * u * u - 2.0 * u * u * u;
    const double dblend = in_ds ? 0.0 : (6.0 * u - 6.0 * u * u) / (ss_len * d.dt);
    const double ox = px + blend * (nx - px), oy = py + blend * (ny - py), oyaw = pyaw + blend * (nyaw - pyaw);
    const double oz = in_ds ? 0.0 : 4.0 * d.step_height * u * (1.0 - u);
    const double tw[6] = {dblend * (nx - px), dblend * (ny - py), in_ds ? 0.0 : 4.0 * d.step_height * (1.0 - 2.0 * u) / (ss_len * d.dt),
                          0.0, 0.0, dblend * (nyaw - pyaw)};
    const bool stance_left = sign_of(d, k, i) > 0.0;
    double* Hs = stance_left ? o.lf_H : o.rf_H; double* Ho = stance_left ? o.rf_H : o.lf_H;
    if (Hs) put_H(Hs, B, i, sx, sy, 0.0, syaw);
    if (Ho) put_H(Ho, B, i, ox, oy, oz, oyaw);
    double* ts = stance_left ? o.lf_twist : o.rf_twist; double* to = stance_left ? o.rf_twist : o.lf_twist;
#pragma unroll
    for (int e = 0; e < 6; e++) {
        if (ts) ts[(size_t)e * B + i] = 0.0;
        if (to) to[(size_t)e * B + i] = tw[e];
    }
    if (o.l_contact) o.l_contact[i] = (stance_left || in_ds) ? 1 : 0;
    if (o.r_contact) o.r_contact[i] = (!stance_left || in_ds) ? 1 : 0;
    if (o.left_is_fixed) o.left_is_fixed[i] = stance_left ? 1 : 0;
    if (o.com_height) o.com_height[i] = d.com_height; // CoMHeightTrajectoryGenerator with comHeightDelta = 0
    if (o.com_height_vel) o.com_height_vel[i] = 0.0;
}

TrajDev dev_view(const wc_traj* h)
{
    TrajDev d;
    d.B = h->B; d.S2 = h->S2; d.step_ticks = h->cfg.step_ticks; d.ds_ticks = h->cfg.ds_ticks;
    d.dt = h->cfg.sampling_time; d.omega = h->omega; d.width = h->cfg.nominal_width; d.step_height = h->cfg.step_height;
    d.zmp_dx = h->cfg.zmp_delta[0]; d.zmp_dy = h->cfg.zmp_delta[1]; d.com_height = h->cfg.com_height;
    d.tbl = h->tbl; d.phase = h->phase; d.first_left = h->first_left;
    d.tick_dev = h->tick_dev;
    return d;
}

} // namespace

extern "C" {

wc_status wc_traj_create(const wc_traj_config* cfg, int32_t B, int32_t device, wc_traj** out)
{
    if (!cfg || !out || B < 1) return WC_ERR_INVALID_ARG;
    if (cfg->step_ticks < 2 || cfg->ds_ticks < 0 || cfg->ds_ticks >= cfg->step_ticks || cfg->n_steps < 1) return WC_ERR_INVALID_ARG;
    if (!(cfg->sampling_time > 0) || !(cfg->com_height > 0) || !(cfg->gravity > 0)) return WC_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return WC_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return WC_ERR_CUDA;
    wc_traj* h = new (std::nothrow) wc_traj();
    if (!h) return WC_ERR_ALLOC;
    h->cfg = *cfg; h->B = B; h->device = device; h->launches = 0; h->gait_set = false; h->tick_dev = nullptr;
    h->S2 = cfg->n_steps + 2;
    h->omega = std::sqrt(cfg->gravity / cfg->com_height);
    const size_t Bz = (size_t)B;
    if (cudaMalloc(&h->tbl, (size_t)T_COUNT * h->S2 * Bz * sizeof(double)) != cudaSuccess) { delete h; return WC_ERR_ALLOC; }
    if (cudaMalloc(&h->phase, Bz * sizeof(int32_t)) != cudaSuccess || cudaMalloc(&h->first_left, Bz) != cudaSuccess) {
        cudaFree(h->tbl); cudaFree(h->phase); delete h; return WC_ERR_ALLOC;
    }
    *out = h;
    return WC_OK;
}

void wc_traj_destroy(wc_traj* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->tbl); cudaFree(h->phase); cudaFree(h->first_left);
    delete h;
}

int64_t wc_traj_launch_count(const wc_traj* h) { return h ? h->launches : 0; }

wc_status wc_traj_set_gait_dev(wc_traj* h, const int32_t* phase, const double* step_len, const double* dtheta,
                               const double* theta0, const uint8_t* first_left, void* stream)
{
    if (!h || !step_len) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    traj_tables_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(dev_view(h), phase, step_len, dtheta, theta0, first_left);
    h->launches += 1; h->gait_set = true;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_traj_set_tick_source_dev(wc_traj* h, const int32_t* tick_dev)
{
    if (!h) return WC_ERR_INVALID_ARG;
    h->tick_dev = tick_dev;
    return WC_OK;
}

wc_status wc_traj_dcm_dev(wc_traj* h, int32_t first_sample, int32_t n_samples, double* dcm_pos, double* dcm_vel, void* stream)
{
    if (!h || n_samples < 1 || n_samples > 65535 || (!dcm_pos && !dcm_vel)) return WC_ERR_INVALID_ARG;
    if (!h->gait_set) return WC_ERR_STATE;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const dim3 grid((h->B + 127) / 128, n_samples);
    traj_dcm_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(dev_view(h), first_sample, n_samples, dcm_pos, dcm_vel);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_traj_measured_dcm_dev(wc_traj* h, int32_t tick, double sigma, uint64_t seed, double* dcm, void* stream)
{
    if (!h || !dcm) return WC_ERR_INVALID_ARG;
    if (!h->gait_set) return WC_ERR_STATE;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    traj_measured_dcm_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(dev_view(h), tick, sigma, seed, dcm);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_traj_feet_dev(wc_traj* h, int32_t tick, const wc_traj_feet* out, void* stream)
{
    if (!h || !out) return WC_ERR_INVALID_ARG;
    if (!h->gait_set) return WC_ERR_STATE;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    traj_feet_kernel<<<(h->B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(dev_view(h), tick, *out);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

// Host-pointer convenience for the drop-in class TrajectoryGenerator (small batches): gait parameters in, n_samples ticks of
// every signal out, sample-major: signal element e of sample s of instance i at [(s * E + e) * B + i].
wc_status wc_traj_generate_host(wc_traj* h, const int32_t* phase, const double* step_len, const double* dtheta, const double* theta0,
                                const uint8_t* first_left, int32_t first_sample, int32_t n_samples, double* dcm_pos, double* dcm_vel,
                                double* lf_H, double* rf_H, double* lf_twist, double* rf_twist, uint8_t* l_contact,
                                uint8_t* r_contact, uint8_t* left_is_fixed, double* com_height, double* com_height_vel)
{
    if (!h || !step_len || n_samples < 1) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t B = (size_t)h->B, n = (size_t)n_samples;
    const size_t nd = (4 + 2 * n * 2 + n * (12 + 12 + 6 + 6 + 2)) * B; // doubles
    const size_t nb = (4 + 1 + 3 * n) * B;                              // bytes (phase as 4)
    double* dd = nullptr; uint8_t* db = nullptr;
    if (cudaMalloc(&dd, nd * sizeof(double)) != cudaSuccess) return WC_ERR_ALLOC;
    if (cudaMalloc(&db, nb) != cudaSuccess) { cudaFree(dd); return WC_ERR_ALLOC; }
    double* d_len = dd; double* d_dth = dd + B; double* d_th0 = dd + 2 * B; double* p = dd + 4 * B;
    double* d_pos = p; p += 2 * n * B; double* d_vel = p; p += 2 * n * B;
    double* d_lf = p; p += 12 * n * B; double* d_rf = p; p += 12 * n * B; double* d_lt = p; p += 6 * n * B; double* d_rt = p; p += 6 * n * B;
    double* d_ch = p; p += n * B; double* d_cv = p;
    int32_t* d_phase = reinterpret_cast<int32_t*>(db); uint8_t* d_fl = db + 4 * B; uint8_t* d_lc = d_fl + B; uint8_t* d_rc = d_lc + n * B;
    uint8_t* d_lfix = d_rc + n * B;
    cudaMemcpy(d_len, step_len, B * sizeof(double), cudaMemcpyHostToDevice);
    if (dtheta) cudaMemcpy(d_dth, dtheta, B * sizeof(double), cudaMemcpyHostToDevice);
    if (theta0) cudaMemcpy(d_th0, theta0, B * sizeof(double), cudaMemcpyHostToDevice);
    if (phase) cudaMemcpy(d_phase, phase, B * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (first_left) cudaMemcpy(d_fl, first_left, B, cudaMemcpyHostToDevice);
    wc_status st = wc_traj_set_gait_dev(h, phase ? d_phase : nullptr, d_len, dtheta ? d_dth : nullptr, theta0 ? d_th0 : nullptr,
                                        first_left ? d_fl : nullptr, nullptr);
    if (st == WC_OK) st = wc_traj_dcm_dev(h, first_sample, n_samples, d_pos, d_vel, nullptr);
    for (size_t s = 0; s < n && st == WC_OK; s++) {
        wc_traj_feet f = {d_lf + 12 * s * B, d_rf + 12 * s * B, d_lt + 6 * s * B, d_rt + 6 * s * B, d_lc + s * B, d_rc + s * B,
                          d_lfix + s * B, d_ch + s * B, d_cv + s * B};
        st = wc_traj_feet_dev(h, first_sample + (int32_t)s, &f, nullptr);
    }
    auto back = [&](void* dst, const void* src, size_t bytes) { if (dst && st == WC_OK && cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) st = WC_ERR_CUDA; };
    back(dcm_pos, d_pos, 2 * n * B * sizeof(double)); back(dcm_vel, d_vel, 2 * n * B * sizeof(double));
    back(lf_H, d_lf, 12 * n * B * sizeof(double)); back(rf_H, d_rf, 12 * n * B * sizeof(double));
    back(lf_twist, d_lt, 6 * n * B * sizeof(double)); back(rf_twist, d_rt, 6 * n * B * sizeof(double));
    back(l_contact, d_lc, n * B); back(r_contact, d_rc, n * B); back(left_is_fixed, d_lfix, n * B);
    back(com_height, d_ch, n * B * sizeof(double)); back(com_height_vel, d_cv, n * B * sizeof(double));
    cudaFree(dd); cudaFree(db);
    return st;
}

} // extern "C"
