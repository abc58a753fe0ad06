// wc_mpc_api.cu — C ABI of the batched DCM-MPC (include/wc_capi.h): handle management,
// device state allocation, kernel launches, pinned host staging.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/wc_capi.h"
#include "host_pipe.cuh"
#include "mpc_kernels.cuh"

using namespace wcb;

struct wc_mpc {
    MpcDev d;
    wc_mpc_config cfg;
    int device;
    int lanes;
    int chain; // 2: split-chain path (mpc_split.cu), 1: chain-per-thread path (mpc_chain.cu), 0: cooperative path (mpc_kernels.cu)
    int64_t launches;
    int64_t tick;
    std::vector<void*> allocs;
    // staging for the *_host entry points
    HostPipe pipe; // double-buffered staging of the host-pointer entry points
    // split-chain path: osqp_setup + phase 1 of the instances with a new workspace run on this stream next to the phase 1 of
    // everyone else (forked from and joined back into the tick's stream, also under CUDA-graph capture)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // optional per-kernel timing (wc_mpc_enable_profiling): events recorded on the launch stream
    int profiling;
    std::vector<cudaEvent_t> ev; // quintuples (start, after K1, after setup, after phase 1, after K2)
    double prep_ms_acc, admm_ms_acc, setup_ms_acc, phase1_ms_acc, phase2_ms_acc;
    int64_t timed_ticks;
    double detail[3]; // setup / phase 1 / phase 2 of the last wc_mpc_kernel_times_ms window
};

namespace {

template <typename T>
cudaError_t dalloc(wc_mpc* h, T** p, size_t count)
{
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemset(q, 0, count * sizeof(T));
    if (e != cudaSuccess) return e;
    h->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return cudaSuccess;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int env_int(const char* name, int dflt)
{
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

} // namespace

extern "C" {

const char* wc_version(void) { return "wc_b200 0.1 sm_100a"; }

wc_status wc_mpc_create(const wc_mpc_config* cfg, int32_t B, int32_t device, wc_mpc** out)
{
    if (!cfg || !out || B < 1 || cfg->horizon < 1) return WC_ERR_INVALID_ARG;
    *out = nullptr;
    // diagonal weights (all shipped configurations) take the per-axis chain kernels; arbitrary 2x2 triplets the cooperative kernel's
    // block chain
    const bool coupled = cfg->Q[1] != 0.0 || cfg->Q[2] != 0.0 || cfg->R[1] != 0.0 || cfg->R[2] != 0.0;
    if (!(cfg->sampling_time > 0) || !(cfg->com_height > 0) || !(cfg->gravity > 0)) return WC_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return WC_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return WC_ERR_CUDA;

    wc_mpc* h = new (std::nothrow) wc_mpc();
    if (!h) return WC_ERR_ALLOC;
    h->cfg = *cfg;
    h->device = device;
    h->launches = 0;
    h->tick = 0;
    h->profiling = 0; h->prep_ms_acc = 0; h->admm_ms_acc = 0; h->timed_ticks = 0;
    h->setup_ms_acc = h->phase1_ms_acc = h->phase2_ms_acc = 0;
    MpcDev& d = h->d;
    std::memset(&d, 0, sizeof(d));
    d.B = B;
    d.N = cfg->horizon;
    const double omega = std::sqrt(cfg->gravity / cfg->com_height);
    d.a_dyn = std::exp(omega * cfg->sampling_time);
    d.b_dyn = 1.0 - std::exp(omega * cfg->sampling_time);
    const double x1 = cfg->foot_x[0], x2 = cfg->foot_x[1], y1 = cfg->foot_y[0], y2 = cfg->foot_y[1];
    d.foot_front = std::fabs(x1 > x2 ? x1 : x2);
    d.foot_back = std::fabs(x1 < x2 ? x1 : x2);
    d.foot_left = std::fabs(y1 > y2 ? y1 : y2);
    d.foot_right = std::fabs(y1 < y2 ? y1 : y2);
    d.init_zmp[0] = cfg->initial_zmp[0];
    d.init_zmp[1] = cfg->initial_zmp[1];
    d.hull_tol = cfg->hull_tolerance;
    d.st = cfg->osqp;
    d.rho_interval = cfg->osqp.adaptive_rho_interval;
    if (cfg->osqp.adaptive_rho && d.rho_interval == 0)
        d.rho_interval = cfg->osqp.check_termination ? 4 * cfg->osqp.check_termination : 100;
    d.ring_head = 0;
    d.coupled = coupled ? 1 : 0;
    d.Qoff[0] = cfg->Q[1]; d.Qoff[1] = cfg->Q[2]; d.Roff[0] = cfg->R[1]; d.Roff[1] = cfg->R[2];

    const size_t Bz = (size_t)B;
    const int nA = mpc_nA(d.N), nR = mpc_nR(d.N);
    cudaError_t e = cudaSuccess;
#define TRY(x) do { if (e == cudaSuccess) e = (x); } while (0)
    TRY(dalloc(h, &d.Qd, 2 * Bz)); TRY(dalloc(h, &d.Rd, 2 * Bz));
    TRY(dalloc(h, &d.x, 2 * nA * Bz)); TRY(dalloc(h, &d.yeq, 2 * nR * Bz));
    TRY(dalloc(h, &d.yh, 8 * Bz)); TRY(dalloc(h, &d.zh, 8 * Bz)); TRY(dalloc(h, &d.z0, 2 * Bz));
    TRY(dalloc(h, &d.D, 2 * nA * Bz)); TRY(dalloc(h, &d.Eeq, 2 * nR * Bz)); TRY(dalloc(h, &d.Eh, 8 * Bz));
    TRY(dalloc(h, &d.cscale, Bz)); TRY(dalloc(h, &d.rho, Bz));
    TRY(dalloc(h, &d.grad, 2 * nR * Bz));
    TRY(dalloc(h, &d.hullA, 16 * Bz)); TRY(dalloc(h, &d.hullb, 8 * Bz)); TRY(dalloc(h, &d.hullv, 16 * Bz));
    TRY(dalloc(h, &d.uprev, 2 * Bz));
    TRY(dalloc(h, &d.nhull, Bz)); TRY(dalloc(h, &d.feet, Bz)); TRY(dalloc(h, &d.have_solver, Bz));
    TRY(dalloc(h, &d.inited, Bz)); TRY(dalloc(h, &d.rho_updates, Bz)); TRY(dalloc(h, &d.tick_flags, Bz));
    if (e != cudaSuccess) { wc_mpc_destroy(h); return e == cudaErrorMemoryAllocation ? WC_ERR_ALLOC : WC_ERR_CUDA; }
    {
        std::vector<double> w(2 * Bz), r(2 * Bz), c1(Bz, 1.0), rho(Bz, cfg->osqp.rho), up(2 * Bz);
        for (size_t i = 0; i < Bz; i++) {
            w[i] = cfg->Q[0]; w[Bz + i] = cfg->Q[3];
            r[i] = cfg->R[0]; r[Bz + i] = cfg->R[3];
            up[i] = cfg->initial_zmp[0]; up[Bz + i] = cfg->initial_zmp[1];
        }
        TRY(cudaMemcpy(d.Qd, w.data(), 2 * Bz * sizeof(double), cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(d.Rd, r.data(), 2 * Bz * sizeof(double), cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(d.cscale, c1.data(), Bz * sizeof(double), cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(d.rho, rho.data(), Bz * sizeof(double), cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(d.uprev, up.data(), 2 * Bz * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (e != cudaSuccess) { wc_mpc_destroy(h); return WC_ERR_CUDA; }
#undef TRY
    // lanes per instance for the ADMM kernel: more lanes for longer horizons (more parallel
    // vector work per sequential sweep).  WC_MPC_LANES overrides (tuning / tests).
    int lanes = d.N <= 8 ? 8 : (d.N <= 128 ? 16 : 32);
    // the segment-parallel chain solve wants a whole warp per instance (16 segments per axis) as soon as its staging fits
    // (measured at B = 65536: N = 100 2.08e6 ticks/s with 32 lanes against 1.66e6 with 16; N = 128 1.0e6 with 16 lanes)
    if (!d.coupled && d.N >= kParMinHorizon && mpc_par_ok(d.N, 32)) lanes = 32;
    lanes = env_int("WC_MPC_LANES", lanes);
    if (!(lanes == 2 || lanes == 4 || lanes == 8 || lanes == 16 || lanes == 32)) lanes = 8;
    // the per-instance working set must fit in one CTA's shared memory
    int stride = mpc_smem_doubles(d.N, d.coupled);
    while ((stride % 16) != 8) stride++;
    while (lanes < 32 && (size_t)stride * 8 * (32 / lanes) > (size_t)mpc_admm_max_smem()) lanes *= 2;
    if ((size_t)stride * 8 * (32 / lanes) > (size_t)mpc_admm_max_smem()) { wc_mpc_destroy(h); return WC_ERR_UNSUPPORTED; }
    // long horizons: segment-parallel chain solve (every lane one segment of one axis chain) when its extra working set fits;
    // WC_MPC_PAR=0|1 overrides (tests / tuning)
    {
        int want = (d.N >= kParMinHorizon && !d.coupled) ? 1 : 0;
        want = env_int("WC_MPC_PAR", want);
        if (want && !d.coupled && mpc_par_ok(d.N, lanes)) {
            int sp = mpc_smem_doubles(d.N, 0, 1);
            while ((sp % 16) != 8) sp++;
            if ((size_t)sp * 8 * (32 / lanes) <= (size_t)mpc_admm_max_smem()) d.par = 1;
        }
    }
    h->lanes = lanes;
    // short horizons: one thread per axis chain (16 instances per warp) when the working set of a warp
    // fits in shared memory; WC_MPC_PATH=coop|chain overrides (tuning / tests)
    // and four lanes per instance with the iterate in registers (mpc_split.cu) at the compile-time horizons
    // (measured at B = 65536: chain 6.5e6 ticks/s at N = 50; from there the cooperative kernel's segment-parallel solve wins --
    // N = 60: 3.95e6 against 3.2e6, N = 80: 3.2e6 against 2.5e6, N = 100: 2.08e6 against 2.0e6)
    h->chain = d.coupled ? 0 : (mpc_split_supported(d.N) ? 2 : ((mpc_chain_preferred(d.N) && !(d.par && d.N >= 54)) ? 1 : 0));
    if (const char* pth = std::getenv("WC_MPC_PATH")) {
        if (!std::strcmp(pth, "coop") || d.coupled) h->chain = 0;
        else if (!std::strcmp(pth, "chain")) {
            if (!mpc_chain_supported(d.N)) { wc_mpc_destroy(h); return WC_ERR_UNSUPPORTED; }
            h->chain = 1;
        } else if (!std::strcmp(pth, "split")) {
            if (!mpc_split_supported(d.N)) { wc_mpc_destroy(h); return WC_ERR_UNSUPPORTED; }
            h->chain = 2;
        }
    }
    if (h->chain) {
        cudaError_t e2 = cudaSuccess;
        if (e2 == cudaSuccess) e2 = dalloc(h, &d.Dinv, 2 * nA * Bz);
        if (e2 == cudaSuccess) e2 = dalloc(h, &d.Einv, 2 * nR * Bz);
        if (e2 == cudaSuccess) e2 = dalloc(h, &d.Ehinv, 8 * Bz);
        if (h->chain == 1) { // the split-chain kernels keep neither the factors nor delta_y in HBM
            if (e2 == cudaSuccess) e2 = dalloc(h, &d.dy_scratch, (2 * nR + 8) * Bz);
            if (e2 == cudaSuccess) e2 = dalloc(h, &d.L1, 2 * nA * Bz);
            if (e2 == cudaSuccess) e2 = dalloc(h, &d.L2, 2 * nA * Bz);
            if (e2 == cudaSuccess) e2 = dalloc(h, &d.DI, 2 * nA * Bz);
            if (e2 == cudaSuccess) e2 = dalloc(h, &d.CF, 10 * Bz);
        }
        if (e2 == cudaSuccess) e2 = dalloc(h, &d.setup_list, Bz);
        if (e2 == cudaSuccess) e2 = dalloc(h, &d.setup_count, 1);
        if (e2 == cudaSuccess) e2 = dalloc(h, &d.cont_list, Bz);
        if (e2 == cudaSuccess) e2 = dalloc(h, &d.cont_count, 1);
        // measured and left OFF: next to phase 1 the side kernels take resident-CTA slots from it (shared memory is what
        // limits its occupancy), and the listed instances' phase 1 reads through L2 synchronously: 0.777 against 0.761 ms per MPC tick
        if (h->chain == 2 && env_int("WC_MPC_SETUP_OVERLAP", 0)) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            if (e2 == cudaSuccess) e2 = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi);
            if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
            if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
            d.setup_overlap = 1;
        }
        if (e2 != cudaSuccess) { wc_mpc_destroy(h); return e2 == cudaErrorMemoryAllocation ? WC_ERR_ALLOC : WC_ERR_CUDA; }
    }
    *out = h;
    return WC_OK;
}

void wc_mpc_destroy(wc_mpc* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (void* p : h->allocs) cudaFree(p);
    h->pipe.destroy();
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
}

int32_t wc_mpc_batch(const wc_mpc* h) { return h ? h->d.B : 0; }
int32_t wc_mpc_horizon(const wc_mpc* h) { return h ? h->d.N : 0; }
int64_t wc_mpc_launch_count(const wc_mpc* h) { return h ? h->launches : 0; }

wc_status wc_mpc_set_weights_host(wc_mpc* h, const double* q_diag, const double* r_diag)
{
    if (!h || !q_diag || !r_diag) return WC_ERR_INVALID_ARG;
    cudaSetDevice(h->device);
    if (h->pipe.ready && h->pipe.sync_all() != cudaSuccess) return WC_ERR_CUDA; // host ticks in flight read Qd / Rd
    const size_t bytes = 2 * (size_t)h->d.B * sizeof(double);
    if (cudaMemcpy(h->d.Qd, q_diag, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return WC_ERR_CUDA;
    if (cudaMemcpy(h->d.Rd, r_diag, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return WC_ERR_CUDA;
    return WC_OK;
}

wc_status wc_mpc_reset(wc_mpc* h, const uint8_t* mask_host)
{
    if (!h) return WC_ERR_INVALID_ARG;
    cudaSetDevice(h->device);
    if (h->pipe.ready && h->pipe.sync_all() != cudaSuccess) return WC_ERR_CUDA; // host ticks in flight read / write `feet`
    const size_t B = (size_t)h->d.B;
    if (!mask_host) {
        if (cudaMemset(h->d.feet, 0, B) != cudaSuccess) return WC_ERR_CUDA;
        return WC_OK;
    }
    std::vector<uint8_t> f(B);
    if (cudaMemcpy(f.data(), h->d.feet, B, cudaMemcpyDeviceToHost) != cudaSuccess) return WC_ERR_CUDA;
    for (size_t i = 0; i < B; i++) if (mask_host[i]) f[i] = 0;
    if (cudaMemcpy(h->d.feet, f.data(), B, cudaMemcpyHostToDevice) != cudaSuccess) return WC_ERR_CUDA;
    return WC_OK;
}

} // extern "C"

// one tick on stream s; ring_newest / ring_window != NULL: the reference window lives in the device-side ring
static wc_status mpc_tick_impl(wc_mpc* h, MpcTickIO io, const double* ring_newest, const double* ring_window, cudaStream_t s)
{
    // gradient ring: stage k of this tick lives in slot (head + k) mod (N+1); advancing the head by
    // one per tick is MPCSolver::setGradient's block shift (MPCSolver.cpp:216-221)
    h->d.ring_head = (int)(h->tick % (int64_t)(h->d.N + 1));
    if (ring_newest || ring_window) {
        if (!h->d.ref_ring) {
            if (!ring_window) return WC_ERR_INVALID_ARG; // the first ring tick needs the whole window
            if (dalloc(h, &h->d.ref_ring, (size_t)2 * (h->d.N + 1) * (size_t)h->d.B) != cudaSuccess) return WC_ERR_ALLOC;
        }
        launch_mpc_ref_ring_update(h->d, ring_newest, ring_window, s);
        h->launches += 1;
        io.ref = nullptr; io.ref_is_ring = 1;
    }
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (h->profiling) {
        for (int k = 0; k < 5; k++) { cudaEventCreate(&ev[k]); h->ev.push_back(ev[k]); }
        cudaEventRecord(ev[0], s);
    }
    launch_mpc_prepare(h->d, io, s);
    cudaError_t e = cudaGetLastError();
    if (h->profiling) cudaEventRecord(ev[1], s);
    if (e == cudaSuccess) {
        if (h->chain == 2) e = launch_mpc_split(h->d, io, s, h->profiling ? &ev[2] : nullptr, h->side, h->ev_fork, h->ev_join);
        else if (h->chain) e = launch_mpc_chain(h->d, io, s, h->profiling ? &ev[2] : nullptr);
        else {
            if (h->profiling) { cudaEventRecord(ev[2], s); }
            e = launch_mpc_admm(h->d, io, h->lanes, s);
            if (h->profiling) { cudaEventRecord(ev[3], s); }
        }
    }
    if (h->profiling) cudaEventRecord(ev[4], s);
    h->launches += h->chain ? (h->d.setup_overlap ? 5 : 4) : 2;
    h->tick += 1;
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[wc_mpc_tick_dev] CUDA error: %s\n", cudaGetErrorString(e));
        return WC_ERR_CUDA;
    }
    return WC_OK;
}

extern "C" {

wc_status wc_mpc_tick_dev(wc_mpc* h, const double* lf_H, const double* rf_H,
                          const uint8_t* l_contact, const uint8_t* r_contact,
                          const double* dcm_meas, const double* ref, const uint8_t* reset,
                          double* zmp, int32_t* osqp_status, int32_t* iters, int32_t* flags, void* stream)
{
    if (!h || !lf_H || !rf_H || !l_contact || !r_contact || !dcm_meas || !ref) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    MpcTickIO io{lf_H, rf_H, l_contact, r_contact, dcm_meas, ref, reset, zmp, osqp_status, iters, flags, 0};
    return mpc_tick_impl(h, io, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

wc_status wc_mpc_tick_ring_dev(wc_mpc* h, const double* lf_H, const double* rf_H,
                               const uint8_t* l_contact, const uint8_t* r_contact,
                               const double* dcm_meas, const double* ref_newest, const double* ref_window,
                               const uint8_t* reset, double* zmp, int32_t* osqp_status, int32_t* iters, int32_t* flags,
                               void* stream)
{
    if (!h || !lf_H || !rf_H || !l_contact || !r_contact || !dcm_meas || (!ref_newest && !ref_window)) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    MpcTickIO io{lf_H, rf_H, l_contact, r_contact, dcm_meas, nullptr, reset, zmp, osqp_status, iters, flags, 1};
    return mpc_tick_impl(h, io, ref_newest, ref_window, static_cast<cudaStream_t>(stream));
}

wc_status wc_mpc_tick_host_async(wc_mpc* h, const double* lf_H, const double* rf_H,
                           const uint8_t* l_contact, const uint8_t* r_contact,
                           const double* dcm_meas, const double* ref, const uint8_t* reset,
                           double* zmp, int32_t* osqp_status, int32_t* iters, int32_t* flags)
{
    if (!h || !lf_H || !rf_H || !l_contact || !r_contact || !dcm_meas || !ref) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t B = (size_t)h->d.B;
    const int N = h->d.N;
    // one packed input block and one packed output block -> one H2D and one D2H copy per tick
    const size_t o_lf = 0, o_rf = o_lf + 12 * B * 8, o_dcm = o_rf + 12 * B * 8, o_ref = o_dcm + 2 * B * 8;
    const size_t o_lc = o_ref + (size_t)2 * (N + 1) * B * 8, o_rc = o_lc + align_up(B, 8), o_rs = o_rc + align_up(B, 8);
    const size_t in_bytes = o_rs + align_up(B, 8);
    const size_t p_zmp = 0, p_st = p_zmp + 2 * B * 8, p_it = p_st + align_up(B * 4, 8), p_fl = p_it + align_up(B * 4, 8);
    const size_t out_bytes = p_fl + align_up(B * 4, 8);
    if (h->pipe.init(in_bytes, out_bytes) != cudaSuccess) return WC_ERR_ALLOC;
    // copies go straight from / to the caller's buffers: pinned buffers (cudaHostAlloc / torch
    // pin_memory) make them true async DMA transfers, pageable ones are staged by the driver
    cudaError_t ce = cudaSuccess;
    const int slot = h->pipe.begin(&ce);
    if (ce != cudaSuccess) return WC_ERR_CUDA;
    cudaStream_t s = h->pipe.up;
    char* dbase = static_cast<char*>(h->pipe.dev_in[slot]);
#define H2D(off, src, bytes) do { if (ce == cudaSuccess) ce = cudaMemcpyAsync(dbase + (off), (src), (bytes), cudaMemcpyHostToDevice, s); } while (0)
    H2D(o_lf, lf_H, 12 * B * 8); H2D(o_rf, rf_H, 12 * B * 8); H2D(o_dcm, dcm_meas, 2 * B * 8);
    H2D(o_ref, ref, (size_t)2 * (N + 1) * B * 8); H2D(o_lc, l_contact, B); H2D(o_rc, r_contact, B);
    if (reset) H2D(o_rs, reset, B);
    else if (ce == cudaSuccess) ce = cudaMemsetAsync(dbase + o_rs, 0, B, s);
#undef H2D
    if (ce == cudaSuccess) ce = h->pipe.before_run(slot);
    if (ce != cudaSuccess) return WC_ERR_CUDA;
    char* di = dbase;
    char* dout = static_cast<char*>(h->pipe.dev_out[slot]);
    wc_status rc = wc_mpc_tick_dev(h, (const double*)(di + o_lf), (const double*)(di + o_rf),
                                   (const uint8_t*)(di + o_lc), (const uint8_t*)(di + o_rc),
                                   (const double*)(di + o_dcm), (const double*)(di + o_ref),
                                   (const uint8_t*)(di + o_rs),
                                   (double*)(dout + p_zmp), (int32_t*)(dout + p_st), (int32_t*)(dout + p_it),
                                   (int32_t*)(dout + p_fl), h->pipe.run);
    if (rc != WC_OK) return rc;
    cudaError_t e = h->pipe.before_down(slot);
    s = h->pipe.down;
#define D2H(dst, off, bytes) do { if ((dst) && e == cudaSuccess) e = cudaMemcpyAsync((dst), dout + (off), (bytes), cudaMemcpyDeviceToHost, s); } while (0)
    D2H(zmp, p_zmp, 2 * B * 8); D2H(osqp_status, p_st, B * 4); D2H(iters, p_it, B * 4); D2H(flags, p_fl, B * 4);
#undef D2H
    if (e == cudaSuccess) e = h->pipe.end(slot);
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[wc_mpc_tick_host_async] CUDA error: %s\n", cudaGetErrorString(e));
        return WC_ERR_CUDA;
    }
    return WC_OK;
}

// Fused host tick of both blocks (wc_capi.h): one packed upload, the MPC kernels on the pipe's kernel stream and the
// FK + QP-IK kernels on its second one (the two blocks of a tick are independent given the host's inputs), one packed
// download.  The feet transforms are uploaded once for both blocks, the stance-foot transform is selected on the device
// when base_H is NULL, and the DCM reference travels as ONE new sample per tick (device-side ring).
wc_status wc_tick_host_async(wc_mpc* h, wc_ik* ik, const wc_tick_host_in* in, const wc_tick_host_out* out)
{
    if (!h || !ik || !in || !out) return WC_ERR_INVALID_ARG;
    if (!in->lf_H || !in->rf_H || !in->l_contact || !in->r_contact || !in->left_is_fixed || !in->dcm_meas ||
        (!in->ref_newest && !in->ref_window) || !in->q_meas || !in->q_lin || !in->lf_twist || !in->rf_twist ||
        !in->com_pos_des || !in->com_vel_des || !in->com_meas || !in->neck_R_des)
        return WC_ERR_INVALID_ARG;
    if (wc_ik_batch(ik) != h->d.B) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t B = (size_t)h->d.B, nJ = (size_t)wc_ik_dofs(ik);
    const int N = h->d.N;
    // packed input block: doubles first (rows per instance), then the byte flags
    enum { I_LF, I_RF, I_DCM, I_REFN, I_QM, I_QL, I_LT, I_RT, I_CP, I_CV, I_CM, I_NR, I_BASE, I_REFW, I_ND };
    const size_t rows[I_ND] = {12, 12, 2, 2, nJ, nJ, 6, 6, 3, 3, 3, 9, 12, (size_t)2 * (N + 1)};
    const double* src[I_ND] = {in->lf_H, in->rf_H, in->dcm_meas, in->ref_newest, in->q_meas, in->q_lin, in->lf_twist, in->rf_twist,
                               in->com_pos_des, in->com_vel_des, in->com_meas, in->neck_R_des, in->base_H, in->ref_window};
    size_t off[I_ND + 1]; off[0] = 0;
    for (int k = 0; k < I_ND; k++) off[k + 1] = off[k] + rows[k] * B * 8;
    const size_t o_lc = off[I_ND], o_rc = o_lc + align_up(B, 8), o_lif = o_rc + align_up(B, 8), o_rs = o_lif + align_up(B, 8);
    const size_t in_bytes = o_rs + align_up(B, 8);
    const size_t p_zmp = 0, p_dq = p_zmp + 2 * B * 8, p_le = p_dq + nJ * B * 8, p_re = p_le + 6 * B * 8;
    const size_t p_st = p_re + 6 * B * 8, p_it = p_st + align_up(B * 4, 8), p_fl = p_it + align_up(B * 4, 8);
    const size_t p_qs = p_fl + align_up(B * 4, 8), p_nw = p_qs + align_up(B * 4, 8), out_bytes = p_nw + align_up(B * 4, 8);
    if (h->pipe.init(in_bytes, out_bytes) != cudaSuccess) return WC_ERR_ALLOC;
    cudaError_t ce = cudaSuccess;
    const int slot = h->pipe.begin(&ce);
    if (ce != cudaSuccess) return WC_ERR_CUDA;
    cudaStream_t s = h->pipe.up;
    char* di = static_cast<char*>(h->pipe.dev_in[slot]);
    char* dout = static_cast<char*>(h->pipe.dev_out[slot]);
    // One copy per RUN of arrays that are contiguous on both sides: a caller that keeps its inputs in one arena in the packed
    // order of this block (wc_capi.h) gets two copies per tick instead of sixteen -- a 1-12 MB copy reaches half the PCIe rate
    // of a 55 MB one on this box (tools/pcie_probe.py: 24 against 51 GB/s).
    {
        struct Seg { size_t dst; const char* src; size_t bytes; };
        Seg segs[I_ND + 4]; int ns = 0;
        for (int k = 0; k < I_ND; k++) if (src[k]) segs[ns++] = Seg{off[k], (const char*)src[k], rows[k] * B * 8};
        segs[ns++] = Seg{o_lc, (const char*)in->l_contact, B};
        segs[ns++] = Seg{o_rc, (const char*)in->r_contact, B};
        segs[ns++] = Seg{o_lif, (const char*)in->left_is_fixed, B};
        if (in->reset) segs[ns++] = Seg{o_rs, (const char*)in->reset, B};
        for (int k = 0; k < ns && ce == cudaSuccess;) {
            size_t bytes = segs[k].bytes; int j = k + 1;
            // (the byte flags sit align_up(B, 8) apart on the device: mergeable when the host keeps the same spacing)
            while (j < ns && segs[j].src - segs[k].src == (ptrdiff_t)(segs[j].dst - segs[k].dst) && segs[j].dst >= segs[k].dst + bytes &&
                   segs[j].dst - (segs[k].dst + bytes) < 8) { bytes = segs[j].dst - segs[k].dst + segs[j].bytes; j++; }
            ce = cudaMemcpyAsync(di + segs[k].dst, segs[k].src, bytes, cudaMemcpyHostToDevice, s);
            k = j;
        }
        if (!in->reset && ce == cudaSuccess) ce = cudaMemsetAsync(di + o_rs, 0, B, s);
    }
    if (ce == cudaSuccess) ce = h->pipe.before_run(slot);
    if (ce != cudaSuccess) return WC_ERR_CUDA;
    auto dp = [&](int k) { return (const double*)(di + off[k]); };
    // fork: the IK kernels on the second kernel stream, beside the MPC kernels
    if (cudaEventRecord(h->pipe.fork, h->pipe.run) != cudaSuccess || cudaStreamWaitEvent(h->pipe.run2, h->pipe.fork, 0) != cudaSuccess)
        return WC_ERR_CUDA;
    MpcTickIO io{dp(I_LF), dp(I_RF), (const uint8_t*)(di + o_lc), (const uint8_t*)(di + o_rc), dp(I_DCM), nullptr,
                 (const uint8_t*)(di + o_rs), (double*)(dout + p_zmp), (int32_t*)(dout + p_st), (int32_t*)(dout + p_it),
                 (int32_t*)(dout + p_fl), 1};
    wc_status rc = mpc_tick_impl(h, io, in->ref_window ? nullptr : dp(I_REFN), in->ref_window ? dp(I_REFW) : nullptr, h->pipe.run);
    if (rc != WC_OK) return rc;
    const double* base = in->base_H ? dp(I_BASE) : nullptr; // NULL: the stance foot's transform, selected in the kernel
    rc = wc_ik_tick_dev(ik, dp(I_QM), dp(I_QL), base, (const uint8_t*)(di + o_lif), dp(I_LF), dp(I_RF), dp(I_LT), dp(I_RT),
                        dp(I_CP), dp(I_CV), dp(I_CM), dp(I_NR), (double*)(dout + p_dq), (double*)(dout + p_le),
                        (double*)(dout + p_re), (int32_t*)(dout + p_qs), (int32_t*)(dout + p_nw), h->pipe.run2);
    if (rc != WC_OK) return rc;
    // join
    if (cudaEventRecord(h->pipe.join, h->pipe.run2) != cudaSuccess || cudaStreamWaitEvent(h->pipe.run, h->pipe.join, 0) != cudaSuccess)
        return WC_ERR_CUDA;
    cudaError_t e = h->pipe.before_down(slot);
    s = h->pipe.down;
    { // results: likewise one copy per run that is contiguous on both sides (packed order: zmp, dq, lf_err, rf_err, osqp_status, iters,
      // flags, qp_status, nwsr)
        struct Seg { size_t src; char* dst; size_t bytes; };
        Seg segs[9]; int ns = 0;
        auto add = [&](void* dst, size_t o, size_t bytes) { if (dst) segs[ns++] = Seg{o, (char*)dst, bytes}; };
        add(out->zmp, p_zmp, 2 * B * 8); add(out->dq, p_dq, nJ * B * 8); add(out->lf_err, p_le, 6 * B * 8); add(out->rf_err, p_re, 6 * B * 8);
        add(out->osqp_status, p_st, B * 4); add(out->iters, p_it, B * 4); add(out->flags, p_fl, B * 4);
        add(out->qp_status, p_qs, B * 4); add(out->nwsr, p_nw, B * 4);
        for (int k = 0; k < ns && e == cudaSuccess;) {
            size_t bytes = segs[k].bytes; int j = k + 1;
            while (j < ns && segs[j].dst - segs[k].dst == (ptrdiff_t)(segs[j].src - segs[k].src) && segs[j].src >= segs[k].src + bytes &&
                   segs[j].src - (segs[k].src + bytes) < 8) { bytes = segs[j].src - segs[k].src + segs[j].bytes; j++; }
            e = cudaMemcpyAsync(segs[k].dst, dout + segs[k].src, bytes, cudaMemcpyDeviceToHost, s);
            k = j;
        }
    }
    if (e == cudaSuccess) e = h->pipe.end(slot);
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[wc_tick_host_async] CUDA error: %s\n", cudaGetErrorString(e));
        return WC_ERR_CUDA;
    }
    return WC_OK;
}

wc_status wc_mpc_sync(wc_mpc* h)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (!h->pipe.ready) return WC_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const cudaError_t e = h->pipe.sync_all();
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[wc_mpc_sync] CUDA error: %s\n", cudaGetErrorString(e));
        return WC_ERR_CUDA;
    }
    return WC_OK;
}

wc_status wc_mpc_sync_one(wc_mpc* h)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (!h->pipe.ready) return WC_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    return h->pipe.sync_one() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_mpc_tick_host(wc_mpc* h, const double* lf_H, const double* rf_H,
                           const uint8_t* l_contact, const uint8_t* r_contact,
                           const double* dcm_meas, const double* ref, const uint8_t* reset,
                           double* zmp, int32_t* osqp_status, int32_t* iters, int32_t* flags)
{
    const wc_status rc = wc_mpc_tick_host_async(h, lf_H, rf_H, l_contact, r_contact, dcm_meas, ref, reset, zmp, osqp_status, iters, flags);
    return rc != WC_OK ? rc : wc_mpc_sync(h);
}

wc_status wc_mpc_debug_get_host(wc_mpc* h, double* primal, double* rho, int32_t* rho_updates)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t B = (size_t)h->d.B;
    const size_t n = (size_t)(4 * h->d.N + 2);
    if (cudaDeviceSynchronize() != cudaSuccess) return WC_ERR_CUDA;
    if (primal) {
        double* tmp = nullptr;
        if (cudaMalloc(&tmp, n * B * sizeof(double)) != cudaSuccess) return WC_ERR_ALLOC;
        launch_mpc_export_primal(h->d, tmp, nullptr);
        h->launches += 1;
        cudaError_t e = cudaMemcpy(primal, tmp, n * B * sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(tmp);
        if (e != cudaSuccess) return WC_ERR_CUDA;
    }
    if (rho && cudaMemcpy(rho, h->d.rho, B * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return WC_ERR_CUDA;
    if (rho_updates && cudaMemcpy(rho_updates, h->d.rho_updates, B * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
        return WC_ERR_CUDA;
    return WC_OK;
}

wc_status wc_mpc_enable_profiling(wc_mpc* h, int32_t on)
{
    if (!h) return WC_ERR_INVALID_ARG;
    h->profiling = on ? 1 : 0;
    return WC_OK;
}

wc_status wc_mpc_kernel_times_ms(wc_mpc* h, double* prepare_ms, double* admm_ms, int64_t* ticks)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return WC_ERR_CUDA;
    for (size_t k = 0; k + 4 < h->ev.size(); k += 5) {
        float a = 0, b = 0, c1 = 0, c2 = 0, c3 = 0;
        cudaEventElapsedTime(&a, h->ev[k], h->ev[k + 1]);
        cudaEventElapsedTime(&b, h->ev[k + 1], h->ev[k + 4]);
        cudaEventElapsedTime(&c1, h->ev[k + 1], h->ev[k + 2]);
        cudaEventElapsedTime(&c2, h->ev[k + 2], h->ev[k + 3]);
        cudaEventElapsedTime(&c3, h->ev[k + 3], h->ev[k + 4]);
        h->prep_ms_acc += a; h->admm_ms_acc += b; h->timed_ticks += 1;
        h->setup_ms_acc += c1; h->phase1_ms_acc += c2; h->phase2_ms_acc += c3;
    }
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    h->ev.clear();
    if (prepare_ms) *prepare_ms = h->prep_ms_acc;
    if (admm_ms) *admm_ms = h->admm_ms_acc;
    if (ticks) *ticks = h->timed_ticks;
    h->detail[0] = h->setup_ms_acc; h->detail[1] = h->phase1_ms_acc; h->detail[2] = h->phase2_ms_acc;
    h->prep_ms_acc = 0; h->admm_ms_acc = 0; h->timed_ticks = 0;
    h->setup_ms_acc = h->phase1_ms_acc = h->phase2_ms_acc = 0;
    return WC_OK;
}

wc_status wc_mpc_kernel_times_detail_ms(wc_mpc* h, double* setup_ms, double* phase1_ms, double* phase2_ms)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (setup_ms) *setup_ms = h->detail[0];
    if (phase1_ms) *phase1_ms = h->detail[1];
    if (phase2_ms) *phase2_ms = h->detail[2];
    return WC_OK;
}

int32_t wc_mpc_path(const wc_mpc* h) { return h ? h->chain : -1; }

wc_status wc_mpc_set_tick_source_dev(wc_mpc* h, const int32_t* tick_dev)
{
    if (!h) return WC_ERR_INVALID_ARG;
    h->d.tick_dev = tick_dev;
    return WC_OK;
}

} // extern "C"
