// wc_ik_api.cu — C ABI of the batched FK + QP-IK (include/wc_capi.h): model pre-processing,
// handle management, kernel launches, pinned host staging.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/wc_capi.h"
#include "host_pipe.cuh"
#include "ik_kernels.cuh"

using namespace wcb;

struct wc_ik {
    IkDev d;
    int device;
    int64_t launches;
    IkModelDev* model_dev;
    double* kpos; double* katt;
    double* osqp_state; // one allocation behind d.ox .. d.osetups
    int32_t* user_setups; int32_t* user_flags;
    HostPipe pipe; // double-buffered staging of the host-pointer tick
    int profiling;
    std::vector<cudaEvent_t> ev; // triples (start, after the fast kernel, end)
    double fast_ms_acc, detail[2];
    double ms_acc; int64_t timed;
};

namespace {

bool anc_or_self(const wc_model* m, int anc, int link)
{
    while (link >= 0) { if (link == anc) return true; link = m->parent[link]; }
    return false;
}

// host-side pre-processing of the kinematic tree (what iDynTree::Traversal caches for a base link)
bool build_model(const wc_model* m, IkModelDev* o)
{
    std::memset(o, 0, sizeof(*o));
    const int L = m->n_links;
    if (L < 1 || L > WC_MAX_LINKS || m->n_dofs < 1 || m->n_dofs > WC_MAX_DOFS) return false;
    o->n_links = L; o->n_dofs = m->n_dofs;
    for (int d = 0; d < WC_MAX_DOFS; d++) o->link_of_dof[d] = -1;
    double total = 0.0;
    for (int i = 0; i < L; i++) {
        if (i == 0 ? m->parent[i] != -1 : (m->parent[i] < 0 || m->parent[i] >= i)) return false;
        o->parent[i] = m->parent[i];
        o->dof[i] = m->dof[i];
        if (m->dof[i] >= m->n_dofs) return false;
        if (m->dof[i] >= 0) { if (o->link_of_dof[m->dof[i]] != -1) return false; o->link_of_dof[m->dof[i]] = i; }
        o->depth[i] = i == 0 ? 0 : o->depth[m->parent[i]] + 1;
        if (o->depth[i] > o->max_depth) o->max_depth = o->depth[i];
        std::memcpy(o->joint_R[i], m->joint_R[i], sizeof(double) * 9);
        std::memcpy(o->joint_p[i], m->joint_p[i], sizeof(double) * 3);
        std::memcpy(o->axis[i], m->axis[i], sizeof(double) * 3);
        std::memcpy(o->com[i], m->com[i], sizeof(double) * 3);
        o->mass[i] = m->mass[i];
        total += m->mass[i];
    }
    if (!(total > 0.0)) return false;
    for (int d = 0; d < m->n_dofs; d++) if (o->link_of_dof[d] < 0) return false;
    o->total_mass = total;
    for (int c = 0; c < L; c++) {
        unsigned long long mask = 0; double ms = 0.0;
        for (int k = 0; k < L; k++) if (anc_or_self(m, c, k)) { mask |= 1ull << k; ms += m->mass[k]; }
        o->sub_mask[c] = mask; o->sub_mass[c] = ms;
    }
    { // pointer-jumping schedule (fk_stage<., true>)
        int anc[WC_MAX_LINKS];
        for (int i = 0; i < L; i++) anc[i] = m->parent[i];
        int r = 0;
        for (; r < WC_MAX_JUMP; r++) {
            bool any = false;
            for (int i = 0; i < L; i++) { o->jump[r][i] = (signed char)anc[i]; any = any || anc[i] >= 0; }
            if (!any) break;
            int nxt[WC_MAX_LINKS];
            for (int i = 0; i < L; i++) nxt[i] = anc[i] >= 0 ? anc[anc[i]] : -1;
            for (int i = 0; i < L; i++) anc[i] = nxt[i];
        }
        o->n_jump = r;
        bool left_over = false;
        for (int i = 0; i < L; i++) left_over = left_over || anc[i] >= 0;
        if (left_over) o->n_jump = 0; // deeper than 2^WC_MAX_JUMP: level-by-level composition
    }
    const int fl[3] = {m->l_sole_link, m->r_sole_link, m->neck_link};
    for (int f = 0; f < 3; f++) if (fl[f] < 0 || fl[f] >= L) return false;
    { // thread-per-instance schedule (ik_tpi.cu): transform slots by linear scan, contiguous sub-trees
        int last_use[WC_MAX_LINKS]; bool needs[WC_MAX_LINKS];
        for (int i = 0; i < L; i++) { last_use[i] = i; needs[i] = false; }
        for (int i = 1; i < L; i++) {
            const int p = m->parent[i];
            if (p != i - 1) { needs[p] = true; if (i > last_use[p]) last_use[p] = i; }
        }
        for (int f = 0; f < 3; f++) { needs[fl[f]] = true; last_use[fl[f]] = L; } // read after the loop
        int owner[kTpiMaxSlots]; for (int k = 0; k < kTpiMaxSlots; k++) owner[k] = -1;
        bool ok = true; int used = 0;
        for (int i = 0; i < L && ok; i++) {
            for (int k = 0; k < kTpiMaxSlots; k++) if (owner[k] >= 0 && last_use[owner[k]] < i) owner[k] = -1;
            o->tpi_pslot[i] = (signed char)(i == 0 ? -1 : (m->parent[i] == i - 1 ? kTpiPrev : o->tpi_slot[m->parent[i]]));
            o->tpi_slot[i] = -1;
            if (needs[i]) {
                int k = 0;
                while (k < kTpiMaxSlots && owner[k] >= 0) k++;
                if (k == kTpiMaxSlots) { ok = false; break; }
                owner[k] = i; o->tpi_slot[i] = (signed char)k; if (k + 1 > used) used = k + 1;
            }
        }
        for (int d = 0; d < m->n_dofs && ok; d++) {
            const int c = o->link_of_dof[d];
            int last = c;
            while (last + 1 < L && ((o->sub_mask[c] >> (last + 1)) & 1ull)) last++;
            unsigned long long run = 0;
            for (int k = c; k <= last; k++) run |= 1ull << k;
            if (run != o->sub_mask[c]) ok = false;
            o->tpi_sub_last[d] = (signed char)last;
        }
        o->tpi_ok = ok ? 1 : 0; o->tpi_slots = used;
    }
    o->base_link[1] = m->l_sole_link; o->base_link[0] = m->r_sole_link;
    for (int f = 0; f < 3; f++) o->frame_link[f] = fl[f];
    std::memcpy(o->frame_R[0], m->l_sole_R, sizeof(double) * 9); std::memcpy(o->frame_p[0], m->l_sole_p, sizeof(double) * 3);
    std::memcpy(o->frame_R[1], m->r_sole_R, sizeof(double) * 9); std::memcpy(o->frame_p[1], m->r_sole_p, sizeof(double) * 3);
    std::memcpy(o->frame_R[2], m->neck_R, sizeof(double) * 9); std::memcpy(o->frame_p[2], m->neck_p, sizeof(double) * 3);
    for (int left = 0; left < 2; left++) {
        const int base = o->base_link[left];
        for (int d = 0; d < m->n_dofs; d++) {
            const int c = o->link_of_dof[d];
            const bool on_b = anc_or_self(m, c, base);
            o->base_in_sub[left][d] = on_b ? 1 : 0;
            for (int f = 0; f < 3; f++) {
                const bool on_f = anc_or_self(m, c, fl[f]);
                o->s_frame[left][f][d] = (on_f == on_b) ? 0 : (on_f ? 1 : -1);
            }
        }
    }
    return true;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

} // namespace

// ---- run-time number of actuated joints (WalkingQPIK_qpOASES::initialize reads actuatedDOFs from the configuration,
// WM/src/WalkingQPInverseKinematics_qpOASES.cpp:53-80): one instantiation of the warp-per-instance kernels per joint count,
// 12 (two legs) .. 26 (6 + 26 = 32 QP variables = one warp)
namespace wcb {
#define WC_IK_NJ_LIST(X) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26)
#define WC_IK_DECL(n) cudaError_t launch_ik_nj##n(const IkDev&, const IkTickIO&, int, cudaStream_t, cudaEvent_t);
WC_IK_NJ_LIST(WC_IK_DECL)
#undef WC_IK_DECL
bool ik_supported(int n_dofs) { return n_dofs >= 12 && n_dofs <= 26; }
cudaError_t launch_ik(const IkDev& d, const IkTickIO& io, int mode, cudaStream_t s, cudaEvent_t mid)
{
    switch (d.nJ) {
#define WC_IK_CASE(n) case n: return launch_ik_nj##n(d, io, mode, s, mid);
        WC_IK_NJ_LIST(WC_IK_CASE)
#undef WC_IK_CASE
    default: return cudaErrorInvalidValue;
    }
}
} // namespace wcb

extern "C" {

wc_status wc_ik_create(const wc_ik_config* cfg, const wc_model* model, int32_t B, int32_t device, wc_ik** out)
{
    if (!cfg || !model || !out || B < 1) return WC_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->n_dofs != model->n_dofs) return WC_ERR_INVALID_ARG;
    if (!ik_supported(cfg->n_dofs)) return WC_ERR_UNSUPPORTED; // 12 .. 26 actuated joints: 6 + n QP variables fit one warp
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return WC_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return WC_ERR_CUDA;
    IkModelDev hm;
    if (!build_model(model, &hm)) return WC_ERR_INVALID_ARG;
    wc_ik* h = new (std::nothrow) wc_ik();
    if (!h) return WC_ERR_ALLOC;
    h->launches = 0; h->model_dev = nullptr; h->kpos = h->katt = nullptr; h->osqp_state = nullptr; h->user_setups = nullptr; h->user_flags = nullptr;
    h->profiling = 0; h->ms_acc = 0; h->timed = 0; h->fast_ms_acc = 0; h->detail[0] = h->detail[1] = 0;
    std::memset(&h->d, 0, sizeof(h->d));
    h->device = device;
    h->d.B = B; h->d.nJ = cfg->n_dofs; h->d.cfg = *cfg;
    if (cudaMalloc(&h->model_dev, (sizeof(IkModelDev) + 15) / 16 * 16) != cudaSuccess) // (whole 16-byte units: the fast kernel fetches it with one bulk copy)
        { delete h; return WC_ERR_ALLOC; }
    if (cudaMemcpy(h->model_dev, &hm, sizeof(IkModelDev), cudaMemcpyHostToDevice) != cudaSuccess) { wc_ik_destroy(h); return WC_ERR_CUDA; }
    h->d.model = h->model_dev;
    // fused-tick fast path (ik_fast_kernel) + compacted general solve; WC_IK_PATH=general disables it
    {
        const char* pth = std::getenv("WC_IK_PATH");
        const bool fast = cfg->use_com_as_constraint && !(pth && !std::strcmp(pth, "general"));
        if (fast) {
            if (cudaMalloc(&h->d.slow_list, (size_t)B * sizeof(int32_t)) != cudaSuccess ||
                cudaMalloc(&h->d.second_list, (size_t)2 * B * sizeof(int32_t)) != cudaSuccess ||
                cudaMalloc(&h->d.slow_count, 2 * sizeof(int32_t)) != cudaSuccess) { wc_ik_destroy(h); return WC_ERR_ALLOC; }
            cudaMemset(h->d.slow_count, 0, 2 * sizeof(int32_t));
            // thread-per-instance fast kernels (ik_tpi.cu) when the model's link numbering allows their schedule;
            // WC_IK_PATH=warp keeps the warp-per-instance fast kernel (ik_fast_kernel)
            if (hm.tpi_ok && !(pth && !std::strcmp(pth, "warp"))) {
                const size_t rows = (size_t)ik_tpi_rows(hm.n_links, cfg->n_dofs);
                if (cudaMalloc(&h->d.tpi_scratch, (rows * (size_t)B + 32) * sizeof(double)) != cudaSuccess) { wc_ik_destroy(h); return WC_ERR_ALLOC; }
                h->d.tpi_slots = hm.tpi_slots;
            }
        }
    }
    *out = h;
    return WC_OK;
}

void wc_ik_destroy(wc_ik* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->model_dev) cudaFree(h->model_dev);
    if (h->d.slow_list) cudaFree(h->d.slow_list);
    if (h->d.second_list) cudaFree(h->d.second_list);
    if (h->d.tpi_scratch) cudaFree(h->d.tpi_scratch);
    if (h->d.slow_count) cudaFree(h->d.slow_count);
    if (h->kpos) cudaFree(h->kpos);
    if (h->katt) cudaFree(h->katt);
    if (h->osqp_state) cudaFree(h->osqp_state);
    h->pipe.destroy();
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    delete h;
}

int32_t wc_ik_batch(const wc_ik* h) { return h ? h->d.B : 0; }
int32_t wc_ik_dofs(const wc_ik* h) { return h ? h->d.nJ : 0; }
int64_t wc_ik_launch_count(const wc_ik* h) { return h ? h->launches : 0; }

wc_status wc_ik_set_gains_host(wc_ik* h, const double* k_pos_foot, const double* k_att_foot)
{
    if (!h || !k_pos_foot || !k_att_foot) return WC_ERR_INVALID_ARG;
    cudaSetDevice(h->device);
    if (h->pipe.ready && h->pipe.sync_all() != cudaSuccess) return WC_ERR_CUDA; // host ticks in flight read the gains
    const size_t bytes = (size_t)h->d.B * sizeof(double);
    if (!h->kpos && cudaMalloc(&h->kpos, bytes) != cudaSuccess) return WC_ERR_ALLOC;
    if (!h->katt && cudaMalloc(&h->katt, bytes) != cudaSuccess) return WC_ERR_ALLOC;
    if (cudaMemcpy(h->kpos, k_pos_foot, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return WC_ERR_CUDA;
    if (cudaMemcpy(h->katt, k_att_foot, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return WC_ERR_CUDA;
    h->d.kpos = h->kpos; h->d.katt = h->katt;
    return WC_OK;
}

wc_status wc_ik_reset_solver(wc_ik* h)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (!h->osqp_state) return WC_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    if (h->pipe.ready && h->pipe.sync_all() != cudaSuccess) return WC_ERR_CUDA; // host ticks in flight use the workspaces
    const size_t B = (size_t)h->d.B;
    // oinit and osetups: no solver yet
    if (cudaMemset(h->d.oinit, 0, B) != cudaSuccess || cudaMemset(h->d.osetups, 0, B * sizeof(int32_t)) != cudaSuccess) return WC_ERR_CUDA;
    return WC_OK;
}

wc_status wc_ik_set_solver(wc_ik* h, int32_t solver, const wc_osqp_settings* settings)
{
    if (!h || (solver != WC_IK_SOLVER_QPOASES && solver != WC_IK_SOLVER_OSQP)) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    if (solver == WC_IK_SOLVER_OSQP) {
        if (settings) h->d.osqp = *settings; else wc_osqp_default_settings(&h->d.osqp);
        if (h->d.osqp.max_iter < 1 || h->d.osqp.scaling < 0 || h->d.osqp.check_termination < 0) return WC_ERR_INVALID_ARG;
        if (!h->osqp_state) {
            const size_t B = (size_t)h->d.B, nV = (size_t)h->d.nJ + 6, M = 15 + (size_t)h->d.nJ;
            const size_t doubles = (nV + 2 * M + 1) * B;
            const size_t bytes = doubles * sizeof(double) + nV * B * sizeof(uint32_t) + 2 * B * sizeof(int32_t) + align_up(B, 16);
            if (cudaMalloc(&h->osqp_state, bytes) != cudaSuccess) return WC_ERR_ALLOC;
            cudaMemset(h->osqp_state, 0, bytes);
            h->d.ox = h->osqp_state; h->d.oy = h->d.ox + nV * B; h->d.oz = h->d.oy + M * B; h->d.orho = h->d.oz + M * B;
            h->d.omask = reinterpret_cast<uint32_t*>(h->d.orho + B);
            h->d.osetups = reinterpret_cast<int32_t*>(h->d.omask + nV * B);
            h->d.oflags = h->d.osetups + B;
            h->d.oinit = reinterpret_cast<uint8_t*>(h->d.oflags + B);
        }
    }
    h->d.solver = solver;
    return wc_ik_reset_solver(h);
}

wc_status wc_ik_osqp_outputs_dev(wc_ik* h, int32_t* flags, int32_t* setups)
{
    if (!h) return WC_ERR_INVALID_ARG;
    h->user_flags = flags;
    h->user_setups = setups;
    return WC_OK;
}

wc_status wc_ik_osqp_info_host(wc_ik* h, int32_t* flags, int32_t* setups)
{
    if (!h || !h->osqp_state) return WC_ERR_STATE;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return WC_ERR_CUDA;
    const size_t bytes = (size_t)h->d.B * sizeof(int32_t);
    if (flags && cudaMemcpy(flags, h->d.oflags, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return WC_ERR_CUDA;
    if (setups && cudaMemcpy(setups, h->d.osetups, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return WC_ERR_CUDA;
    return WC_OK;
}

static wc_status launch_checked(wc_ik* h, const IkTickIO& io, int mode, cudaStream_t s, const char* who)
{
    cudaEvent_t e0 = nullptr, em = nullptr, e1 = nullptr;
    if (h->profiling) {
        cudaEventCreate(&e0); cudaEventCreate(&em); cudaEventCreate(&e1);
        h->ev.push_back(e0); h->ev.push_back(em); h->ev.push_back(e1);
        cudaEventRecord(e0, s);
    }
    const bool osqp = h->d.solver == WC_IK_SOLVER_OSQP && mode != IK_MODE_FK_ONLY;
    const bool kin_part = mode == IK_MODE_TICK_KIN; // a launch of its own only with the thread-per-instance fast path
    const bool two = (mode == IK_MODE_TICK || mode == IK_MODE_TICK_SOLVE) && h->d.slow_list && !osqp;
    if (h->profiling && !two) cudaEventRecord(em, s);
    cudaError_t e = launch_ik(h->d, io, mode, s, (h->profiling && two) ? em : nullptr);
    if (h->profiling) cudaEventRecord(e1, s);
    if (osqp && h->user_setups && e == cudaSuccess)
        e = cudaMemcpyAsync(h->user_setups, h->d.osetups, (size_t)h->d.B * sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    if (osqp && h->user_flags && e == cudaSuccess)
        e = cudaMemcpyAsync(h->user_flags, h->d.oflags, (size_t)h->d.B * sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    if (kin_part) h->launches += (h->d.tpi_scratch && h->d.slow_list && !osqp && h->d.cfg.use_com_as_constraint) ? 1 : 0;
    else h->launches += two ? (h->d.tpi_scratch ? 2 : 3) : 1; // fast kernel (thread per instance: one launch; warp per instance: two), general kernel
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[%s] CUDA error: %s\n", who, cudaGetErrorString(e));
        return WC_ERR_CUDA;
    }
    return WC_OK;
}

wc_status wc_ik_tick_dev(wc_ik* h, const double* q_meas, const double* q_lin, const double* base_H,
                         const uint8_t* left_is_fixed, const double* lf_H_des, const double* rf_H_des,
                         const double* lf_twist, const double* rf_twist, const double* com_pos_des,
                         const double* com_vel_des, const double* com_meas, const double* neck_R_des,
                         double* dq, double* lf_err, double* rf_err, int32_t* qp_status, int32_t* nwsr, void* stream)
{
    // base_H == NULL: the stance foot's desired transform (lf_H_des / rf_H_des by left_is_fixed)
    if (!h || !q_meas || !q_lin || !left_is_fixed || !lf_H_des || !rf_H_des || !lf_twist || !rf_twist ||
        !com_pos_des || !com_vel_des || !com_meas || !neck_R_des)
        return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    IkTickIO io;
    std::memset(&io, 0, sizeof(io));
    io.q_meas = q_meas; io.q_lin = q_lin; io.base_H = base_H; io.left_is_fixed = left_is_fixed;
    io.lf_H_des = lf_H_des; io.rf_H_des = rf_H_des; io.lf_twist = lf_twist; io.rf_twist = rf_twist;
    io.com_pos_des = com_pos_des; io.com_vel_des = com_vel_des; io.com_meas = com_meas; io.neck_R_des = neck_R_des;
    io.dq = dq; io.lf_err = lf_err; io.rf_err = rf_err; io.status = qp_status; io.nwsr = nwsr;
    return launch_checked(h, io, IK_MODE_TICK, static_cast<cudaStream_t>(stream), "wc_ik_tick_dev");
}

wc_status wc_ik_tick_part_dev(wc_ik* h, int32_t part, const double* q_meas, const double* q_lin, const double* base_H,
                              const uint8_t* left_is_fixed, const double* lf_H_des, const double* rf_H_des,
                              const double* lf_twist, const double* rf_twist, const double* com_pos_des,
                              const double* com_vel_des, const double* com_meas, const double* neck_R_des, double* com_out,
                              double* dq, double* lf_err, double* rf_err, int32_t* qp_status, int32_t* nwsr, void* stream)
{
    if (part != WC_IK_PART_KINEMATICS && part != WC_IK_PART_SOLVE) return WC_ERR_INVALID_ARG;
    if (!h || !q_meas || !q_lin || !left_is_fixed || !lf_H_des || !rf_H_des || !lf_twist || !rf_twist || !neck_R_des)
        return WC_ERR_INVALID_ARG;
    if (part == WC_IK_PART_SOLVE && (!com_pos_des || !com_vel_des || !com_meas)) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    IkTickIO io;
    std::memset(&io, 0, sizeof(io));
    io.q_meas = q_meas; io.q_lin = q_lin; io.base_H = base_H; io.left_is_fixed = left_is_fixed;
    io.lf_H_des = lf_H_des; io.rf_H_des = rf_H_des; io.lf_twist = lf_twist; io.rf_twist = rf_twist;
    io.com_pos_des = com_pos_des; io.com_vel_des = com_vel_des; io.com_meas = com_meas; io.neck_R_des = neck_R_des;
    if (part == WC_IK_PART_KINEMATICS) {
        io.o_com = com_out;
        return launch_checked(h, io, IK_MODE_TICK_KIN, static_cast<cudaStream_t>(stream), "wc_ik_tick_part_dev (kinematics)");
    }
    io.dq = dq; io.lf_err = lf_err; io.rf_err = rf_err; io.status = qp_status; io.nwsr = nwsr;
    return launch_checked(h, io, IK_MODE_TICK_SOLVE, static_cast<cudaStream_t>(stream), "wc_ik_tick_part_dev (solve)");
}

int32_t wc_ik_tick_is_split(const wc_ik* h)
{
    return (h && h->d.tpi_scratch && h->d.slow_list && h->d.solver != WC_IK_SOLVER_OSQP && h->d.cfg.use_com_as_constraint) ? 1 : 0;
}

wc_status wc_fk_dev(wc_ik* h, const double* q, const double* base_H, const uint8_t* left_is_fixed,
                    double* lf_H, double* rf_H, double* neck_R, double* com,
                    double* J_lf, double* J_rf, double* J_neck, double* J_com, void* stream)
{
    if (!h || !q || !base_H || !left_is_fixed) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    IkTickIO io;
    std::memset(&io, 0, sizeof(io));
    io.q_lin = q; io.base_H = base_H; io.left_is_fixed = left_is_fixed;
    io.o_lf_H = lf_H; io.o_rf_H = rf_H; io.o_neck_R = neck_R; io.o_com = com;
    io.o_J_lf = J_lf; io.o_J_rf = J_rf; io.o_J_neck = J_neck; io.o_J_com = J_com;
    if (!J_lf && !J_rf && !J_neck && !J_com && h->d.tpi_scratch) {
        // transforms / CoM only (the measured CoM of the closed loop): thread-per-instance kinematics, no Jacobian columns
        const cudaError_t e = launch_ik_tpi_fk(h->d, io, static_cast<cudaStream_t>(stream));
        h->launches += 1;
        if (e != cudaSuccess) { std::fprintf(stderr, "[wc_fk_dev] CUDA error: %s\n", cudaGetErrorString(e)); return WC_ERR_CUDA; }
        return WC_OK;
    }
    return launch_checked(h, io, IK_MODE_FK_ONLY, static_cast<cudaStream_t>(stream), "wc_fk_dev");
}

wc_status wc_ik_solve_with_jacobians_dev(wc_ik* h, const double* q_meas,
                         const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                         const double* J_lf, const double* J_rf, const double* J_neck, const double* J_com,
                         const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                         const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                         double* dq, double* lf_err, double* rf_err, int32_t* qp_status, int32_t* nwsr, void* stream)
{
    if (!h || !q_meas || !lf_H || !rf_H || !neck_R || !com_meas || !J_lf || !J_rf || !J_neck || !J_com || !lf_H_des ||
        !rf_H_des || !lf_twist || !rf_twist || !com_pos_des || !com_vel_des || !neck_R_des)
        return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    IkTickIO io;
    std::memset(&io, 0, sizeof(io));
    io.q_meas = q_meas; io.lf_H = lf_H; io.rf_H = rf_H; io.neck_R = neck_R; io.com_meas = com_meas;
    io.J_lf = J_lf; io.J_rf = J_rf; io.J_neck = J_neck; io.J_com = J_com;
    io.lf_H_des = lf_H_des; io.rf_H_des = rf_H_des; io.lf_twist = lf_twist; io.rf_twist = rf_twist;
    io.com_pos_des = com_pos_des; io.com_vel_des = com_vel_des; io.neck_R_des = neck_R_des;
    io.dq = dq; io.lf_err = lf_err; io.rf_err = rf_err; io.status = qp_status; io.nwsr = nwsr;
    return launch_checked(h, io, IK_MODE_EXT_KIN, static_cast<cudaStream_t>(stream), "wc_ik_solve_with_jacobians_dev");
}

wc_status wc_ik_tick_host_async(wc_ik* h, const double* q_meas, const double* q_lin, const double* base_H,
                          const uint8_t* left_is_fixed, const double* lf_H_des, const double* rf_H_des,
                          const double* lf_twist, const double* rf_twist, const double* com_pos_des,
                          const double* com_vel_des, const double* com_meas, const double* neck_R_des,
                          double* dq, double* lf_err, double* rf_err, int32_t* qp_status, int32_t* nwsr)
{
    if (!h || !q_meas || !q_lin || !base_H || !left_is_fixed || !lf_H_des || !rf_H_des || !lf_twist || !rf_twist ||
        !com_pos_des || !com_vel_des || !com_meas || !neck_R_des)
        return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t B = (size_t)h->d.B, nJ = (size_t)h->d.nJ;
    // rows of each packed double input, in order
    const size_t rows[11] = {nJ, nJ, 12, 12, 12, 6, 6, 3, 3, 3, 9};
    const double* src[11] = {q_meas, q_lin, base_H, lf_H_des, rf_H_des, lf_twist, rf_twist, com_pos_des, com_vel_des, com_meas, neck_R_des};
    size_t off[12]; off[0] = 0;
    for (int k = 0; k < 11; k++) off[k + 1] = off[k] + rows[k] * B * 8;
    const size_t o_flag = off[11];
    const size_t in_bytes = o_flag + align_up(B, 8);
    const size_t p_dq = 0, p_le = p_dq + nJ * B * 8, p_re = p_le + 6 * B * 8, p_st = p_re + 6 * B * 8;
    const size_t p_nw = p_st + align_up(B * 4, 8), out_bytes = p_nw + align_up(B * 4, 8);
    if (h->pipe.init(in_bytes, out_bytes) != cudaSuccess) return WC_ERR_ALLOC;
    // direct copies from / to the caller's (ideally pinned) buffers
    cudaError_t ce = cudaSuccess;
    const int slot = h->pipe.begin(&ce);
    if (ce != cudaSuccess) return WC_ERR_CUDA;
    cudaStream_t s = h->pipe.up;
    char* dbase = static_cast<char*>(h->pipe.dev_in[slot]);
    for (int k = 0; k < 11; k++)
        if (cudaMemcpyAsync(dbase + off[k], src[k], rows[k] * B * 8, cudaMemcpyHostToDevice, s) != cudaSuccess) return WC_ERR_CUDA;
    if (cudaMemcpyAsync(dbase + o_flag, left_is_fixed, B, cudaMemcpyHostToDevice, s) != cudaSuccess) return WC_ERR_CUDA;
    if (h->pipe.before_run(slot) != cudaSuccess) return WC_ERR_CUDA;
    const char* di = dbase;
    char* dout = static_cast<char*>(h->pipe.dev_out[slot]);
    auto D = [&](int k) { return reinterpret_cast<const double*>(di + off[k]); };
    wc_status rc = wc_ik_tick_dev(h, D(0), D(1), D(2), reinterpret_cast<const uint8_t*>(di + o_flag), D(3), D(4), D(5), D(6),
                                  D(7), D(8), D(9), D(10),
                                  reinterpret_cast<double*>(dout + p_dq), reinterpret_cast<double*>(dout + p_le),
                                  reinterpret_cast<double*>(dout + p_re), reinterpret_cast<int32_t*>(dout + p_st),
                                  reinterpret_cast<int32_t*>(dout + p_nw), h->pipe.run);
    if (rc != WC_OK) return rc;
    cudaError_t e = h->pipe.before_down(slot);
    s = h->pipe.down;
#define D2H(dst, off_, bytes) do { if ((dst) && e == cudaSuccess) e = cudaMemcpyAsync((dst), dout + (off_), (bytes), cudaMemcpyDeviceToHost, s); } while (0)
    D2H(dq, p_dq, nJ * B * 8); D2H(lf_err, p_le, 6 * B * 8); D2H(rf_err, p_re, 6 * B * 8);
    D2H(qp_status, p_st, B * 4); D2H(nwsr, p_nw, B * 4);
#undef D2H
    if (e == cudaSuccess) e = h->pipe.end(slot);
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[wc_ik_tick_host_async] CUDA error: %s\n", cudaGetErrorString(e));
        return WC_ERR_CUDA;
    }
    return WC_OK;
}

wc_status wc_ik_sync(wc_ik* h)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (!h->pipe.ready) return WC_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const cudaError_t e = h->pipe.sync_all();
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[wc_ik_sync] CUDA error: %s\n", cudaGetErrorString(e));
        return WC_ERR_CUDA;
    }
    return WC_OK;
}

wc_status wc_ik_sync_one(wc_ik* h)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (!h->pipe.ready) return WC_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    return h->pipe.sync_one() == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_ik_tick_host(wc_ik* h, const double* q_meas, const double* q_lin, const double* base_H,
                          const uint8_t* left_is_fixed, const double* lf_H_des, const double* rf_H_des,
                          const double* lf_twist, const double* rf_twist, const double* com_pos_des,
                          const double* com_vel_des, const double* com_meas, const double* neck_R_des,
                          double* dq, double* lf_err, double* rf_err, int32_t* qp_status, int32_t* nwsr)
{
    const wc_status rc = wc_ik_tick_host_async(h, q_meas, q_lin, base_H, left_is_fixed, lf_H_des, rf_H_des, lf_twist, rf_twist,
                                               com_pos_des, com_vel_des, com_meas, neck_R_des, dq, lf_err, rf_err, qp_status, nwsr);
    return rc != WC_OK ? rc : wc_ik_sync(h);
}

namespace {
// scratch arena for the small host-pointer entry points
struct Arena {
    char* base = nullptr; size_t cap = 0, used = 0;
    ~Arena() { if (base) cudaFree(base); }
    void* take(size_t bytes) { void* p = base + used; used += (bytes + 255) / 256 * 256; return p; }
};
} // namespace

wc_status wc_fk_host(wc_ik* h, const double* q, const double* base_H, const uint8_t* left_is_fixed,
                     double* lf_H, double* rf_H, double* neck_R, double* com,
                     double* J_lf, double* J_rf, double* J_neck, double* J_com)
{
    if (!h || !q || !base_H || !left_is_fixed) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t B = (size_t)h->d.B, nJ = (size_t)h->d.nJ, nV = nJ + 6;
    const size_t rows_in[2] = {nJ, 12};
    const double* src[2] = {q, base_H};
    const size_t rows_out[8] = {12, 12, 9, 3, 6 * nV, 6 * nV, 6 * nV, 3 * nV};
    double* dst[8] = {lf_H, rf_H, neck_R, com, J_lf, J_rf, J_neck, J_com};
    Arena ar;
    size_t total = 4096;
    for (size_t r : rows_in) total += r * B * 8 + 256;
    for (size_t r : rows_out) total += r * B * 8 + 256;
    total += B + 256;
    if (cudaMalloc(reinterpret_cast<void**>(&ar.base), total) != cudaSuccess) return WC_ERR_ALLOC;
    ar.cap = total;
    double* din[2]; double* dout[8];
    for (int k = 0; k < 2; k++) {
        din[k] = static_cast<double*>(ar.take(rows_in[k] * B * 8));
        if (cudaMemcpy(din[k], src[k], rows_in[k] * B * 8, cudaMemcpyHostToDevice) != cudaSuccess) return WC_ERR_CUDA;
    }
    uint8_t* dflag = static_cast<uint8_t*>(ar.take(B));
    if (cudaMemcpy(dflag, left_is_fixed, B, cudaMemcpyHostToDevice) != cudaSuccess) return WC_ERR_CUDA;
    for (int k = 0; k < 8; k++) dout[k] = dst[k] ? static_cast<double*>(ar.take(rows_out[k] * B * 8)) : nullptr;
    wc_status rc = wc_fk_dev(h, din[0], din[1], dflag, dout[0], dout[1], dout[2], dout[3], dout[4], dout[5], dout[6], dout[7], nullptr);
    if (rc != WC_OK) return rc;
    if (cudaDeviceSynchronize() != cudaSuccess) return WC_ERR_CUDA;
    for (int k = 0; k < 8; k++)
        if (dst[k] && cudaMemcpy(dst[k], dout[k], rows_out[k] * B * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return WC_ERR_CUDA;
    return WC_OK;
}

wc_status wc_ik_solve_with_jacobians_host(wc_ik* h, const double* q_meas,
                         const double* lf_H, const double* rf_H, const double* neck_R, const double* com_meas,
                         const double* J_lf, const double* J_rf, const double* J_neck, const double* J_com,
                         const double* lf_H_des, const double* rf_H_des, const double* lf_twist, const double* rf_twist,
                         const double* com_pos_des, const double* com_vel_des, const double* neck_R_des,
                         double* dq, double* lf_err, double* rf_err, int32_t* qp_status, int32_t* nwsr)
{
    if (!h || !q_meas || !lf_H || !rf_H || !neck_R || !com_meas || !J_lf || !J_rf || !J_neck || !J_com || !lf_H_des ||
        !rf_H_des || !lf_twist || !rf_twist || !com_pos_des || !com_vel_des || !neck_R_des)
        return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    const size_t B = (size_t)h->d.B, nJ = (size_t)h->d.nJ, nV = nJ + 6;
    const size_t rows_in[16] = {nJ, 12, 12, 9, 3, 6 * nV, 6 * nV, 6 * nV, 3 * nV, 12, 12, 6, 6, 3, 3, 9};
    const double* src[16] = {q_meas, lf_H, rf_H, neck_R, com_meas, J_lf, J_rf, J_neck, J_com, lf_H_des, rf_H_des,
                             lf_twist, rf_twist, com_pos_des, com_vel_des, neck_R_des};
    Arena ar;
    size_t total = 8192 + (nJ + 12) * B * 8 + 2 * B * 4 + 2048;
    for (size_t r : rows_in) total += r * B * 8 + 256;
    if (cudaMalloc(reinterpret_cast<void**>(&ar.base), total) != cudaSuccess) return WC_ERR_ALLOC;
    ar.cap = total;
    double* din[16];
    for (int k = 0; k < 16; k++) {
        din[k] = static_cast<double*>(ar.take(rows_in[k] * B * 8));
        if (cudaMemcpy(din[k], src[k], rows_in[k] * B * 8, cudaMemcpyHostToDevice) != cudaSuccess) return WC_ERR_CUDA;
    }
    double* d_dq = static_cast<double*>(ar.take(nJ * B * 8));
    double* d_le = static_cast<double*>(ar.take(6 * B * 8));
    double* d_re = static_cast<double*>(ar.take(6 * B * 8));
    int32_t* d_st = static_cast<int32_t*>(ar.take(B * 4));
    int32_t* d_nw = static_cast<int32_t*>(ar.take(B * 4));
    wc_status rc = wc_ik_solve_with_jacobians_dev(h, din[0], din[1], din[2], din[3], din[4], din[5], din[6], din[7], din[8],
                                                  din[9], din[10], din[11], din[12], din[13], din[14], din[15],
                                                  d_dq, d_le, d_re, d_st, d_nw, nullptr);
    if (rc != WC_OK) return rc;
    if (cudaDeviceSynchronize() != cudaSuccess) return WC_ERR_CUDA;
    cudaError_t e = cudaSuccess;
    if (dq && e == cudaSuccess) e = cudaMemcpy(dq, d_dq, nJ * B * 8, cudaMemcpyDeviceToHost);
    if (lf_err && e == cudaSuccess) e = cudaMemcpy(lf_err, d_le, 6 * B * 8, cudaMemcpyDeviceToHost);
    if (rf_err && e == cudaSuccess) e = cudaMemcpy(rf_err, d_re, 6 * B * 8, cudaMemcpyDeviceToHost);
    if (qp_status && e == cudaSuccess) e = cudaMemcpy(qp_status, d_st, B * 4, cudaMemcpyDeviceToHost);
    if (nwsr && e == cudaSuccess) e = cudaMemcpy(nwsr, d_nw, B * 4, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

wc_status wc_ik_enable_profiling(wc_ik* h, int32_t on)
{
    if (!h) return WC_ERR_INVALID_ARG;
    h->profiling = on ? 1 : 0;
    return WC_OK;
}

wc_status wc_ik_kernel_times_ms(wc_ik* h, double* ik_ms, int64_t* launches)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return WC_ERR_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return WC_ERR_CUDA;
    for (size_t k = 0; k + 2 < h->ev.size(); k += 3) {
        float a = 0, f = 0;
        cudaEventElapsedTime(&a, h->ev[k], h->ev[k + 2]);
        cudaEventElapsedTime(&f, h->ev[k], h->ev[k + 1]);
        h->ms_acc += a; h->fast_ms_acc += f; h->timed += 1;
    }
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    h->ev.clear();
    if (ik_ms) *ik_ms = h->ms_acc;
    if (launches) *launches = h->timed;
    h->detail[0] = h->fast_ms_acc; h->detail[1] = h->ms_acc - h->fast_ms_acc;
    h->ms_acc = 0; h->timed = 0; h->fast_ms_acc = 0;
    return WC_OK;
}

wc_status wc_ik_kernel_times_detail_ms(wc_ik* h, double* fast_ms, double* general_ms)
{
    if (!h) return WC_ERR_INVALID_ARG;
    if (fast_ms) *fast_ms = h->detail[0];
    if (general_ms) *general_ms = h->detail[1];
    return WC_OK;
}

wc_status wc_ik_path_counts_host(wc_ik* h, int32_t* second_pass, int32_t* general)
{
    if (!h) return WC_ERR_INVALID_ARG;
    int32_t c[2] = {0, 0};
    if (h->d.slow_count) {
        cudaSetDevice(h->device);
        if (cudaDeviceSynchronize() != cudaSuccess ||
            cudaMemcpy(c, h->d.slow_count, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return WC_ERR_CUDA;
    }
    if (general) *general = c[0];
    if (second_pass) *second_pass = c[1];
    return WC_OK;
}

} // extern "C"
