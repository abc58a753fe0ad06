// mpc_chain.cu — "chain-per-thread" variant of the batched OSQP-equivalent ADMM (K2); the benchmark path.
//
// Same mathematics, same iterates as mpc_admm_kernel (mpc_kernels.cu; mapping to the reference in
// mpc_kernels.cuh): OSQP's KKT solve done on the reduced system (P + sI + A' rho A) x~ = rhs, which is
// pentadiagonal per axis plus a dense 4x4 corner on (u_0, x_0) of both axes.
//
// Mapping.  The ncu profile of the cooperative kernel (profiles/r1a_*) shows it is ISSUE bound: the
// sequential triangular sweeps run on 2 of the 16 lanes of an instance, so 45 % of all warp
// instructions execute with 4 of 32 lanes active.  Here ONE THREAD owns ONE AXIS CHAIN of one
// instance: a warp holds 16 instances (lanes 0-15 the x axes, lanes 16-31 the y axes of the same 16
// instances; the pair meets by shfl_xor 16 for the corner and the hull rows).  Every lane is a
// sweeper, and because a thread walks its chain stage by stage anyway, the right-hand-side build is
// fused into the forward sweep and the z / y / x updates into the backward sweep: one pass down, one
// pass up per ADMM iteration, ~8x fewer warp instructions per instance-tick (ncu, profiles/r1a vs r1g).
// Longer horizons use the same kernels with 8, 4, 2 or 1 instances per CTA (a partial warp), whatever fits
// one CTA's shared memory; wc_mpc_create prefers them while at least 8 instances share a CTA (N <= 106).
//
// Shared memory: every per-thread array is lane-interleaved (element k of lane l at [k*32 + l]) so
// all accesses are bank-conflict free and the HBM <-> shared staging is coalesced (field-major SoA
// in HBM: consecutive instances are consecutive addresses).  Per thread (doubles):
//   XX[nR] XU[N] iterate | QX[nR] scaled q (state block; the input block is zero except u_0) |
//   WX[nR] WU[N] forward-sweep result | L1X L1U L2X L2U DIX DIU  pentadiagonal LDL' |
//   YE[nR] equality duals | CM CA CB [nR] scaled rows of A
// Per instance: HX HY UH ZH YH EH [8] hull rows.  N = 25: 385 doubles/thread -> 104.7 KB per warp,
// two single-warp CTAs per SM.
//
// The LDL' factors depend only on (D, E, c, rho), which change at osqp_setup and at a rho update,
// so they live in HBM next to the rest of the workspace (OSQP keeps its factorisation in the workspace
// too) and are staged in with it; chain_factor runs in the setup kernel and after adapt_rho only.
// The horizon is a template parameter for the benchmark configuration (N = 25): all shared-memory
// offsets fold into immediates and the stage loops unroll 6x, which lets ptxas hoist the loads of later
// stages above the dependent FMA chain of the current one.  What bounds the kernel (shared-memory pipe,
// two lone warps per SM) and what was tried against it is in DESIGN.md §5.
//
// Work distribution: instances whose OSQP workspace is new (support phase changed, ~2 %/tick) are
// compacted into a list by K1 and equilibrated by mpc_chain_setup_kernel; phase 1 runs every
// instance up to the first termination check; the few that have not converged are compacted again
// and finished by phase 2 (a small grid striding over the list).  A warp therefore never idles 15 lanes
// while one instance runs the Ruiz passes or another 25 iterations.
#include "mpc_kernels.cuh"

#include <cmath>
#include <cstdio>
#include <type_traits>

namespace wcb {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kIPWmax = 16; // instances per warp at short horizons; 8, 4, 2, 1 as the per-chain working set grows
// stages per trip of the forward / backward loops: enough straight-line code for ptxas to overlap the
// loads of later stages with the FMA chain of the current one, small enough to stay in the instruction cache
#ifndef WC_CHAIN_UNROLL
#define WC_CHAIN_UNROLL 6
#endif
constexpr int kChainUnroll = WC_CHAIN_UNROLL;

__device__ __forceinline__ double limit_scaling_c(double v)
{
    v = v < kMinScaling ? 1.0 : v;
    v = v > kMaxScaling ? kMaxScaling : v;
    return v;
}

// lane-interleaved per-thread arrays; IPW = instances per CTA (one partial warp of 2*IPW threads)
template <int IPW>
struct ChainSmem {
    double* base;  // + lane
    double* hbase; // per-instance block, + slot
    int nR, N;
    __device__ __forceinline__ double& at(int arr_off, int k) const { return base[(arr_off + k) * (2 * IPW)]; }
    __device__ __forceinline__ double& hat(int arr, int h) const { return hbase[(arr * 8 + h) * IPW]; }
};

// array offsets (in elements per thread) for horizon N
struct Off {
    int XX, XU, QX, WX, WU, L1X, L1U, L2X, L2U, DIX, DIU, YE, CM, CA, CB, total;
    __host__ __device__ explicit Off(int N)
    {
        const int nR = N + 1;
        int p = 0;
        XX = p; p += nR; XU = p; p += N; QX = p; p += nR; WX = p; p += nR; WU = p; p += N;
        L1X = p; p += nR; L1U = p; p += N; L2X = p; p += nR; L2U = p; p += N; DIX = p; p += nR; DIU = p; p += N;
        YE = p; p += nR; CM = p; p += nR; CA = p; p += nR; CB = p; p += nR;
        total = p;
    }
};
enum { H_HX = 0, H_HY, H_UH, H_ZH, H_YH, H_EH, H_RX, H_RY, H_RB, H_COUNT }; // H_R*: the unscaled hull rows (epilogue margin test)

// 8-byte asynchronous copy HBM -> shared (LDGSTS): no register staging, so a thread can keep its whole
// ~340-element stage-in in flight at once (two warps per SM cannot cover HBM latency with plain loads).
// src_bytes = 0 zero-fills the destination (cold start: x = y = z = 0).
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool nonzero)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = nonzero ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// Bulk asynchronous copies (cp.async.bulk, the TMA engine's 1-D form; SASS UBLKCP) completing on an mbarrier: one
// instruction moves the 128-byte half row that 16 lanes otherwise fetch with 16 LDGSTS and their address arithmetic
// (714 copies per CTA, 23 per lane, instead of ~340 LDGSTS per lane).  MEASURED AND SWITCHED OFF (-DWC_CHAIN_BULK=1
// builds it): parity green, but phase 1 went from 1.64 to 1.92 ms at B = 65 536 -- with a field-major layout a row
// holds only 128 contiguous bytes of one CTA's instances, and 714 requests of 128 bytes per CTA are bound by the
// copy engine's request rate, not by bytes.  What it needs is fewer, larger requests: a group-major workspace (one
// contiguous block per 16 instances) or 3-D tensor maps whose box {16 instances, 2 axes, all stages} lands as
// [stage][axis][16] = this kernel's shared-memory layout.
#ifndef WC_CHAIN_BULK
#define WC_CHAIN_BULK 0
#endif
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// 128 bytes HBM -> shared; both addresses 16-byte aligned
__device__ __forceinline__ void bulk128(unsigned dst, const double* gsrc, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];\n"
                 ::"r"(dst), "l"(gsrc), "r"(bar) : "memory");
}
// generic-proxy accesses to shared memory ordered before the async-proxy writes that follow
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// exchanges between the two axis lanes of an instance: always under the PAIR mask, so that they are
// legal inside branches that are uniform for the pair but not for the warp
__device__ __forceinline__ double pair_max(double v, unsigned pm, int dist) { return fmax(v, __shfl_xor_sync(pm, v, dist)); }

// dense 4x4 LDL' of the corner, order [u0x, u0y, x0x, x0y]; CF: L (6 strict-lower, row-major) + 4 Dinv
__device__ __forceinline__ void factor_corner_c(const double* Sx, const double* Sy, double kxy, double* CF)
{
    double A[4][4] = {{Sx[0], kxy, Sx[1], 0.0}, {kxy, Sy[0], 0.0, Sy[1]}, {Sx[1], 0.0, Sx[2], 0.0}, {0.0, Sy[1], 0.0, Sy[2]}};
    double L[4][4], Dv[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double dj = A[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) dj -= L[j][k] * L[j][k] * Dv[k];
        Dv[j] = dj;
#pragma unroll
        for (int i = j + 1; i < 4; i++) {
            double v = A[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k] * Dv[k];
            L[i][j] = v / dj;
        }
    }
    CF[0] = L[1][0]; CF[1] = L[2][0]; CF[2] = L[2][1]; CF[3] = L[3][0]; CF[4] = L[3][1]; CF[5] = L[3][2];
    CF[6] = 1.0 / Dv[0]; CF[7] = 1.0 / Dv[1]; CF[8] = 1.0 / Dv[2]; CF[9] = 1.0 / Dv[3];
}

__device__ __forceinline__ void solve_corner_c(const double* CF, double* r)
{
    double t0 = r[0];
    double t1 = r[1] - CF[0] * t0;
    double t2 = r[2] - CF[1] * t0 - CF[2] * t1;
    double t3 = r[3] - CF[3] * t0 - CF[4] * t1 - CF[5] * t2;
    t0 *= CF[6]; t1 *= CF[7]; t2 *= CF[8]; t3 *= CF[9];
    t2 -= CF[5] * t3;
    t1 -= CF[2] * t2 + CF[4] * t3;
    t0 -= CF[0] * t1 + CF[1] * t2 + CF[3] * t3;
    r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}

} // namespace

// Pentadiagonal LDL' of one axis chain (x_N, u_{N-1}, x_{N-1}, ..., x_1) + Schur complement on its
// (u_0, x_0) corner + the dense 4x4 corner factor, from (D, E, E_hull, c, rho) in HBM to (L1, L2, DI, CF)
// in HBM.  Called by both lanes of an instance (pair-uniform); `write` masks the stores.
__device__ __noinline__ void chain_factor(const MpcDev& d, int i, int a, int N, double c, double rho_now, int nh,
                                          unsigned pmask, bool write, const double* Dsrc = nullptr,
                                          const double* Esrc = nullptr, int pair_dist = 16)
{
    const size_t B = (size_t)d.B;
    const int nA = 2 * N + 1, nR = N + 1;
    const double sigma = d.st.sigma;
    const double rho_e = kRhoEqOverIneq * rho_now;
    const double cQ = c * d.Qd[(size_t)a * B + i], cR = c * d.Rd[(size_t)a * B + i];
    // D, E of this axis: HBM (stride B) or, in the setup kernel, the shared-memory copy (stride 1)
    const size_t sB = Dsrc ? (size_t)1 : B;
    const double* Dg = Dsrc ? Dsrc : d.D + (size_t)(a * nA) * B + i;
    const double* Eg = Esrc ? Esrc : d.Eeq + (size_t)(a * nR) * B + i;
    double* l1g = d.L1 + (size_t)(a * nA) * B + i;
    double* l2g = d.L2 + (size_t)(a * nA) * B + i;
    double* dig = d.DI + (size_t)(a * nA) * B + i;
    const double du0 = Dg[(size_t)(nR) * sB];
    const double du0o = __shfl_xor_sync(pmask, du0, pair_dist);
    const double dux = a == 0 ? du0 : du0o, duy = a == 0 ? du0o : du0;
    double hd = 0.0, kxy = 0.0;
    for (int h = 0; h < nh; h++) {
        const double e = d.Eh[(size_t)h * B + i];
        const double hx = e * d.hullA[(size_t)(2 * h) * B + i] * dux;
        const double hy = e * d.hullA[(size_t)(2 * h + 1) * B + i] * duy;
        const double v = (a == 0) ? hx : hy;
        hd += rho_now * v * v;
        kxy += rho_now * hx * hy;
    }
    double Dm1 = 1.0, Dm2 = 1.0, l1m1 = 0.0, e1m1 = 0.0, e2m2 = 0.0, e2m1 = 0.0;
    double ca_next = 0.0; // CA[k+1]
    for (int k = N; k >= 1; k--) {
        const double e = Eg[(size_t)(k) * sB];
        const double dx = Dg[(size_t)(k) * sB];
        const double cm = -e * dx;
        const double ca = d.a_dyn * e * Dg[(size_t)((k - 1)) * sB];
        const double cb = d.b_dyn * e * Dg[(size_t)((nR + k - 1)) * sB];
        { // x_k : d, e1 (with u_{k-1}), e2 (with x_{k-1})
            const double dj = cQ * dx * dx + sigma + rho_e * (cm * cm + (k < N ? ca_next * ca_next : 0.0));
            const double e1j = rho_e * cm * cb, e2j = rho_e * cm * ca;
            const double l2 = e2m2 / Dm2;
            const double l1 = (e1m1 - l2 * Dm2 * l1m1) / Dm1;
            const double Dj = dj - l1 * l1 * Dm1 - l2 * l2 * Dm2;
            if (write) { l1g[(size_t)k * B] = l1; l2g[(size_t)k * B] = l2; dig[(size_t)k * B] = 1.0 / Dj; }
            e2m2 = e2m1; e2m1 = e2j; e1m1 = e1j; Dm2 = Dm1; Dm1 = Dj; l1m1 = l1;
        }
        if (k >= 2) { // u_{k-1} : d, e1 (with x_{k-1}), e2 (with u_{k-2})
            const int ku = k - 1;
            const double du = Dg[(size_t)((nR + ku)) * sB];
            const double dum = Dg[(size_t)((nR + ku - 1)) * sB];
            const double sc = (ku < N - 1) ? 2.0 : 1.0;
            const double dj = cR * sc * du * du + sigma + rho_e * cb * cb;
            const double e1j = rho_e * ca * cb, e2j = -cR * du * dum;
            const double l2 = e2m2 / Dm2;
            const double l1 = (e1m1 - l2 * Dm2 * l1m1) / Dm1;
            const double Dj = dj - l1 * l1 * Dm1 - l2 * l2 * Dm2;
            if (write) { l1g[(size_t)(nR + ku) * B] = l1; l2g[(size_t)(nR + ku) * B] = l2; dig[(size_t)(nR + ku) * B] = 1.0 / Dj; }
            e2m2 = e2m1; e2m1 = e2j; e1m1 = e1j; Dm2 = Dm1; Dm1 = Dj; l1m1 = l1;
        }
        ca_next = ca;
    }
    double S3[3];
    { // corner rows: u_0 then x_0  (Dm1 = pivot of x_1, Dm2 = pivot of u_1)
        const double e1 = Eg[(size_t)(1) * sB], e0 = Eg[0];
        const double dx0 = Dg[0];
        const double cb1 = d.b_dyn * e1 * du0, ca1 = d.a_dyn * e1 * dx0, cm0 = -e0 * dx0;
        const double sc = (0 < N - 1) ? 2.0 : 1.0;
        const double dj = cR * sc * du0 * du0 + sigma + rho_e * cb1 * cb1 + hd;
        const double e1j = rho_e * ca1 * cb1; // K[u_0, x_0]
        const double l2 = e2m2 / Dm2;
        const double l1 = (e1m1 - l2 * Dm2 * l1m1) / Dm1;
        S3[0] = dj - l1 * l1 * Dm1 - l2 * l2 * Dm2;   // S11
        const double l2x = e2m1 / Dm1;                  // L[x_0, x_1]
        S3[1] = e1j - l2x * Dm1 * l1;                   // S12
        const double d0 = cQ * dx0 * dx0 + sigma + rho_e * (cm0 * cm0 + ca1 * ca1);
        S3[2] = d0 - l2x * l2x * Dm1;                   // S22
        if (write) {
            l1g[(size_t)nR * B] = l1; l2g[(size_t)nR * B] = l2; dig[(size_t)nR * B] = 0.0;
            l1g[0] = 0.0; l2g[0] = l2x; dig[0] = 0.0;
        }
    }
    double So[3], CF[10];
#pragma unroll
    for (int r = 0; r < 3; r++) So[r] = __shfl_xor_sync(pmask, S3[r], pair_dist);
    if (a == 0) factor_corner_c(S3, So, kxy, CF);
    else factor_corner_c(So, S3, kxy, CF);
    if (write && a == 0) {
#pragma unroll
        for (int r = 0; r < 10; r++) d.CF[(size_t)r * B + i] = CF[r];
    }
}

// =====================================================================================
// osqp_setup for the instances of the setup list (new MPCSolver after a support-phase change):
// scale_data = Ruiz equilibration of [[P, A'],[A, 0]] with cost normalisation (OSQP scaling.c), then
// the numeric factorisation.  ONE WARP PER INSTANCE, lanes over the stages: the ten Ruiz passes are
// ten rounds of independent per-stage norms, so a warp finishes them in a few thousand cycles where a
// single thread per chain needed 300 us for the ~2 % of instances that change phase in a tick.
// P and A are never formed: their scaled entries follow from D, E, c in closed form.
// Shared memory per warp (doubles): DD[2nA] FD[2nA] QB[2nA] EE[2nR] FE[2nR] HX HY EH FH [8].
// =====================================================================================
constexpr int kSetupWarps = 4;

__device__ __forceinline__ double warp_max_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__global__ void __launch_bounds__(kSetupWarps * 32) mpc_chain_setup_kernel(MpcDev d, int warp_doubles, int do_factor)
{
    extern __shared__ double smem_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int count = *d.setup_count;
    const int cta_warps = blockDim.x >> 5;
    for (int e = blockIdx.x * cta_warps + warp; e < count; e += gridDim.x * cta_warps) { // whole warps
    const int i = d.setup_list[e];
    const size_t B = (size_t)d.B;
    const int N = d.N, nA = 2 * N + 1, nR = N + 1;
    const wc_osqp_settings& st = d.st;
    double* DD = smem_all + (size_t)warp * warp_doubles;
    double* FD = DD + 2 * nA; double* QB = FD + 2 * nA; double* EE = QB + 2 * nA; double* FE = EE + 2 * nR;
    double* HX = FE + 2 * nR; double* HY = HX + 8; double* EH = HY + 8; double* FH = EH + 8;
    const int nh = d.nhull[i];
    const double Q[2] = {d.Qd[i], d.Qd[B + i]};
    const double R[2] = {d.Rd[i], d.Rd[B + i]};
    const int head = mpc_ring_head(d);
    for (int j = lane; j < 2 * nA; j += 32) { DD[j] = 1.0; QB[j] = 0.0; }
    for (int j = lane; j < 2 * nR; j += 32) EE[j] = 1.0;
    if (lane < kMaxHull) {
        EH[lane] = 1.0;
        HX[lane] = (lane < nh) ? d.hullA[(size_t)(2 * lane) * B + i] : 0.0;
        HY[lane] = (lane < nh) ? d.hullA[(size_t)(2 * lane + 1) * B + i] : 0.0;
    }
    __syncwarp();
    // unscaled gradient (MPCSolver::setGradient): state block from the ring, input block G u_prev
    for (int k = lane; k < nR; k += 32) {
        const int slot_k = (head + k) % nR;
        QB[k] = d.grad[(size_t)slot_k * B + i];
        QB[nA + k] = d.grad[(size_t)(nR + slot_k) * B + i];
    }
    if (lane == 0) { // G = -Theta' R~ e1 -> only the u_0 block (MPCSolver.cpp:244-245)
        QB[nR] = -(R[0] * d.uprev[i]);
        QB[nA + nR] = -(R[1] * d.uprev[B + i]);
    }
    __syncwarp();
    const double ad = fabs(d.a_dyn), bd = fabs(d.b_dyn);
    double c = 1.0;
    for (int it = 0; it < st.scaling; it++) {
        for (int k = lane; k < nR; k += 32) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const double* DX = DD + a * nA; const double* DU = DX + nR; const double* Ea = EE + a * nR;
                const double dx = DX[k];
                double nrm = fabs(c * Q[a] * dx * dx);
                nrm = amax(nrm, Ea[k] * dx);
                if (k < N) nrm = amax(nrm, ad * Ea[k + 1] * dx);
                FD[a * nA + k] = 1.0 / sqrt(limit_scaling_c(nrm));
                if (k < N) {
                    const double du = DU[k];
                    const double sc = (k < N - 1) ? 2.0 : 1.0;
                    double nu = fabs(c * sc * R[a] * du * du);
                    if (k >= 1) nu = amax(nu, c * R[a] * du * DU[k - 1]);
                    if (k < N - 1) nu = amax(nu, c * R[a] * du * DU[k + 1]);
                    nu = amax(nu, bd * Ea[k + 1] * du);
                    if (k == 0) {
                        const double* Hc = (a == 0) ? HX : HY;
                        for (int h = 0; h < nh; h++) nu = amax(nu, (EH[h] * Hc[h]) * du); // du > 0: |a| du = |a du|, same bits
                    }
                    FD[a * nA + nR + k] = 1.0 / sqrt(limit_scaling_c(nu));
                }
                double nr = Ea[k] * dx;
                if (k >= 1) {
                    nr = amax(nr, ad * Ea[k] * DX[k - 1]);
                    nr = amax(nr, bd * Ea[k] * DU[k - 1]);
                }
                FE[a * nR + k] = 1.0 / sqrt(limit_scaling_c(nr));
            }
        }
        if (lane < nh) {
            const double nrm = fmax(fabs(EH[lane] * HX[lane]) * DD[nR], fabs(EH[lane] * HY[lane]) * DD[nA + nR]);
            FH[lane] = 1.0 / sqrt(limit_scaling_c(nrm));
        }
        __syncwarp();
        for (int j = lane; j < 2 * nA; j += 32) DD[j] *= FD[j];
        for (int j = lane; j < 2 * nR; j += 32) EE[j] *= FE[j];
        if (lane < nh) EH[lane] *= FH[lane];
        __syncwarp();
        // cost normalisation: c <- c / max(mean_j ||P_j||_inf, ||q||_inf)
        double sum = 0.0, qn = 0.0;
        for (int k = lane; k < nR; k += 32) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const double* DX = DD + a * nA; const double* DU = DX + nR;
                const double dx = DX[k];
                sum += fabs(c * Q[a] * dx * dx);
                qn = amax(qn, c * dx * QB[a * nA + k]);
                if (k < N) {
                    const double du = DU[k];
                    const double sc = (k < N - 1) ? 2.0 : 1.0;
                    double nu = fabs(c * sc * R[a] * du * du);
                    if (k >= 1) nu = amax(nu, c * R[a] * du * DU[k - 1]);
                    if (k < N - 1) nu = amax(nu, c * R[a] * du * DU[k + 1]);
                    sum += nu;
                    qn = amax(qn, c * du * QB[a * nA + nR + k]);
                }
            }
        }
        sum = warp_sum_d(sum);
        qn = warp_max_d(qn);
        double c_temp = sum / (double)(2 * nA);
        c_temp = fmax(c_temp, limit_scaling_c(qn));
        c_temp = limit_scaling_c(c_temp);
        c *= 1.0 / c_temp;
    }
    // publish the scaling of the new workspace (OSQP keeps Dinv, Einv next to D, E)
    for (int j = lane; j < 2 * nA; j += 32) {
        const double v = DD[j];
        d.D[(size_t)j * B + i] = v; d.Dinv[(size_t)j * B + i] = 1.0 / v;
    }
    for (int j = lane; j < 2 * nR; j += 32) {
        const double v = EE[j];
        d.Eeq[(size_t)j * B + i] = v; d.Einv[(size_t)j * B + i] = 1.0 / v;
    }
    if (lane < kMaxHull) {
        const double v = EH[lane];
        d.Eh[(size_t)lane * B + i] = v; d.Ehinv[(size_t)lane * B + i] = 1.0 / v;
    }
    const double rho0 = fmin(fmax(st.rho, kRhoMin), kRhoMax);
    if (lane == 0) {
        d.cscale[i] = c;
        d.rho[i] = rho0;
        d.inited[i] = 2; // workspace exists, cold start pending (consumed by phase 1)
    }
    __syncwarp();
    // numeric factorisation (osqp_setup -> init_linsys_solver): the two axis chains on lanes 0 and 16
    // (the split-chain kernels, mpc_split.cu, rebuild their factors in the ADMM kernel and skip this)
    if (do_factor && (lane & 15) == 0) chain_factor(d, i, lane >> 4, N, c, rho0, nh, 0x00010001u, true, DD + (lane >> 4) * nA, EE + (lane >> 4) * nR);
    __syncwarp();
    } // list entries
}

// =====================================================================================
// ADMM.  PHASE2 == false: all instances, iterations 1 .. first termination check.
//        PHASE2 == true : the instances of the continuation list, until osqp_solve returns.
// =====================================================================================
template <bool PHASE2, int NT, int IPW>
__global__ void __launch_bounds__(2 * IPW) mpc_chain_kernel(MpcDev d, MpcTickIO io)
{
    extern __shared__ double smem_all[];
    // one (partial) warp per CTA: lanes [0, IPW) the x axes, [IPW, 2 IPW) the y axes of the same IPW instances
    constexpr int kIPW = IPW;
    constexpr unsigned WMASK = (2 * IPW == 32) ? 0xffffffffu : ((1u << (2 * IPW)) - 1u);
    const int lane = threadIdx.x;
    const int slot = lane % IPW, a = lane / IPW;
    const unsigned pmask = (1u << slot) | (1u << (slot + IPW));
    const size_t B = (size_t)d.B;
    const int N = NT > 0 ? NT : d.N, nA = 2 * N + 1, nR = N + 1; // NT > 0: compile-time horizon
    const Off O(N);
    if (!PHASE2 && blockIdx.x == 0 && lane == 0) *d.setup_count = 0; // consumed by the setup kernel of this tick
    // Phase 1: one group of IPW instances per CTA.  Phase 2: a small grid strides over the (few) groups of the
    // continuation list, instead of thousands of CTAs whose only work is to find the list empty.
    const int count = PHASE2 ? *d.cont_count : d.B;
    const int ngroups = (count + kIPW - 1) / kIPW;
    ChainSmem<IPW> s{smem_all + lane, smem_all + (size_t)O.total * (2 * IPW) + slot, nR, N};
#define A_(o, k) s.at(O.o, (k))
#define H_(arr, h) s.hat((arr), (h))
    // bulk staging: phase 1 at 16 instances per CTA, every group full and its rows 16-byte aligned
    constexpr bool BULK = WC_CHAIN_BULK && !PHASE2 && IPW == 16;
    const bool bulk = BULK && (d.B % 16) == 0;
    const unsigned sbase = smem_u32(smem_all);
    const unsigned sbar = smem_u32(smem_all + (size_t)O.total * (2 * IPW) + (size_t)H_COUNT * 8 * IPW);
    unsigned bar_parity = 0;
    if (bulk) {
        if (lane == 0) mbar_init(sbar, 1);
        __syncwarp(WMASK);
    }
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int li = grp * kIPW + slot;
    const bool valid = li < count;
    const int i = PHASE2 ? d.cont_list[valid ? li : count - 1] : (valid ? li : d.B - 1);

    const wc_osqp_settings& st = d.st;
    const double sigma = st.sigma, alpha = st.alpha, infty = st.infty;
    const int kflags = d.tick_flags[i];
    const bool skip = !valid || (kflags & (WC_MPC_FLAG_NO_CONTACT | WC_MPC_FLAG_NO_SOLVER)) != 0;
    const int nh = skip ? 0 : d.nhull[i];
    const double Qa = d.Qd[(size_t)a * B + i], Ra = d.Rd[(size_t)a * B + i];
    const int init_state = d.inited[i];
    const bool cold = !PHASE2 && !skip && init_state == 2; // new OSQP workspace: x = y = z = 0
    const double c = d.cscale[i];
    double rho = d.rho[i];
    const double upa = d.uprev[(size_t)a * B + i];
    const int head = mpc_ring_head(d);
    const double cQ = c * Qa, cR = c * Ra;

    // ------------------------------------------------------------------ stage in (all copies in flight at once)
    auto load_factor = [&]() {
        const double* l1 = d.L1 + (size_t)(a * nA) * B + i;
        const double* l2 = d.L2 + (size_t)(a * nA) * B + i;
        const double* di = d.DI + (size_t)(a * nA) * B + i;
#pragma unroll
        for (int k = 0; k < nR; k++) {
            cp_async8(&A_(L1X, k), l1 + (size_t)k * B, true);
            cp_async8(&A_(L2X, k), l2 + (size_t)k * B, true);
            cp_async8(&A_(DIX, k), di + (size_t)k * B, true);
            if (k < N) {
                cp_async8(&A_(L1U, k), l1 + (size_t)(nR + k) * B, true);
                cp_async8(&A_(L2U, k), l2 + (size_t)(nR + k) * B, true);
                cp_async8(&A_(DIU, k), di + (size_t)(nR + k) * B, true);
            }
        }
    };
    // half row (16 instances of one axis) of array element `el` <- row `grow` of a field-major HBM array
    const size_t i0 = (size_t)grp * kIPW;
    auto brow = [&](int el, int aa, const double* gbase, int grow) {
        bulk128(sbase + (unsigned)((el * 32 + aa * 16) * 8), gbase + (size_t)grow * B + i0, sbar);
    };
    if (bulk) {
        fence_async_smem();
        if (lane == 0) mbar_expect_tx(sbar, (unsigned)(((8 * nR + 5 * N) * 2 + 6 * kMaxHull) * 128));
        __syncwarp(WMASK);
#pragma unroll
        for (int p = lane; p < 2 * nR; p += 32) { // lane <-> (stage k, axis aa); same parking as the LDGSTS path below
            const int k = p >> 1, aa = p & 1;
            const int slot_k = head + k < nR ? head + k : head + k - nR;
            brow(O.XX + k, aa, d.x, aa * nA + k);
            brow(O.YE + k, aa, d.yeq, aa * nR + k);
            brow(O.QX + k, aa, d.grad, aa * nR + slot_k);
            brow(O.CM + k, aa, d.Eeq, aa * nR + k);
            brow(O.WX + k, aa, d.D, aa * nA + k);
            brow(O.L1X + k, aa, d.L1, aa * nA + k);
            brow(O.L2X + k, aa, d.L2, aa * nA + k);
            brow(O.DIX + k, aa, d.DI, aa * nA + k);
            if (k < N) {
                brow(O.XU + k, aa, d.x, aa * nA + nR + k);
                brow(O.WU + k, aa, d.D, aa * nA + nR + k);
                brow(O.L1U + k, aa, d.L1, aa * nA + nR + k);
                brow(O.L2U + k, aa, d.L2, aa * nA + nR + k);
                brow(O.DIU + k, aa, d.DI, aa * nA + nR + k);
            }
        }
        for (int cidx = lane; cidx < 6 * kMaxHull; cidx += 32) { // per-instance hull block: rows of 16 doubles
            const int which = cidx / kMaxHull, h = cidx % kMaxHull;
            const double* src = which == 0 ? d.Eh + (size_t)h * B : which == 1 ? d.hullA + (size_t)(2 * h) * B
                              : which == 2 ? d.hullA + (size_t)(2 * h + 1) * B : which == 3 ? d.hullb + (size_t)h * B
                              : which == 4 ? d.yh + (size_t)h * B : d.zh + (size_t)h * B;
            const int harr = which == 0 ? H_EH : which == 1 ? H_RX : which == 2 ? H_RY : which == 3 ? H_RB : which == 4 ? H_YH : H_ZH;
            bulk128(sbase + (unsigned)((O.total * (2 * IPW) + (harr * 8 + h) * IPW) * 8), src + i0, sbar);
        }
    } else {
        const double* Dg = d.D + (size_t)(a * nA) * B + i;
        const double* Eg = d.Eeq + (size_t)(a * nR) * B + i;
        const double* xg = d.x + (size_t)(a * nA) * B + i;
        const double* yg = d.yeq + (size_t)(a * nR) * B + i;
        const double* gg = d.grad + (size_t)(a * nR) * B + i;
        // raw data: E parked in CM, D in WX / WU, the unscaled gradient in QX
#pragma unroll
        for (int k = 0; k < nR; k++) {
            const int slot_k = head + k < nR ? head + k : head + k - nR;
            cp_async8(&A_(XX, k), xg + (size_t)k * B, !cold);
            cp_async8(&A_(YE, k), yg + (size_t)k * B, !cold);
            cp_async8(&A_(QX, k), gg + (size_t)slot_k * B, true);
            cp_async8(&A_(CM, k), Eg + (size_t)k * B, true);
            cp_async8(&A_(WX, k), Dg + (size_t)k * B, true);
            if (k < N) {
                cp_async8(&A_(XU, k), xg + (size_t)(nR + k) * B, !cold);
                cp_async8(&A_(WU, k), Dg + (size_t)(nR + k) * B, true);
            }
        }
        load_factor();
        if (a == 0) {
#pragma unroll
            for (int h = 0; h < kMaxHull; h++) {
                cp_async8(&H_(H_EH, h), d.Eh + (size_t)h * B + i, true);
                cp_async8(&H_(H_RX, h), d.hullA + (size_t)(2 * h) * B + i, true);
                cp_async8(&H_(H_RY, h), d.hullA + (size_t)(2 * h + 1) * B + i, true);
                cp_async8(&H_(H_RB, h), d.hullb + (size_t)h * B + i, true);
                cp_async8(&H_(H_YH, h), d.yh + (size_t)h * B + i, !cold);
                cp_async8(&H_(H_ZH, h), d.zh + (size_t)h * B + i, !cold);
            }
        }
    }
    double CF[10];
#pragma unroll
    for (int r = 0; r < 10; r++) CF[r] = d.CF[(size_t)r * B + i];
    const double dcm_a = io.dcm[(size_t)a * B + i];
    const double z0_in = d.z0[(size_t)a * B + i];
    cp_async_wait_all();
    if (bulk) {
        mbar_wait(sbar, bar_parity); bar_parity ^= 1;
        if (cold) { // new OSQP workspace: x = y = z = 0 (the LDGSTS path zero-fills instead)
#pragma unroll
            for (int k = 0; k < nR; k++) { A_(XX, k) = 0.0; A_(YE, k) = 0.0; if (k < N) A_(XU, k) = 0.0; }
            if (a == 0) {
#pragma unroll
                for (int h = 0; h < kMaxHull; h++) { H_(H_YH, h) = 0.0; H_(H_ZH, h) = 0.0; }
            }
        }
        __syncwarp(WMASK);
    }
    // scaled problem data from the raw copies: q (osqp_update_lin_cost), rows of A, bounds
    const double dx0_scale = A_(WX, 0), du0_scale = A_(WU, 0); // D of x_0 / u_0
    const double e0 = A_(CM, 0);
#pragma unroll
    for (int k = nR - 1; k >= 0; k--) {
        const double e = A_(CM, k), dx = A_(WX, k);
        A_(CA, k) = (k >= 1) ? d.a_dyn * e * A_(WX, k - 1) : 0.0;
        A_(CB, k) = (k >= 1) ? d.b_dyn * e * A_(WU, k - 1) : 0.0;
        A_(CM, k) = -e * dx;
        A_(QX, k) = (A_(QX, k) * dx) * c;
    }
    const double qu0 = ((-(Ra * upa)) * du0_scale) * c; // scaled q of u_0; the other input entries are zero
    const double du0_other = __shfl_xor_sync(pmask, du0_scale, IPW);
    if (a == 0) {
        const double dux = du0_scale, duy = du0_other;
#pragma unroll
        for (int h = 0; h < kMaxHull; h++) {
            const bool on = h < nh;
            const double e = H_(H_EH, h);
            H_(H_HX, h) = on ? e * H_(H_RX, h) * dux : 0.0;
            H_(H_HY, h) = on ? e * H_(H_RY, h) * duy : 0.0;
            H_(H_UH, h) = on ? e * H_(H_RB, h) : 0.0;
        }
    }
    // equality row 0: l = u = -dcm (MPCSolver.cpp:143-146), scaled by E
    const double lb0 = -dcm_a * e0;
    double z0prev = (cold || PHASE2) ? (PHASE2 ? lb0 : 0.0) : z0_in;
    __syncwarp(WMASK);

    // ------------------------------------------------------------------ ADMM (osqp_solve)
    int status = WC_OSQP_UNSOLVED;
    int iter_done = 0, rho_updates = PHASE2 ? d.rho_updates[i] : 0;
    bool done = skip;
    double pri_res = 0.0, dua_res = 0.0;
    double nz_u = 0, nAx_u = 0, nq_u = 0, nAty_u = 0, nPx_u = 0;
    double rp_s = 0, rd_s = 0, nz_s = 0, nAx_s = 0, nq_s = 0, nAty_s = 0, nPx_s = 0;
    const double cinv = 1.0 / c;
    double* dyg = d.dy_scratch + (size_t)(a * nR) * B + i;      // delta_y of the equality rows (checks only)
    double* dyhg = d.dy_scratch + (size_t)(2 * nR) * B + i;     // delta_y of the hull rows

    // (A' y)_j and (P x)_j of stage k for this axis; V = vector in the role of x (VX/VU array offsets),
    // Y in the role of y (shared array offset or global pointer)
    auto aty_hull = [&](const double* yh_glob, bool from_glob) {
        double acc = 0.0;
        for (int h = 0; h < nh; h++) {
            const double hc = (a == 0) ? H_(H_HX, h) : H_(H_HY, h);
            acc += hc * (from_glob ? yh_glob[(size_t)h * B] : H_(H_YH, h));
        }
        return acc;
    };

    // update_info + the norms of check_termination for the CURRENT (x, z, y)
    // SC: also the norms of the SCALED residuals (compute_rho_estimate); only an adaptive-rho iteration needs them
    auto update_info = [&](auto sc_tag) {
        constexpr bool SC = decltype(sc_tag)::value;
        double m_rp_u = 0, m_rp_s = 0, m_z_u = 0, m_z_s = 0, m_ax_u = 0, m_ax_s = 0;
        double m_rd_u = 0, m_rd_s = 0, m_q_u = 0, m_q_s = 0, m_aty_u = 0, m_aty_s = 0, m_px_u = 0, m_px_s = 0;
        const double* Dg = d.D + (size_t)(a * nA) * B + i;
        const double* Dig = d.Dinv + (size_t)(a * nA) * B + i;
        const double* Eig = d.Einv + (size_t)(a * nR) * B + i;
        if (!PHASE2) {
            // Phase 1 evaluates the residuals exactly once, after its last iteration (phase1_len is the
            // first check / rho update / max_iter): the factor arrays are dead (a rho update reloads
            // them), so D, Dinv, Einv land there with every copy in flight at once.
            if (bulk) {
                __syncwarp(WMASK); // every lane is done reading the factor rows
                fence_async_smem();
                if (lane == 0) mbar_expect_tx(sbar, (unsigned)((3 * nR + 2 * N) * 2 * 128));
                __syncwarp(WMASK);
#pragma unroll
                for (int p = lane; p < 2 * nR; p += 32) {
                    const int k = p >> 1, aa = p & 1;
                    brow(O.L1X + k, aa, d.D, aa * nA + k);
                    brow(O.L2X + k, aa, d.Dinv, aa * nA + k);
                    brow(O.DIX + k, aa, d.Einv, aa * nR + k);
                    if (k < N) {
                        brow(O.L1U + k, aa, d.D, aa * nA + nR + k);
                        brow(O.L2U + k, aa, d.Dinv, aa * nA + nR + k);
                    }
                }
                mbar_wait(sbar, bar_parity); bar_parity ^= 1;
            } else {
#pragma unroll
            for (int k = 0; k < nR; k++) {
                cp_async8(&A_(L1X, k), Dg + (size_t)k * B, true);
                cp_async8(&A_(L2X, k), Dig + (size_t)k * B, true);
                cp_async8(&A_(DIX, k), Eig + (size_t)k * B, true);
                if (k < N) {
                    cp_async8(&A_(L1U, k), Dg + (size_t)(nR + k) * B, true);
                    cp_async8(&A_(L2U, k), Dig + (size_t)(nR + k) * B, true);
                }
            }
            cp_async_wait_all();
            }
        }
        auto DXv = [&](int k) { return PHASE2 ? Dg[(size_t)k * B] : A_(L1X, k); };
        auto DUv = [&](int k) { return PHASE2 ? Dg[(size_t)(nR + k) * B] : A_(L1U, k); };
        auto DIXv = [&](int k) { return PHASE2 ? Dig[(size_t)k * B] : A_(L2X, k); };
        auto DIUv = [&](int k) { return PHASE2 ? Dig[(size_t)(nR + k) * B] : A_(L2U, k); };
        auto EIv = [&](int k) { return PHASE2 ? Eig[(size_t)k * B] : A_(DIX, k); };
        double xx_prev = 0.0, xu_prev = 0.0;     // X of stage k-1
        double du_m1 = 0.0, vu_m1 = 0.0;         // DU, XU of stage k-1
        double du_k = DUv(0), vu_k = A_(XU, 0);
#pragma unroll 2
        for (int k = 0; k < nR; k++) {
            const double xxk = A_(XX, k);
            const double cm = A_(CM, k);
            double ax = cm * xxk;
            if (k >= 1) ax += A_(CA, k) * xx_prev + A_(CB, k) * xu_prev;
            const double z = (k == 0) ? lb0 : 0.0;
            const double einv = EIv(k);
            const double r = ax - z;
            if (SC) m_rp_s = amax(m_rp_s, r); m_rp_u = amax(m_rp_u, einv * r);
            if (SC) m_z_s = amax(m_z_s, z); m_z_u = amax(m_z_u, einv * z);
            if (SC) m_ax_s = amax(m_ax_s, ax); m_ax_u = amax(m_ax_u, einv * ax);
            const double yk = A_(YE, k);
            const double yk1 = (k < N) ? A_(YE, k + 1) : 0.0;
            {
                const double dx = DXv(k);
                const double px_x = cQ * dx * dx * xxk;
                const double aty_x = cm * yk + (k < N ? A_(CA, k + 1) * yk1 : 0.0);
                const double q = A_(QX, k), dinv = DIXv(k);
                const double rr = q + px_x + aty_x;
                if (SC) m_rd_s = amax(m_rd_s, rr); m_rd_u = amax(m_rd_u, dinv * rr);
                if (SC) m_q_s = amax(m_q_s, q); m_q_u = amax(m_q_u, dinv * q);
                if (SC) m_aty_s = amax(m_aty_s, aty_x); m_aty_u = amax(m_aty_u, dinv * aty_x);
                if (SC) m_px_s = amax(m_px_s, px_x); m_px_u = amax(m_px_u, dinv * px_x);
            }
            if (k < N) {
                const double du_p1 = (k < N - 1) ? DUv(k + 1) : 0.0;
                const double vu_p1 = (k < N - 1) ? A_(XU, k + 1) : 0.0;
                const double sc = (k < N - 1) ? 2.0 : 1.0;
                double t = sc * du_k * vu_k;
                if (k >= 1) t -= du_m1 * vu_m1;
                if (k < N - 1) t -= du_p1 * vu_p1;
                const double px_u = cR * du_k * t;
                double aty_u = A_(CB, k + 1) * yk1;
                if (k == 0) aty_u += aty_hull(nullptr, false);
                const double q = (k == 0) ? qu0 : 0.0, dinv = DIUv(k);
                const double rr = q + px_u + aty_u;
                if (SC) m_rd_s = amax(m_rd_s, rr); m_rd_u = amax(m_rd_u, dinv * rr);
                if (SC) m_q_s = amax(m_q_s, q); m_q_u = amax(m_q_u, dinv * q);
                if (SC) m_aty_s = amax(m_aty_s, aty_u); m_aty_u = amax(m_aty_u, dinv * aty_u);
                if (SC) m_px_s = amax(m_px_s, px_u); m_px_u = amax(m_px_u, dinv * px_u);
                xu_prev = vu_k;
                du_m1 = du_k; vu_m1 = vu_k; du_k = du_p1; vu_k = vu_p1;
            }
            xx_prev = xxk;
        }
        // hull rows: split between the two lanes of the pair
        {
            const double xu0 = A_(XU, 0);
            const double xu0o = __shfl_xor_sync(pmask, xu0, IPW);
            const double ux = a == 0 ? xu0 : xu0o, uy = a == 0 ? xu0o : xu0;
            for (int h = a; h < nh; h += 2) {
                const double ax = H_(H_HX, h) * ux + H_(H_HY, h) * uy;
                const double z = H_(H_ZH, h);
                const double einv = 1.0 / H_(H_EH, h); // = Ehinv of osqp_setup
                const double r = ax - z;
                if (SC) m_rp_s = amax(m_rp_s, r); m_rp_u = amax(m_rp_u, einv * r);
                if (SC) m_z_s = amax(m_z_s, z); m_z_u = amax(m_z_u, einv * z);
                if (SC) m_ax_s = amax(m_ax_s, ax); m_ax_u = amax(m_ax_u, einv * ax);
            }
        }
        pri_res = pair_max(m_rp_u, pmask, IPW); if (SC) rp_s = pair_max(m_rp_s, pmask, IPW);
        nz_u = pair_max(m_z_u, pmask, IPW); if (SC) nz_s = pair_max(m_z_s, pmask, IPW);
        nAx_u = pair_max(m_ax_u, pmask, IPW); if (SC) nAx_s = pair_max(m_ax_s, pmask, IPW);
        dua_res = cinv * pair_max(m_rd_u, pmask, IPW); if (SC) rd_s = pair_max(m_rd_s, pmask, IPW);
        nq_u = cinv * pair_max(m_q_u, pmask, IPW); if (SC) nq_s = pair_max(m_q_s, pmask, IPW);
        nAty_u = cinv * pair_max(m_aty_u, pmask, IPW); if (SC) nAty_s = pair_max(m_aty_s, pmask, IPW);
        nPx_u = cinv * pair_max(m_px_u, pmask, IPW); if (SC) nPx_s = pair_max(m_px_s, pmask, IPW);
    };

    // is_primal_infeasible (OSQP auxil.c) on delta_y (global scratch written by the last iteration)
    auto primal_infeasible = [&](double eps) -> int {
        const double* Eg = d.Eeq + (size_t)(a * nR) * B + i;
        double nrm = 0.0, lhs = 0.0;
        for (int k = 0; k < nR; k++) {
            const double dy = dyg[(size_t)k * B];
            nrm = amax(nrm, Eg[(size_t)k * B] * dy);
            const double bnd = (k == 0) ? lb0 : 0.0;
            lhs += bnd * fmax(dy, 0.0) + bnd * fmin(dy, 0.0);
        }
        if (a == 0) {
            // project delta_y of the hull rows on the polar of the recession cone (l = -inf -> max(dy, 0))
            for (int h = 0; h < nh; h++) {
                double dy = dyhg[(size_t)h * B];
                const double e = H_(H_EH, h);
                const double uh = H_(H_UH, h), lh = e * (-infty);
                const bool u_inf = uh > infty * kMinScaling, l_inf = lh < -infty * kMinScaling;
                if (u_inf) dy = l_inf ? 0.0 : fmin(dy, 0.0);
                else if (l_inf) dy = fmax(dy, 0.0);
                dyhg[(size_t)h * B] = dy;
                nrm = amax(nrm, e * dy);
                lhs += uh * fmax(dy, 0.0) + lh * fmin(dy, 0.0);
            }
        }
        nrm = pair_max(nrm, pmask, IPW);
        {
            const double lo = __shfl_xor_sync(pmask, lhs, IPW);
            lhs = a == 0 ? lhs + lo : lo + lhs;
        }
        __syncwarp(WMASK);
        int res = 0;
        if (nrm > 1.0 / infty && lhs < -eps * nrm) {
            const double* Dig = d.Dinv + (size_t)(a * nA) * B + i;
            double m = 0.0;
            for (int k = 0; k < nR; k++) {
                const double yk = dyg[(size_t)k * B];
                const double yk1 = (k < N) ? dyg[(size_t)(k + 1) * B] : 0.0;
                const double aty_x = A_(CM, k) * yk + (k < N ? A_(CA, k + 1) * yk1 : 0.0);
                m = amax(m, aty_x * Dig[(size_t)k * B]); // note: cooperative kernel divides by D
                if (k < N) {
                    double aty_u = A_(CB, k + 1) * yk1;
                    if (k == 0) aty_u += aty_hull(dyhg, true);
                    m = amax(m, aty_u * Dig[(size_t)(nR + k) * B]);
                }
            }
            m = pair_max(m, pmask, IPW);
            res = m < eps * nrm;
        }
        return res;
    };
    // is_dual_infeasible on delta_x (left in WX / WU by the last iteration)
    auto dual_infeasible = [&](double eps) -> int {
        const double* Dg = d.D + (size_t)(a * nA) * B + i;
        const double* Dig = d.Dinv + (size_t)(a * nA) * B + i;
        const double* Eig = d.Einv + (size_t)(a * nR) * B + i;
        double nrm = 0.0, qdx = 0.0;
        for (int k = 0; k < nR; k++) {
            const double wx = A_(WX, k);
            nrm = amax(nrm, Dg[(size_t)k * B] * wx);
            qdx += A_(QX, k) * wx;
        }
        for (int k = 0; k < N; k++) {
            const double wu = A_(WU, k);
            nrm = amax(nrm, Dg[(size_t)(nR + k) * B] * wu);
            if (k == 0) qdx += qu0 * wu;
        }
        nrm = pair_max(nrm, pmask, IPW);
        {
            const double qo = __shfl_xor_sync(pmask, qdx, IPW);
            qdx = a == 0 ? qdx + qo : qo + qdx;
        }
        int res = 0;
        if (nrm > 1.0 / infty && qdx < -c * eps * nrm) {
            double m = 0.0;
            for (int k = 0; k < nR; k++) {
                const double dx = Dg[(size_t)k * B];
                const double px_x = cQ * dx * dx * A_(WX, k);
                m = amax(m, px_x * Dig[(size_t)k * B]);
                if (k < N) {
                    const double sc = (k < N - 1) ? 2.0 : 1.0;
                    const double du = Dg[(size_t)(nR + k) * B];
                    double t = sc * du * A_(WU, k);
                    if (k >= 1) t -= Dg[(size_t)(nR + k - 1) * B] * A_(WU, k - 1);
                    if (k < N - 1) t -= Dg[(size_t)(nR + k + 1) * B] * A_(WU, k + 1);
                    const double px_u = cR * du * t;
                    m = amax(m, px_u * Dig[(size_t)(nR + k) * B]);
                }
            }
            m = pair_max(m, pmask, IPW);
            if (m < c * eps * nrm) {
                int bad = 0;
                for (int k = 0; k < nR; k++) {
                    double ax = A_(CM, k) * A_(WX, k);
                    if (k >= 1) ax += A_(CA, k) * A_(WX, k - 1) + A_(CB, k) * A_(WU, k - 1);
                    ax *= Eig[(size_t)k * B];
                    if (ax > eps * nrm || ax < -eps * nrm) bad = 1; // both bounds finite on equality rows
                }
                {
                    const double wu0 = A_(WU, 0);
                    const double wu0o = __shfl_xor_sync(pmask, wu0, IPW);
                    const double ux = a == 0 ? wu0 : wu0o, uy = a == 0 ? wu0o : wu0;
                    for (int h = a; h < nh; h += 2) {
                        double ax = H_(H_HX, h) * ux + H_(H_HY, h) * uy;
                        ax *= d.Ehinv[(size_t)h * B + i];
                        const double e = H_(H_EH, h);
                        const double uh = H_(H_UH, h), lh = e * (-infty);
                        if ((uh < infty * kMinScaling && ax > eps * nrm) || (lh > -infty * kMinScaling && ax < -eps * nrm)) bad = 1;
                    }
                }
                bad = (int)pair_max((double)bad, pmask, IPW);
                res = !bad;
            }
        }
        return res;
    };
    // check_termination (OSQP auxil.c).  Called by the whole warp (the infeasibility tests contain pair
    // shuffles); returns the status the solve stops with, or 0 to continue.  The caller applies it
    // only to lanes that are still iterating.
    auto check_termination = [&](int approximate) -> int {
        double eps_abs = st.eps_abs, eps_rel = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
        const bool non_cvx = pri_res > infty || dua_res > infty;
        if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
        int prc = 0, drc = 0, prim_inf = 0, dual_inf = 0;
        const double eps_prim = eps_abs + eps_rel * fmax(nz_u, nAx_u);
        if (pri_res < eps_prim) prc = 1;
        const double eps_dual = eps_abs + eps_rel * fmax(fmax(nq_u, nAty_u), nPx_u);
        if (dua_res < eps_dual) drc = 1;
        const bool live = !done && !non_cvx;
        const bool need_pi = live && !prc, need_di = live && !drc;
        if (__any_sync(WMASK, need_pi)) { const int r = primal_infeasible(epi); if (need_pi) prim_inf = r; }
        if (__any_sync(WMASK, need_di)) { const int r = dual_infeasible(edi); if (need_di) dual_inf = r; }
        if (non_cvx) return WC_OSQP_NON_CVX;
        if (prc && drc) return approximate ? WC_OSQP_SOLVED_INACCURATE : WC_OSQP_SOLVED;
        if (prim_inf) return approximate ? WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE : WC_OSQP_PRIMAL_INFEASIBLE;
        if (dual_inf) return approximate ? WC_OSQP_DUAL_INFEASIBLE_INACCURATE : WC_OSQP_DUAL_INFEASIBLE;
        return 0;
    };

    const int max_iter = st.max_iter;
    const int chk = st.check_termination;
    const int rint = d.rho_interval;
    const bool adaptive = st.adaptive_rho && rint > 0;
    // phase 1 stops after this many iterations (first termination check); 0 = run to the end
    int phase1_len = chk > 0 ? chk : max_iter;
    if (adaptive && rint < phase1_len) phase1_len = rint;
    if (phase1_len > max_iter) phase1_len = max_iter;
    int iter = PHASE2 ? phase1_len : 0;
    int to_check = chk ? chk - (iter % chk) : 0;
    int to_adapt = adaptive ? rint - (iter % rint) : 0;
    // 1 / rho_e, 1 / rho: divisions hoisted out of the iteration (rho changes in adapt_rho only)
    double rho_e_inv = 1.0 / (kRhoEqOverIneq * rho), rho_inv = 1.0 / rho;
    bool warp_busy = __any_sync(WMASK, !done);
    while (warp_busy) {
        iter++;
        const bool act = !done;
        const double rho_e = kRhoEqOverIneq * rho;
        bool check_now = false, adapt_now = false;
        if (chk) { if (--to_check == 0) { check_now = true; to_check = chk; } }
        if (adaptive) { if (--to_adapt == 0) { adapt_now = true; to_adapt = rint; } }
        const bool keep_delta = check_now || iter >= max_iter; // delta_x / delta_y feed the infeasibility tests only
        double xt_u0 = 0.0, xt_u0o = 0.0;
        // ---- forward: right-hand side sigma x - q + A' (rho z - y) fused with L t = rhs
        if (act) {
            double t1 = 0.0, t2 = 0.0; // t at the previous two chain positions
            double w_next = 0.0, ca_next = 0.0;
#pragma unroll kChainUnroll
            for (int k = N; k >= 2; k--) {
                const double wk = rho_e * (0.0 - rho_e_inv * A_(YE, k));
                double rx = sigma * A_(XX, k) - A_(QX, k);
                rx += A_(CM, k) * wk;
                if (k < N) rx += ca_next * w_next;
                const double tx = (rx - A_(L2X, k) * t2) - A_(L1X, k) * t1;
                A_(WX, k) = tx;
                double ru = sigma * A_(XU, k - 1);
                ru += A_(CB, k) * wk;
                const double tu = (ru - A_(L2U, k - 1) * t1) - A_(L1U, k - 1) * tx;
                A_(WU, k - 1) = tu;
                t2 = tx; t1 = tu;
                w_next = wk; ca_next = A_(CA, k);
            }
            {
                const double w1 = rho_e * (0.0 - rho_e_inv * A_(YE, 1));
                double rx = sigma * A_(XX, 1) - A_(QX, 1);
                rx += A_(CM, 1) * w1;
                if (1 < N) rx += ca_next * w_next;
                const double tx = (rx - A_(L2X, 1) * t2) - A_(L1X, 1) * t1; // x_1
                A_(WX, 1) = tx;
                t2 = t1; t1 = tx;
                // corner right-hand sides: u_0 then x_0
                double ru = sigma * A_(XU, 0) - qu0;
                ru += A_(CB, 1) * w1;
                for (int h = 0; h < nh; h++) {
                    const double hc = (a == 0) ? H_(H_HX, h) : H_(H_HY, h);
                    ru += hc * (rho * (H_(H_ZH, h) - rho_inv * H_(H_YH, h)));
                }
                const double w0 = rho_e * (z0prev - rho_e_inv * A_(YE, 0));
                double r0 = sigma * A_(XX, 0) - A_(QX, 0);
                r0 += A_(CM, 0) * w0;
                r0 += A_(CA, 1) * w1;
                xt_u0 = (ru - A_(L2U, 0) * t2) - A_(L1U, 0) * t1;
                xt_u0o = r0 - A_(L2X, 0) * t1; // reduced rhs of x_0 (parked here until the corner solve)
            }
        }
        // ---- corner: dense 4x4 on [u0x, u0y, x0x, x0y], redundantly on both lanes of the pair
        double xt_x0;
        {
            const double ou = __shfl_xor_sync(pmask, xt_u0, IPW), ox = __shfl_xor_sync(pmask, xt_u0o, IPW);
            double r[4];
            if (a == 0) { r[0] = xt_u0; r[1] = ou; r[2] = xt_u0o; r[3] = ox; }
            else { r[0] = ou; r[1] = xt_u0; r[2] = ox; r[3] = xt_u0o; }
            solve_corner_c(CF, r);
            xt_u0 = (a == 0) ? r[0] : r[1];
            xt_x0 = (a == 0) ? r[2] : r[3];
            // ---- hull rows (z~ = A x~): update_z / update_y, rows split between the two lanes
            __syncwarp(WMASK); // all reads of ZH / YH by the forward pass are done
            if (act) {
                for (int h = a; h < nh; h += 2) {
                    const double zt = H_(H_HX, h) * r[0] + H_(H_HY, h) * r[1];
                    const double zp = H_(H_ZH, h), y = H_(H_YH, h);
                    const double zr = alpha * zt + (1.0 - alpha) * zp;
                    double zn = zr + rho_inv * y;
                    zn = fmax(zn, H_(H_EH, h) * (-infty));
                    zn = fmin(zn, H_(H_UH, h));
                    const double dy = rho * (zr - zn);
                    H_(H_ZH, h) = zn;
                    H_(H_YH, h) = y + dy;
                    if (keep_delta) dyhg[(size_t)h * B] = dy;
                }
            }
        }
        // ---- backward: L' x~ = D^-1 t fused with update_y (equality rows) and update_x
        auto backward = [&](auto keep_tag) {
            constexpr bool KEEP = decltype(keep_tag)::value; // delta_x / delta_y wanted (checks only)
            double a1 = xt_u0, a2 = xt_x0; // x~ at chain positions j+1, j+2
            { // stage 0
                const double zt = A_(CM, 0) * xt_x0;
                const double dy = rho_e * (alpha * zt + (1.0 - alpha) * z0prev - lb0);
                A_(YE, 0) += dy;
                if (KEEP) dyg[0] = dy;
                const double xp = A_(XX, 0), xn = alpha * xt_x0 + (1.0 - alpha) * xp;
                A_(XX, 0) = xn;
                const double up = A_(XU, 0), un = alpha * xt_u0 + (1.0 - alpha) * up;
                A_(XU, 0) = un;
                if (KEEP) { A_(WX, 0) = xn - xp; A_(WU, 0) = un - up; }
            }
            double xx_prev = xt_x0, xu_prev = xt_u0;
#pragma unroll kChainUnroll
            for (int k = 1; k < N; k++) {
                const double xx = (A_(WX, k) * A_(DIX, k) - A_(L2X, k - 1) * a2) - A_(L1U, k - 1) * a1; // x_k
                const double xu = (A_(WU, k) * A_(DIU, k) - A_(L2U, k - 1) * a1) - A_(L1X, k) * xx;     // u_k
                double zt = A_(CM, k) * xx;
                zt += A_(CA, k) * xx_prev + A_(CB, k) * xu_prev;
                const double dy = rho_e * (alpha * zt + (1.0 - alpha) * 0.0 - 0.0);
                A_(YE, k) += dy;
                const double xp = A_(XX, k), xn = alpha * xx + (1.0 - alpha) * xp;
                A_(XX, k) = xn;
                const double up = A_(XU, k), un = alpha * xu + (1.0 - alpha) * up;
                A_(XU, k) = un;
                if (KEEP) { dyg[(size_t)k * B] = dy; A_(WX, k) = xn - xp; A_(WU, k) = un - up; }
                a2 = xx; a1 = xu; xx_prev = xx; xu_prev = xu;
            }
            {
                const double xx = (A_(WX, N) * A_(DIX, N) - A_(L2X, N - 1) * a2) - A_(L1U, N - 1) * a1;
                double zt = A_(CM, N) * xx;
                zt += A_(CA, N) * xx_prev + A_(CB, N) * xu_prev;
                const double dy = rho_e * (alpha * zt + (1.0 - alpha) * 0.0 - 0.0);
                A_(YE, N) += dy;
                const double xp = A_(XX, N), xn = alpha * xx + (1.0 - alpha) * xp;
                A_(XX, N) = xn;
                if (KEEP) { dyg[(size_t)N * B] = dy; A_(WX, N) = xn - xp; }
            }
            z0prev = lb0;
            iter_done = iter;
        };
        if (act) {
            if (keep_delta) backward(std::true_type{});
            else backward(std::false_type{});
        }
        __syncwarp(WMASK); // hull z / y of this iteration visible to both lanes
        // ---- termination / adaptive rho (pair-uniform decisions, warp-uniform barriers)
        if (check_now) {
            if (adapt_now) update_info(std::true_type{}); else update_info(std::false_type{});
            const int stop = check_termination(0);
            if (act && stop) { status = stop; done = true; }
        }
        if (adapt_now) {
            if (!check_now) update_info(std::true_type{});
            if (act && !done) {
                // compute_rho_estimate + adapt_rho
                const double pr = rp_s / (fmax(nz_s, nAx_s) + 1e-10);
                const double dr = rd_s / (fmax(fmax(nq_s, nAty_s), nPx_s) + 1e-10);
                double rn = rho * sqrt(pr / (dr + 1e-10));
                rn = fmin(fmax(rn, kRhoMin), kRhoMax);
                if (rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance) {
                    rho = rn;
                    rho_updates++;
                    rho_e_inv = 1.0 / (kRhoEqOverIneq * rho); rho_inv = 1.0 / rho;
                }
            }
            __syncwarp(WMASK);
            // warp-uniform; pairs whose rho did not change recompute the same factors
            chain_factor(d, i, a, N, c, rho, nh, pmask, !skip, nullptr, nullptr, IPW);
            __syncwarp(WMASK);
            load_factor();
            cp_async_wait_all();
#pragma unroll
            for (int r = 0; r < 10; r++) CF[r] = d.CF[(size_t)r * B + i];
            __syncwarp(WMASK);
        }
        if (iter >= max_iter) { // warp-uniform: osqp_solve leaves the loop
            if (!check_now) {
                update_info(std::false_type{});
                const int stop = check_termination(0);
                if (act && !done && stop) { status = stop; done = true; }
            }
            {
                const int stop = check_termination(1);
                if (act && !done) {
                    status = stop ? stop : WC_OSQP_MAX_ITER_REACHED;
                    done = true;
                }
            }
        }
        if (!PHASE2 && iter >= phase1_len) break; // unfinished instances go to the continuation list
        warp_busy = __any_sync(WMASK, !done);
    }
    __syncwarp(WMASK);

    // ------------------------------------------------------------------ epilogue: store_solution + controller checks
    const bool unfinished = !skip && !done; // phase 1 only
    const bool has_sol = !(status == WC_OSQP_PRIMAL_INFEASIBLE || status == WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                           status == WC_OSQP_DUAL_INFEASIBLE || status == WC_OSQP_DUAL_INFEASIBLE_INACCURATE ||
                           status == WC_OSQP_UNSOLVED || status == WC_OSQP_NON_CVX) || unfinished;
    if (!skip) {
        double* xg = d.x + (size_t)(a * nA) * B + i;
        double* yg = d.yeq + (size_t)(a * nR) * B + i;
#pragma unroll
        for (int k = 0; k < nR; k++) {
            xg[(size_t)k * B] = has_sol ? A_(XX, k) : 0.0;
            yg[(size_t)k * B] = has_sol ? A_(YE, k) : 0.0;
            if (k < N) xg[(size_t)(nR + k) * B] = has_sol ? A_(XU, k) : 0.0;
        }
        for (int h = a; h < kMaxHull; h += 2) {
            d.yh[(size_t)h * B + i] = has_sol ? H_(H_YH, h) : 0.0;
            d.zh[(size_t)h * B + i] = has_sol ? H_(H_ZH, h) : 0.0;
        }
        d.z0[(size_t)a * B + i] = has_sol ? z0prev : 0.0;
    }
    // desired ZMP of both axes on the x lane
    const double u0_mine = A_(XU, 0) * du0_scale;
    const double u0_other = __shfl_xor_sync(pmask, u0_mine, IPW);
    const double up_other = __shfl_xor_sync(pmask, upa, IPW);
    if (a == 0 && valid) {
        if (unfinished) {
            const int idx = atomicAdd(d.cont_count, 1);
            d.cont_list[idx] = i;
            d.rho[i] = rho;
            d.rho_updates[i] = rho_updates;
            if (init_state == 2) d.inited[i] = 1;
        } else {
            int flags = kflags;
            double out[2] = {upa, up_other};
            if (!skip) {
                d.rho[i] = rho;
                d.rho_updates[i] = rho_updates;
                if (init_state == 2) d.inited[i] = 1;
                if (osqp_eigen_solve_ok(status, st.osqp_eigen_checks_status)) {
                    out[0] = has_sol ? u0_mine : NAN; out[1] = has_sol ? u0_other : NAN;
                    d.uprev[i] = out[0]; d.uprev[B + i] = out[1];
                    // ConvexHullProjectionConstraint::computeMargin (Controller.cpp:513-517)
                    // margin < -tol  <=>  outside the polygon AND farther than tol from its boundary: the distance to
                    // the hull segments is only needed for a ZMP that violates a hull row
                    // A ZMP on an active hull row violates it by OSQP's tolerance (~1e-4 m): the hull is the convex hull of
                    // foot rectangles, every interior angle is >= 90 degrees, so the distance to the polygon is at most
                    // sqrt(2) x the largest normalised row violation; only a violation that could reach the tolerance
                    // takes the exact point-to-segment distances (and their dependent loads of the vertices)
                    bool inside = true;
                    double best = INFINITY, vmax = 0.0;
#pragma unroll
                    for (int h = 0; h < kMaxHull; h++) {
                        const double a0 = H_(H_RX, h), a1 = H_(H_RY, h), bb = H_(H_RB, h);
                        const double viol = (a0 * out[0] + a1 * out[1]) - bb;
                        if (h < nh && !(viol <= 0.0)) { inside = false; vmax = fmax(vmax, viol / sqrt(a0 * a0 + a1 * a1)); }
                    }
                    const bool exact = !inside && !(4.0 * vmax < d.hull_tol);
                    if (exact) {
                        for (int h = 0; h < nh; h++) {
                            const int h1 = (h + 1) % nh;
                            const double p0x = d.hullv[(size_t)(2 * h) * B + i], p0y = d.hullv[(size_t)(2 * h + 1) * B + i];
                            const double p1x = d.hullv[(size_t)(2 * h1) * B + i], p1y = d.hullv[(size_t)(2 * h1 + 1) * B + i];
                            const double dx = p1x - p0x, dy = p1y - p0y;
                            const double len2 = dx * dx + dy * dy;
                            double t = len2 > 0 ? ((out[0] - p0x) * dx + (out[1] - p0y) * dy) / len2 : 0.0;
                            t = t < 0 ? 0 : (t > 1 ? 1 : t);
                            const double ex = out[0] - (p0x + t * dx), ey = out[1] - (p0y + t * dy);
                            best = fmin(best, sqrt(ex * ex + ey * ey));
                        }
                    } else if (!inside) {
                        best = 0.0; // outside by less than the tolerance: not flagged
                    }
                    const double margin = inside ? best : -best;
                    if (margin < -d.hull_tol) flags |= WC_MPC_FLAG_OUTSIDE_HULL;
                } else {
                    flags |= WC_MPC_FLAG_NOT_SOLVED;
                }
            }
            if (io.zmp) { io.zmp[i] = out[0]; io.zmp[B + i] = out[1]; }
            if (io.status) io.status[i] = skip ? WC_OSQP_UNSOLVED : status;
            if (io.iters) io.iters[i] = skip ? 0 : iter_done;
            if (io.flags) io.flags[i] = flags;
            d.tick_flags[i] = flags;
        }
    }
    __syncwarp(WMASK); // the next group of this CTA reuses the shared-memory working set
    } // groups
#undef A_
#undef H_
}

// ------------------------------------------------------------------------------------------------
// instances per CTA for horizon N: as many as fit in one CTA's shared memory (16 at N <= 53, ... 1 at N <= ~880)
int mpc_chain_ipw(int N)
{
    const Off O(N);
    const size_t maxs = (size_t)mpc_admm_max_smem();
    for (int ipw = kIPWmax; ipw >= 1; ipw >>= 1)
        if (((size_t)O.total * 2 * ipw + (size_t)H_COUNT * 8 * ipw + 2) * sizeof(double) <= maxs) return ipw;
    return 0;
}

size_t mpc_chain_smem_bytes(int N)
{
    const Off O(N);
    const int ipw = mpc_chain_ipw(N);
    return ((size_t)O.total * 2 * ipw + (size_t)H_COUNT * 8 * ipw + 2) * sizeof(double); // + the staging mbarrier
}

static size_t mpc_chain_setup_warp_doubles(int N) { return (size_t)(3 * (4 * N + 2) + 2 * (2 * N + 2) + 32); }

// the chain kernels run any horizon whose per-chain working set fits one CTA (1 instance per CTA at N ~ 880)
bool mpc_chain_supported(int N)
{
    return N >= 2 && mpc_chain_ipw(N) >= 1 &&
           mpc_chain_setup_warp_doubles(N) * kSetupWarps * sizeof(double) <= (size_t)mpc_admm_max_smem();
}

template <int NT, int IPW>
static cudaError_t launch_chain_t(const MpcDev& d, const MpcTickIO& io, cudaStream_t s, cudaEvent_t* mid, size_t smem)
{
    auto k1 = mpc_chain_kernel<false, NT, IPW>;
    auto k2 = mpc_chain_kernel<true, NT, IPW>;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int blocks = (d.B + IPW - 1) / IPW;
    if (mid) cudaEventRecord(mid[0], s); // per-kernel timing (wc_mpc_enable_profiling)
    k1<<<blocks, 2 * IPW, smem, s>>>(d, io);
    if (mid) cudaEventRecord(mid[1], s);
    k2<<<blocks < 2 * 148 ? blocks : 2 * 148, 2 * IPW, smem, s>>>(d, io); // grid-stride over the continuation list
    return cudaGetLastError();
}

// osqp_setup of the instances K1 listed (Ruiz equilibration; with `factor` also the chain kernels' LDL' to HBM)
cudaError_t launch_mpc_chain_setup(const MpcDev& d, cudaStream_t s, bool factor)
{
    const size_t td = mpc_chain_setup_warp_doubles(d.N);
    // next to the split-chain phase 1 (d.setup_overlap) single-warp CTAs: 3.5 KB of shared memory each at N = 25, which fits beside
    // the five 42 KB CTAs of that kernel on an SM instead of taking the place of one
    const int warps = d.setup_overlap ? 1 : kSetupWarps;
    const size_t smem_setup = td * warps * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(mpc_chain_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(td * kSetupWarps * sizeof(double)));
    if (e != cudaSuccess) return e;
    const int need = (d.B + warps - 1) / warps;
    const int cap = 4 * 148 * (kSetupWarps / warps);
    mpc_chain_setup_kernel<<<need < cap ? need : cap, warps * 32, smem_setup, s>>>(d, (int)td, factor ? 1 : 0); // grid-stride over setup_list
    return cudaGetLastError();
}

cudaError_t launch_mpc_chain(const MpcDev& d, const MpcTickIO& io, cudaStream_t s, cudaEvent_t* mid)
{
    const int N = d.N;
    const size_t smem = mpc_chain_smem_bytes(N);
    cudaError_t e = launch_mpc_chain_setup(d, s, true);
    if (e != cudaSuccess) return e;
    switch (mpc_chain_ipw(N)) {
    case 16: return N == 25 ? launch_chain_t<25, 16>(d, io, s, mid, smem) : launch_chain_t<0, 16>(d, io, s, mid, smem);
    case 8: return launch_chain_t<0, 8>(d, io, s, mid, smem);
    case 4: return launch_chain_t<0, 4>(d, io, s, mid, smem);
    case 2: return launch_chain_t<0, 2>(d, io, s, mid, smem);
    case 1: return launch_chain_t<0, 1>(d, io, s, mid, smem);
    default: return cudaErrorInvalidConfiguration;
    }
}

// default choice of wc_mpc_create: measured on B200 (tools/sweep_horizon.sh) the chain kernels win while at least 8
// instances share a CTA (N <= 106: 1.2x at N = 50, 1.1x at N = 100); beyond that both families are bound by the
// sequential sweeps of a few resident instances per SM and tie (N = 200), and the cooperative kernel is kept.
bool mpc_chain_preferred(int N) { return mpc_chain_supported(N) && mpc_chain_ipw(N) >= 8; }

} // namespace wcb
