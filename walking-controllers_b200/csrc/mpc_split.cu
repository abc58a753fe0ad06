// mpc_split.cu — "split-chain" variant of the batched OSQP-equivalent ADMM (K2): FOUR LANES PER INSTANCE, the
// iterate in REGISTERS, the problem in UNSCALED variables.  The benchmark path (N = 25).
//
// Same QP, same iterates as mpc_chain.cu / mpc_admm_kernel (mapping to the reference in mpc_kernels.cuh:
// WalkingController::solve -> MPCSolver::solve -> OSQP, WM/src/WalkingDCMModelPredictiveController.cpp:491-521,
// WM/src/MPCSolver.cpp:295-322).  What changes is how the reduced KKT step is organised on the machine
// (lane-accurate CPU prototype, checked against dense OSQP iterations: tools/split_lane_proto.py).
//
// 1. Unscaled variables.  OSQP iterates on Pb = c D P D, qb = c D q, Ab = E A D.  With X = D xb, Yh = E yb the step
//        (Pb + sigma I + Ab' rho Ab) xt = sigma xb - qb + Ab' (rho zb - yb)
//    is, after a left multiplication by D^-1,
//        (c P + sigma D^-2 + A' (rho E^2) A) Xt = sigma D^-2 X - c q + A' E (rho zb - yb)
//    in which A carries the CONSTANT coefficients (-1, a, b) of the LIPM recursion: the three coefficient arrays of
//    the scaled rows disappear, and the unscaled residuals of check_termination need no scaling vectors at all.
//    The iterates are OSQP's (same scaled problem in exact arithmetic; the persistent workspace in HBM stays OSQP's
//    scaled x, y, z): iteration counts and status match the oracle on every test.
// 2. Split chain.  Per axis the unknowns x_N, u_{N-1}, ..., x_1, u_0, x_0 form a pentadiagonal chain.  Two lanes share
//    it: the TOP lane eliminates x_N, u_{N-1}, ... downwards, the BOTTOM lane x_0, x_1, u_1, ... upwards with u_0
//    carried as a border column; they meet at a two-unknown separator (x_s, u_{s-1}).  What is left is a dense 6x6 on
//    (x_s, u_{s-1}, u_0) of both axes (the hull rows couple u_0x and u_0y), whose inverse is kept in shared memory.
//    A warp = 8 instances x (2 axes x 2 halves); the dependent chain per lane is half as long as in mpc_chain.cu.
// 3. Registers.  A lane's T = ceil((N+1)/2) slots are fully unrolled (compile-time horizon), so its share of the
//    iterate (X, U, Yh) and the forward-sweep result live in registers (5 T doubles); shared memory holds constants
//    only (10 arrays + the border): ~16 LDS per slot and iteration instead of ~30 LDS/STS per stage, which is what
//    bounded the chain kernel (shared-memory pipe).  ~42 KB per single-warp CTA -> 5 CTAs per SM.
// 4. The LDL' factors are rebuilt in the kernel every tick from (D, E, c, rho) -- one ADMM iteration's worth of
//    work out of 25 -- instead of being kept in HBM: no factor arrays, no extra traffic, and a rho update refactors
//    in place.
// 5. Everything outside the iteration loop is written for a small instruction footprint.  With five single-warp CTAs
//    per SM nothing hides a warp's own latencies, and the once-per-warp code (stage-in, update_info, check_termination,
//    epilogue), fully unrolled, was half of the kernel's stall samples -- most of them instruction fetches from L2.
//    Now: raw values travel HBM -> shared memory by cp.async and are turned into constants by a real loop; once the
//    iterations of a pass are over the iterate is copied to shared memory and update_info / the epilogue walk the
//    slots with real loops (phase 1 re-uses the dead factor arrays for that, phase 2 has a dump area of its own);
//    infinity norms are compare + select (OSQP's vec_norm_inf) instead of fmax, a seven-instruction sequence on fp64.
//
// Slots.  top lane   t = 0..T-1   <-> (x_k, u_{k-1}), equality row k,   k = N - t;  slot T-1 is the separator
//         bottom lane t = OFF..T-1 <-> (x_j, u_j),     equality row j+1, j = t - OFF; u_0 (slot OFF) is the border,
//         its last slot holds x_{s-1} and a shadow copy of the separator's u_{s-1} and row s (kept bitwise equal).
#include "mpc_kernels.cuh"

#include <cmath>
#include <cstdio>
#include <type_traits>

namespace wcb {

namespace {

constexpr unsigned FULL = 0xffffffffu;
#ifndef WC_SPLIT_MINB
#define WC_SPLIT_MINB 5 // warps resident per SM (shared memory: 42 KB each at N = 25)
#endif
#ifndef WC_SPLIT_WARPS
#define WC_SPLIT_WARPS 1 // warps per CTA; they share nothing but the instruction stream (each has its own working set)
#endif

template <int N>
struct SplitDims {
    static constexpr int T = (N + 2) / 2;   // slots per lane
    static constexpr int TB = N + 1 - T;    // real slots of the bottom lane
    static constexpr int OFF = T - TB;      // leading dummy slots of the bottom lane (0 or 1)
    enum { A_SX = 0, A_SU, A_CQ, A_EE, A_L1X, A_L2X, A_DIX, A_L1U, A_L2U, A_DIU, A_COUNT };
    static constexpr int oLBX = A_COUNT * T * 32; // border multipliers, bottom lanes only: [T][16]
    static constexpr int oLBU = oLBX + T * 16;
    static constexpr int kMAxis = 19;              // 3 rows x 6 + 1 pad: the two axes land on complementary banks
    static constexpr int oM = oLBU + T * 16;       // rows (x_s, u_{s-1}, u_0) of the 6x6 inverse per axis: [2][19][8]
    enum { H_HXS = 0, H_HYS, H_LO, H_UP, H_EH, H_EHI, H_COUNT };
    static constexpr int oH = oM + 2 * kMAxis * 8; // hull constants [H_COUNT][8 rows][8 instances]
    static constexpr int total = oH + H_COUNT * 8 * 8;
    // copy of the iterate (X, U, Y) and of 1 / D, 1 / E for the rolled update_info / epilogue: phase 1 overlays the six factor arrays
    // (dead once its iterations are over), phase 2 goes back into the loop and gets six arrays of its own (its occupancy is irrelevant)
    static constexpr int oDump1 = A_L1X * T * 32;
    static constexpr int oDump2 = total;
    static constexpr int total2 = total + 6 * T * 32;
};

__device__ __forceinline__ double sx16(double v) { return __shfl_xor_sync(FULL, v, 16); }
__device__ __forceinline__ double sx8(double v) { return __shfl_xor_sync(FULL, v, 8); }
// reductions over the four lanes of an instance (lane bits 3 and 4); symmetric, so all four get the same bits
__device__ __forceinline__ double inst_max(double v) { v = fmax(v, sx8(v)); return fmax(v, sx16(v)); }
// the same over the four lanes of an instance, for accumulators built by amax (never NaN, never negative: symmetric)
__device__ __forceinline__ double inst_amax(double v)
{
    const double p = sx8(v);
    v = p > v ? p : v;
    const double q = sx16(v);
    return q > v ? q : v;
}
__device__ __forceinline__ double inst_sum(double v) { v += sx8(v); return v + sx16(v); }

// LDL' of a symmetric 6x6 (lower triangle of M), then rows of the inverse by two triangular solves
struct Ldl6 {
    double L[6][6], Dinv[6];
    __device__ __forceinline__ void factor(const double (*M)[6])
    {
        double Dv[6];
#pragma unroll
        for (int j = 0; j < 6; j++) {
            double dj = M[j][j];
            double lj[6];
#pragma unroll
            for (int k = 0; k < j; k++) { lj[k] = L[j][k] * Dv[k]; dj -= L[j][k] * lj[k]; }
            Dv[j] = dj;
            Dinv[j] = 1.0 / dj;
#pragma unroll
            for (int i = j + 1; i < 6; i++) {
                double v = M[i][j];
#pragma unroll
                for (int k = 0; k < j; k++) v -= L[i][k] * lj[k];
                L[i][j] = v * Dinv[j];
            }
        }
    }
    __device__ __forceinline__ void solve(double* r) const
    {
#pragma unroll
        for (int i = 1; i < 6; i++) {
#pragma unroll
            for (int k = 0; k < i; k++) r[i] -= L[i][k] * r[k];
        }
#pragma unroll
        for (int i = 0; i < 6; i++) r[i] *= Dinv[i];
#pragma unroll
        for (int i = 4; i >= 0; i--) {
#pragma unroll
            for (int k = i + 1; k < 6; k++) r[i] -= L[k][i] * r[k];
        }
    }
};

} // namespace

// =====================================================================================
// ADMM.  PHASE2 == false: all instances, iterations 1 .. first termination check.
//        PHASE2 == true : the instances of the continuation list, until osqp_solve returns.
// One warp per CTA, 8 instances per warp: lane = half * 16 + axis * 8 + instance slot.
// =====================================================================================
// MODE 0: phase 1 of all instances (with d.setup_overlap: except those on this tick's setup_list), 1: phase 2,
//      2: phase 1 of the instances on the setup_list (side stream, after their osqp_setup)
template <int MODE, int NT>
__global__ void __launch_bounds__(32 * WC_SPLIT_WARPS, WC_SPLIT_MINB / WC_SPLIT_WARPS) mpc_split_kernel(MpcDev d, MpcTickIO io)
{
    constexpr bool PHASE2 = MODE == 1, LISTED = MODE == 2;
    // MODE 2 runs NEXT TO the MODE 0 launch, whose CTAs have read the workspace rows of listed instances (and of their neighbours
    // in the same 32-byte sector) before the set-up kernel rewrote them: an SM's L1 can still hold those lines, and no kernel
    // boundary separates the two launches on that SM.  Everything the set-up kernel writes (D, 1/D, E, 1/E, c, rho, inited) is
    // therefore read through L2 here (ld.global.cg; cp.async.cg only moves 16 bytes, so the stage-in is synchronous in this mode).
    auto ldc = [](const double* p) -> double { return LISTED ? __ldcg(p) : __ldg(p); };
    using SD = SplitDims<NT>;
    constexpr int N = NT, T = SD::T, OFF = SD::OFF, nA = 2 * N + 1, nR = N + 1;
    static_assert(N >= 4, "the split needs a separator away from u_0 / u_1");
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31;
    double* const sm = sm_all + (size_t)(threadIdx.x >> 5) * (PHASE2 ? SD::total2 : SD::total);
    const int s8 = lane & 7, a = (lane >> 3) & 1, q4 = lane >> 3;
    const bool bot = lane >= 16;
    const size_t B = (size_t)d.B;
    // (not const: phase 1 adds a run-time zero to them in every iteration, see there)
    double* sl = sm + lane;                          // full arrays: element (arr, t) at sl[(arr * T + t) * 32]
    double* sb = sm + SD::oLBX + (lane & 15);        // border arrays (bottom lanes)
    double* smr = sm + SD::oM + a * SD::kMAxis * 8 + s8;
    double* sh = sm + SD::oH + s8;                   // hull constants: (arr, h) at sh[(arr * 8 + h) * 8]
#define S_(arr, t) sl[((SD::arr) * T + (t)) * 32]
#define SV_(arr, t) (*(volatile const double*)&sl[((SD::arr) * T + (t)) * 32])
#define LBXV_(t) (*(volatile const double*)&sb[(t) * 16])
#define LBUV_(t) (*(volatile const double*)&sb[(T + (t)) * 16])
#define LBX_(t) sb[(t) * 16]
#define LBU_(t) sb[(T + (t)) * 16]
#define HC_(arr, h) sh[((SD::arr) * 8 + (h)) * 8]
    // setup_count is consumed by the setup kernel of this tick (and by the MODE 2 launch): cleared by the last kernel of the tick
    if ((d.setup_overlap ? PHASE2 : MODE == 0) && blockIdx.x == 0 && threadIdx.x == 0) *d.setup_count = 0;
    const int count = PHASE2 ? *d.cont_count : (LISTED ? *d.setup_count : d.B);
    const int ngroups = (count + 7) / 8;
    const wc_osqp_settings& st = d.st;
    const double sigma = st.sigma, alpha = st.alpha, infty = st.infty;
    const double a_ = d.a_dyn, b_ = d.b_dyn;
    // lane constants of the generic slot program: A'(-Yh) on x = cy0 Yh[t] + cy1 Yh[t-1]; row = b u + (-cy0) x + (-cy1) x_next
    const double cy0 = bot ? -a_ : 1.0, cy1 = bot ? 1.0 : -a_;
    const double cx0 = -cy0, cx1 = -cy1;
    const double k1 = 1.0 / alpha - 1.0;             // X~ = X+ + k1 dX

    for (int grp = blockIdx.x * WC_SPLIT_WARPS + (threadIdx.x >> 5); grp < ngroups; grp += gridDim.x * WC_SPLIT_WARPS) {
    const int li = grp * 8 + s8;
    const bool valid = li < count;
    const int i = PHASE2 ? d.cont_list[valid ? li : count - 1] : (LISTED ? d.setup_list[valid ? li : count - 1] : (valid ? li : d.B - 1));
    // ------------------------------------------------------------------ stage in: raw HBM -> shared memory -> constants / registers
    // Everything a group needs from HBM is REQUESTED before anything is waited for (five single-warp CTAs per SM: an exposed
    // round trip is paid in full): (1a) cp.async of eight raw values per slot straight into shared memory (no register is
    // involved, so all requests of a lane are in flight together); the scalars of the instance; (1b) the gradient ring, whose
    // rows depend on the ring head; the hull rows and equality row 0; then ONE wait; (2) in place, by a real loop, the constants
    // of the iteration (sigma D^-2, c q, E^2) and the unscaled warm start; (3) the warm start into registers.
    //   raw value  ->  array            raw value  ->  array
    //   1/D (x)        A_SX             D (x)          A_L1X   (then X = D x)
    //   1/D (u)        A_SU             x (x part)     A_L2X
    //   gradient       A_CQ             D (u)          A_DIX   (then U = D u)
    //   E              A_EE             x (u part)     A_L1U
    //                                   y              A_L2U   (then Y = E y)
    // A new OSQP workspace (cold: x = y = z = 0) is applied in pass 2, so that no request depends on a loaded flag.
#define WC_CPA8(dst, src)                                                                                              \
    do {                                                                                                               \
        if (LISTED) {                                                                                                  \
            const double v_ = __ldcg(src);                                                                             \
            asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(dst), "d"(v_) : "memory");                                  \
        } else {                                                                                                       \
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");                \
        }                                                                                                              \
    } while (0)
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(sl);
    constexpr unsigned kArr = (unsigned)T * 256u;    // bytes between arrays
    {
        const int kx0 = bot ? -OFF : N;              // x index of slot 0; +1 / -1 per slot
        const ptrdiff_t sB = bot ? (ptrdiff_t)B : -(ptrdiff_t)B;
        const ptrdiff_t xrow0 = ((ptrdiff_t)(a * nA + kx0)) * (ptrdiff_t)B + i;
        const ptrdiff_t urow0 = ((ptrdiff_t)(a * nA + nR + kx0 - (bot ? 0 : 1))) * (ptrdiff_t)B + i;
        const ptrdiff_t erow0 = ((ptrdiff_t)(a * nR + kx0 + (bot ? 1 : 0))) * (ptrdiff_t)B + i;
        const double* gDx = d.D + xrow0;
        const double* gDix = d.Dinv + xrow0;
        const double* gDu = d.D + urow0;
        const double* gDiu = d.Dinv + urow0;
        const double* gE = d.Eeq + erow0;
        const double* gX = d.x + xrow0;
        const double* gU = d.x + urow0;
        const double* gY = d.yeq + erow0;
#pragma unroll(LISTED ? T : 1)
        for (int t = 0; t < T; t++) {
            const unsigned st = s0 + (unsigned)t * 256u;
            const ptrdiff_t o = t * sB;
            if ((t >= OFF) || !bot) {
                WC_CPA8(st + SD::A_SX * kArr, gDix + o);
                WC_CPA8(st + SD::A_SU * kArr, gDiu + o);
                WC_CPA8(st + SD::A_EE * kArr, gE + o);
                WC_CPA8(st + SD::A_L1X * kArr, gDx + o);
                WC_CPA8(st + SD::A_DIX * kArr, gDu + o);
                WC_CPA8(st + SD::A_L2X * kArr, gX + o);
                WC_CPA8(st + SD::A_L1U * kArr, gU + o);
                WC_CPA8(st + SD::A_L2U * kArr, gY + o);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    const int kflags_l = d.tick_flags[i];
    const bool pending = MODE == 0 && (kflags_l & kMpcFlagSetupPending) != 0; // not this launch's instance
    const int kflags = kflags_l & ~kMpcFlagSetupPending;
    const int nh_l = d.nhull[i];
    const double Qa = d.Qd[(size_t)a * B + i], Ra = d.Rd[(size_t)a * B + i];
    const int init_state = LISTED ? __ldcg(d.inited + i) : d.inited[i];
    const double c = LISTED ? __ldcg(d.cscale + i) : d.cscale[i];
    double rho = LISTED ? __ldcg(d.rho + i) : d.rho[i];
    const double upa = d.uprev[(size_t)a * B + i];
    const int head = mpc_ring_head(d);
    const double dcm_a = io.dcm[(size_t)a * B + i];
    // equality row 0 (bottom lanes) and the hull rows q4, q4 + 4: raw values
    const double E0_l = ldc(d.Eeq + ((size_t)(a * nR) * B + i));
    const double Y0_l = d.yeq[(size_t)(a * nR) * B + i];
    const double z0_l = PHASE2 ? 0.0 : d.z0[(size_t)a * B + i];
    const double Ei0_l = PHASE2 ? 0.0 : ldc(d.Einv + ((size_t)(a * nR) * B + i));
    double he_l[2], hx_l[2], hy_l[2], hb_l[2], zh_l[2], yh_l[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int h = q4 + 4 * r;
        he_l[r] = ldc(d.Eh + ((size_t)h * B + i));
        hx_l[r] = __ldg(d.hullA + ((size_t)(2 * h) * B + i)); hy_l[r] = __ldg(d.hullA + ((size_t)(2 * h + 1) * B + i));
        hb_l[r] = __ldg(d.hullb + ((size_t)h * B + i));
        zh_l[r] = d.zh[(size_t)h * B + i]; yh_l[r] = d.yh[(size_t)h * B + i];
    }
    { // (1b) the gradient ring
        int gs = head + (bot ? -OFF : N); gs = gs >= nR ? gs - nR : (gs < 0 ? gs + nR : gs); // ring slot of slot 0's stage
        const int gstep = bot ? 1 : -1;
        const double* gG = d.grad + ((size_t)(a * nR) * B + i);
#pragma unroll 1
        for (int t = 0; t < T; t++) {
            if ((t >= OFF) || !bot) WC_CPA8(s0 + (unsigned)t * 256u + SD::A_CQ * kArr, gG + (size_t)gs * B);
            gs += gstep; gs = gs >= nR ? gs - nR : (gs < 0 ? gs + nR : gs);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
#undef WC_CPA8
    const bool skip = !valid || pending || (kflags & (WC_MPC_FLAG_NO_CONTACT | WC_MPC_FLAG_NO_SOLVER)) != 0;
    const int nh = skip ? 0 : nh_l;
    const bool cold = !PHASE2 && !skip && init_state == 2; // new OSQP workspace: x = y = z = 0
    const double cQ = c * Qa, cR = c * Ra;
    const double cinv = 1.0 / c;
    const double cqu0 = c * (-(Ra * upa));           // c q of u_0 (MPCSolver.cpp:244-245: G u_prev, first block only)
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    double X[T], U[T], Y[T], tX[T], tU[T];
    {
#pragma unroll 1
        for (int t = 0; t < T; t++) {
            const bool real = (t >= OFF) || !bot;
            double dx = 0, dix = 0, du = 0, diu = 0, e = 0, g = 0, xb = 0, ub = 0, yb = 0;
            if (real) {
                dix = S_(A_SX, t); diu = S_(A_SU, t); g = S_(A_CQ, t); e = S_(A_EE, t); dx = S_(A_L1X, t); du = S_(A_DIX, t);
                if (!cold) { xb = S_(A_L2X, t); ub = S_(A_L1U, t); yb = S_(A_L2U, t); }
            }
            S_(A_SX, t) = sigma * (dix * dix);
            S_(A_SU, t) = sigma * (diu * diu);
            S_(A_CQ, t) = c * g;
            S_(A_EE, t) = e * e;
            S_(A_L1X, t) = dx * xb; S_(A_DIX, t) = du * ub; S_(A_L2U, t) = e * yb;
        }
#pragma unroll
        for (int t = 0; t < T; t++) {
            X[t] = S_(A_L1X, t); U[t] = S_(A_DIX, t); Y[t] = S_(A_L2U, t);
            tX[t] = 0.0; tU[t] = 0.0;
        }
    }
    // bottom lanes: equality row 0 (l = u = -dcm, MPCSolver.cpp:143-146) and its unscaled z
    double EE0 = 0.0, Y0 = 0.0, Z0 = 0.0, E0 = 1.0;
    if (bot) {
        E0 = E0_l;
        EE0 = E0 * E0;
        if (!cold) Y0 = E0 * Y0_l;
        Z0 = PHASE2 ? -dcm_a : (cold ? 0.0 : z0_l * Ei0_l);
    }
    // hull rows q4 and q4 + 4 of the instance: constants to shared memory, (z, y) in registers (scaled, as OSQP keeps them)
    double zh[2], yh[2], dyh[2] = {0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int h = q4 + 4 * r;
        const bool on = h < nh;
        const double e = he_l[r], hx = hx_l[r], hy = hy_l[r], hb = hb_l[r];
        HC_(H_HXS, h) = on ? e * hx : 0.0;
        HC_(H_HYS, h) = on ? e * hy : 0.0;
        HC_(H_LO, h) = on ? e * (-infty) : -1.0;
        HC_(H_UP, h) = on ? e * hb : 1.0;
        HC_(H_EH, h) = on ? e : 1.0;
        HC_(H_EHI, h) = on ? 1.0 / e : 1.0;
        zh[r] = (on && !cold) ? zh_l[r] : 0.0;
        yh[r] = (on && !cold) ? yh_l[r] : 0.0;
    }
    __syncwarp();

    // ------------------------------------------------------------------ numeric factorisation (per tick and after a rho update)
    auto build_factor = [&]() {
        const double rho_e = kRhoEqOverIneq * rho;
        const double a2 = a_ * a_, b2 = b_ * b_, ab = a_ * b_;
        double hxx = 0.0, hyy = 0.0, hxy = 0.0;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const double hx = HC_(H_HXS, q4 + 4 * r), hy = HC_(H_HYS, q4 + 4 * r);
            hxx += hx * hx; hyy += hy * hy; hxy += hx * hy;
        }
        hxx = inst_sum(hxx); hyy = inst_sum(hyy); hxy = inst_sum(hxy);
        const double hd = rho * (a == 0 ? hxx : hyy), kxy = rho * hxy;
        // carried: multipliers and (multiplier x pivot) of the previous two unknowns of the elimination order
        double l1p = 0, m1p = 0, l2p = 0, m2p = 0, lbp = 0, l2pp = 0, m2pp = 0, lbpp = 0, Sbb = 0;
        double Wprev = bot ? rho_e * EE0 : 0.0;
        double pxx = 0, puu = 0, pxu = 0, p0x = 0, p0u = 0, p00 = 0, g_u0 = 0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            const double W = rho_e * S_(A_EE, t);
            const double sc = (t == 0) ? (bot ? 2.0 : 1.0) : 2.0; // u_{N-1} (top slot 0) has the single diagonal block of Theta' R Theta
            double g_x = cQ + S_(A_SX, t) + (bot ? (Wprev + a2 * W) : (W + a2 * Wprev));
            double g_u = cR * sc + S_(A_SU, t) + b2 * W;
            double s1x = bot ? ab * W : -b_ * W;
            double s2x = -a_ * W;
            double s1u = bot ? -b_ * W : ab * W;
            double s2u = -cR;
            double bdx = 0.0, bdu = 0.0;
            bool u_dummy = false;
            if (t == OFF) { if (bot) { s1x = 0.0; bdx = ab * W; u_dummy = true; g_u0 = g_u + hd; } }
            if (t == OFF + 1) { if (bot) { bdx = -b_ * (rho_e * S_(A_EE, OFF)); bdu = -cR; } }
            if (t < OFF) { if (bot) { g_x = 1.0; g_u = 1.0; s1x = s2x = s1u = s2u = 0.0; u_dummy = true; } }
            if (t == T - 1) { // the separator seen from the top lane: Schur terms of its last two unknowns
                if (!bot) {
                    pxx = g_x - (l1p * m1p + l2pp * m2pp);
                    puu = g_u - l2p * m2p;
                    pxu = s1x - l2p * m1p;
                }
            }
            { // x part
                const double dd = g_x - l1p * m1p - l2pp * m2pp;
                const double m1 = s1x - l2p * m1p, m2 = s2x, mb = bdx - lbp * m1p - lbpp * m2pp;
                const double di = 1.0 / dd;
                const double l1 = m1 * di, l2 = m2 * di, lb = mb * di;
                S_(A_L1X, t) = l1; S_(A_L2X, t) = l2;
                S_(A_DIX, t) = (t < OFF) ? (bot ? 0.0 : di) : di;
                if (bot) LBX_(t) = lb;
                Sbb += lb * mb;
                l2pp = l2p; m2pp = m2p; lbpp = lbp;
                l1p = l1; m1p = m1; l2p = l2; m2p = m2; lbp = lb;
            }
            if (t == T - 1) { // bottom: near = u_{s-1}, far = x_s
                if (bot) {
                    pxx = -(l2p * m2p);
                    puu = -(l1p * m1p + l2pp * m2pp);
                    pxu = -(l2p * m1p);
                    p0x = -(lbp * m2p);
                    p0u = ((T - 1 - OFF) == 1 ? -cR : 0.0) - (lbp * m1p + lbpp * m2pp);
                    p00 = g_u0 - Sbb;
                }
                break;
            }
            { // u part
                const double dd = g_u - l1p * m1p - l2pp * m2pp;
                double m1 = s1u - l2p * m1p, m2 = s2u, mb = bdu - lbp * m1p - lbpp * m2pp;
                double di = 1.0 / dd;
                double l1 = m1 * di, l2 = m2 * di, lb = mb * di;
                if (u_dummy) { m1 = m2 = mb = 0.0; l1 = l2 = lb = 0.0; di = 0.0; }
                S_(A_L1U, t) = l1; S_(A_L2U, t) = l2; S_(A_DIU, t) = di;
                if (bot) LBU_(t) = lb;
                Sbb += lb * mb;
                l2pp = l2p; m2pp = m2p; lbpp = lbp;
                l1p = l1; m1p = m1; l2p = l2; m2p = m2; lbp = lb;
            }
            if (t >= OFF || !bot) Wprev = W;
        }
        S_(A_L1U, T - 1) = 0.0; S_(A_L2U, T - 1) = 0.0; S_(A_DIU, T - 1) = 0.0;
        if (bot) LBU_(T - 1) = 0.0;
        // the 6x6 on (x_s, u_{s-1}, u_0) of this axis and of the other one
        pxx += sx16(pxx); puu += sx16(puu); pxu += sx16(pxu); p0x += sx16(p0x); p0u += sx16(p0u); p00 += sx16(p00);
        const double oxx = sx8(pxx), ouu = sx8(puu), oxu = sx8(pxu), o0x = sx8(p0x), o0u = sx8(p0u), o00 = sx8(p00);
        double M[6][6];
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int cc = 0; cc < 6; cc++) M[r][cc] = 0.0;
        M[0][0] = pxx; M[1][0] = pxu; M[1][1] = puu; M[2][0] = p0x; M[2][1] = p0u; M[2][2] = p00;
        M[3][3] = oxx; M[4][3] = oxu; M[4][4] = ouu; M[5][3] = o0x; M[5][4] = o0u; M[5][5] = o00;
        M[5][2] = kxy;
        Ldl6 F;
        F.factor(M);
        // rows of the inverse: the top lane solves for x_s and u_{s-1}, the bottom lane for u_0
        double r0[6], r1[6];
#pragma unroll
        for (int k = 0; k < 6; k++) { r0[k] = (k == (bot ? 2 : 0)) ? 1.0 : 0.0; r1[k] = (k == 1) ? 1.0 : 0.0; }
        F.solve(r0);
        F.solve(r1);
#pragma unroll
        for (int k = 0; k < 6; k++) {
            smr[((bot ? 2 : 0) * 6 + k) * 8] = r0[k];
            if (!bot) smr[(6 + k) * 8] = r1[k];
        }
        __syncwarp();
    };

    // ------------------------------------------------------------------ ADMM (osqp_solve)
    int status = WC_OSQP_UNSOLVED;
    int iter_done = 0, rho_updates = PHASE2 ? d.rho_updates[i] : 0;
    bool done = skip;
    double pri_res = 0.0, dua_res = 0.0;
    double nz_u = 0, nAx_u = 0, nq_u = 0, nAty_u = 0, nPx_u = 0;
    double rp_s = 0, rd_s = 0, nz_s = 0, nAx_s = 0, nq_s = 0, nAty_s = 0, nPx_s = 0;
    double Yext = 0.0;                               // Yh_0 - W_0 Z_0: the row-0 term of x_0's right-hand side
    double dY0 = 0.0;
    // hull rows' share of the right-hand side of u_0 (this axis, the other axis), from the current (z, y)
    double hp_m = 0.0, hp_o = 0.0;
    auto hull_partials = [&]() {
        double px = 0.0, py = 0.0;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const double v = rho * zh[r] - yh[r];
            px += HC_(H_HXS, q4 + 4 * r) * v;
            py += HC_(H_HYS, q4 + 4 * r) * v;
        }
        hp_m = a == 0 ? px : py;
        hp_o = a == 0 ? py : px;
    };

    // one ADMM iteration on the whole warp; `act` freezes finished instances (phase 2 only)
    auto iteration = [&](const bool act, const double rho_inv) {
        const double rho_e = kRhoEqOverIneq * rho, arho = alpha * rho_e;
        const double oma = 1.0 - alpha;
        // ---- forward: right-hand side fused with L t = rhs, towards the separator
        double txp = 0.0, tup = 0.0;                 // t of the previous slot's x and u
        double n1u = 0.0, n2x = 0.0, n2u = 0.0;      // l1u, l2x, l2u of the previous slot
        double rb = 0.0, vx = 0.0, vu = 0.0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            double yprev;
            if (t == OFF) yprev = bot ? Yext : (t == 0 ? 0.0 : Y[t > 0 ? t - 1 : 0]);
            else yprev = (t == 0) ? 0.0 : Y[t > 0 ? t - 1 : 0];
            double rx = fma(S_(A_SX, t), X[t], -S_(A_CQ, t));
            rx = fma(cy0, Y[t], rx);
            rx = fma(cy1, yprev, rx);
            double ru = S_(A_SU, t) * U[t];
            ru = fma(-b_, Y[t], ru);
            if (t == OFF) { if (bot) rb = ru - cqu0; }
            const double l1x = S_(A_L1X, t), l2x = S_(A_L2X, t);
            const double tx = fma(-n1u, tup, fma(-n2x, txp, rx));
            if (t < T - 1) {
                const double l1u = S_(A_L1U, t), l2u = S_(A_L2U, t);
                const double tu = fma(-l1x, tx, fma(-n2u, tup, ru));
                tX[t] = tx; tU[t] = tu;
                if (bot) { rb = fma(-LBX_(t), tx, rb); rb = fma(-LBU_(t), tu, rb); }
                txp = tx; tup = tu; n1u = l1u; n2x = l2x; n2u = l2u;
            } else {
                tX[t] = tx;
                if (bot) rb = fma(-LBX_(t), tx, rb);
                // contributions to the separator's right-hand side (x_s, u_{s-1})
                vx = bot ? -(l2x * tx) : tx;
                vu = bot ? fma(-l1x, tx, -(n2u * tup)) : fma(-n2u, tup, ru);
            }
        }
        // ---- the dense 6x6: gather the right-hand side (own axis first), multiply by the stored rows of the inverse.
        // The 18 entries of the rows and the constants of the two hull rows are asked for HERE, as volatile loads ahead of the
        // shuffles: placed by the compiler they sit one by one in front of the fma that uses them (register pressure), and the
        // 26 exposed shared-memory latencies were a fifth of the loop's samples.
        double mi[18], hcx[2], hcy[2], hlo[2], hup[2];
#pragma unroll
        for (int k = 0; k < 18; k++) mi[k] = *(volatile const double*)&smr[k * 8];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            hcx[r] = *(volatile const double*)&HC_(H_HXS, q4 + 4 * r); hcy[r] = *(volatile const double*)&HC_(H_HYS, q4 + 4 * r);
            hlo[r] = *(volatile const double*)&HC_(H_LO, q4 + 4 * r); hup[r] = *(volatile const double*)&HC_(H_UP, q4 + 4 * r);
        }
        double r6[6];
        {
            const double c0m = (bot ? rb : 0.0) + hp_m;
            const double tx_ = vx + sx16(vx), tu_ = vu + sx16(vu), t0m = c0m + sx16(c0m), t0o = hp_o + sx16(hp_o);
            r6[0] = tx_; r6[1] = tu_; r6[2] = t0m + sx8(t0o);
            r6[3] = sx8(tx_); r6[4] = sx8(tu_); r6[5] = sx8(t0m) + t0o;
        }
        double Xs = 0.0, Us = 0.0, U0m = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            Xs = fma(mi[k], r6[k], Xs);
            Us = fma(mi[6 + k], r6[k], Us);
            U0m = fma(mi[12 + k], r6[k], U0m);
        }
        const double U0o = sx8(U0m);
        // ---- hull rows (z~ = A x~): update_z / update_y of rows q4 and q4 + 4
        {
            const double ux = a == 0 ? U0m : U0o, uy = a == 0 ? U0o : U0m;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const double zt = hcx[r] * ux + hcy[r] * uy;
                const double zr = alpha * zt + oma * zh[r];
                double zn = zr + rho_inv * yh[r];
                zn = zn > hlo[r] ? zn : hlo[r]; zn = zn < hup[r] ? zn : hup[r]; // c_min(c_max(.))
                const double dy = rho * (zr - zn);
                if (!PHASE2 || act) { zh[r] = zn; yh[r] += dy; dyh[r] = dy; }
            }
        }
        // (the multipliers both sweeps use are read again here as VOLATILE loads: left to itself ptxas keeps some twenty of them
        // in registers from the forward sweep and parks parts of the iterate in local memory to make room)
        // ---- backward: L' x~ = D^-1 t away from the separator, fused with update_y (equality rows) and update_x
        {
            const double ub = bot ? U0m : 0.0;
            double sx = 0.0, su = 0.0;               // x~ of the next slot (towards the separator)
#pragma unroll
            for (int t = T - 1; t >= 0; t--) {
                double xu, xx, sxr;
                if (t == T - 1) {
                    xu = Us;
                    const double xe = fma(-SV_(A_L1X, t), Us, fma(-SV_(A_L2X, t), Xs, tX[t] * S_(A_DIX, t)));
                    xx = bot ? fma(-LBXV_(T - 1 < 0 ? 0 : T - 1), ub, xe) : Xs;
                    sxr = sx16(xx);                  // top: x_{s-1} of the bottom lane; bottom: x_s
                } else {
                    double pu = fma(-SV_(A_L2U, t), su, tU[t] * S_(A_DIU, t));
                    double px = fma(-SV_(A_L2X, t), sx, tX[t] * S_(A_DIX, t));
                    if (bot) { pu = fma(-LBUV_(t), ub, pu); px = fma(-LBXV_(t), ub, px); }
                    xu = fma(-SV_(A_L1U, t), sx, pu);
                    if (t == OFF) { if (bot) xu = U0m; }
                    xx = fma(-SV_(A_L1X, t), xu, px);
                    sxr = sx;
                }
                // slot T-1: the two lanes of the axis hold copies of row s; the sum of two rounded products is
                // the same number on both (their operands arrive in opposite order), a contracted fma is not
                const double row = (t == T - 1) ? fma(b_, xu, __dadd_rn(__dmul_rn(cx0, xx), __dmul_rn(cx1, sxr)))
                                                : fma(b_, xu, cx0 * xx + cx1 * sxr);
                const double yn = fma(arho, S_(A_EE, t) * row, Y[t]);
                const double xn = fma(alpha, xx, oma * X[t]);
                const double un = fma(alpha, xu, oma * U[t]);
                if (!PHASE2 || act) {
                    tX[t] = xn - X[t]; tU[t] = un - U[t];
                    Y[t] = yn; X[t] = xn; U[t] = un;
                }
                if (t == OFF) {
                    if (bot && (!PHASE2 || act)) { // equality row 0: z = l = u = -dcm
                        const double W0 = rho_e * EE0;
                        dY0 = W0 * (fma(-alpha, xx, oma * Z0) + dcm_a);
                        Y0 += dY0;
                        Z0 = -dcm_a;
                        Yext = fma(W0, dcm_a, Y0);
                    }
                }
                sx = xx; su = xu;
            }
        }
        if (!PHASE2 || act) hull_partials();
    };

    // update_info + the norms of check_termination for the CURRENT (x, z, y), all in unscaled quantities;
    // SC: also the norms of the SCALED residuals (compute_rho_estimate; an adaptive-rho iteration only).
    // SC is a compile-time tag (std::true_type / std::false_type): as a run-time flag every slot's scaled part was a branch over
    // code that is never fetched otherwise, and each of those jumps landed on a cold instruction-cache line.
    // The slot body is written once (info_slot) and used in two forms: unrolled over the iterate in registers (phase 2, which goes
    // back into the ADMM loop afterwards), and ROLLED over a copy of the iterate in shared memory (phase 1, where the factor arrays
    // are dead once the iterations are done): 13 trips through ~100 instructions that stay in the L0 instruction cache instead of
    // ~1 500 instructions fetched once from L2 by every warp.
    struct InfoAcc {
        double rp_u = 0, rp_s = 0, z_u = 0, z_s = 0, ax_u = 0, ax_s = 0;
        double rd_u = 0, rd_s = 0, q_u = 0, q_s = 0, aty_u = 0, aty_s = 0, px_u = 0, px_s = 0;
    };
    // workspace rows of the lane's slot 0 and the step per slot: recomputed where they are used (macros, not variables: as
    // variables they are live across the ADMM loop, eight registers the loop does not have)
#define kx0_ (bot ? -OFF : N)
#define sB_ (bot ? (ptrdiff_t)B : -(ptrdiff_t)B)
#define xrow0_ (((ptrdiff_t)(a * nA + kx0_)) * (ptrdiff_t)B + i)
#define urow0_ (((ptrdiff_t)(a * nA + nR + kx0_ - (bot ? 0 : 1))) * (ptrdiff_t)B + i)
#define erow0_ (((ptrdiff_t)(a * nR + kx0_ + (bot ? 1 : 0))) * (ptrdiff_t)B + i)
    // slot t: Xt, Ut, Yt its iterate; Xn / Un the x / u of slot t + 1 (the partner lane's at the separator), Ua the u of slot t - 1
    // (0 before the first), Yp the multiplier of the previous row (Y0 for the bottom lane's first real slot, 0 before the first)
    auto info_slot = [&](auto sc_tag, const int t, const double Xt, const double Xn, const double Ua, const double Ut, const double Un,
                         const double Yt, const double Yp, const double hy_m, InfoAcc& A) {
        constexpr bool SC = decltype(sc_tag)::value;
        const bool real = (t >= OFF) || !bot;
        const bool shadow = bot && t == T - 1;       // the bottom lane's copy of (u_{s-1}, row s): counted by the top lane
        double Dx = 1.0, Du = 1.0, Ek = 1.0;
        if (SC) { if (real) { Dx = ldc(d.D + (xrow0_ + t * sB_)); Du = ldc(d.D + (urow0_ + t * sB_)); Ek = ldc(d.Eeq + (erow0_ + t * sB_)); } }
        if (real) { // equality row of the slot: z = l = u = 0
            const double ax = fma(b_, Ut, cx0 * Xt + cx1 * Xn);
            A.rp_u = amax(A.rp_u, ax); A.ax_u = amax(A.ax_u, ax);
            if (SC) { A.rp_s = amax(A.rp_s, Ek * ax); A.ax_s = amax(A.ax_s, Ek * ax); }
        }
        { // x unknown
            const double aty = -(cy0 * Yt + cy1 * Yp);
            const double px = cQ * Xt, q = S_(A_CQ, t);
            const double rr = q + px + aty;
            A.rd_u = amax(A.rd_u, rr); A.q_u = amax(A.q_u, q);
            A.aty_u = amax(A.aty_u, aty); A.px_u = amax(A.px_u, px);
            if (SC) {
                A.rd_s = amax(A.rd_s, Dx * rr); A.q_s = amax(A.q_s, Dx * q);
                A.aty_s = amax(A.aty_s, Dx * aty); A.px_s = amax(A.px_s, Dx * px);
            }
        }
        if (!shadow && real) { // u unknown (a dummy's row would pick up its neighbour u_0)
            const double sc = (t == 0) ? (bot ? 2.0 : 1.0) : 2.0;
            const double px = cR * ((sc * Ut - Ua) - Un);
            double aty = b_ * Yt;
            double q = 0.0;
            if (t == OFF) { if (bot) { aty += hy_m; q = cqu0; } }
            const double rr = q + px + aty;
            A.rd_u = amax(A.rd_u, rr); A.q_u = amax(A.q_u, q);
            A.aty_u = amax(A.aty_u, aty); A.px_u = amax(A.px_u, px);
            if (SC) {
                A.rd_s = amax(A.rd_s, Du * rr); A.q_s = amax(A.q_s, Du * q);
                A.aty_s = amax(A.aty_s, Du * aty); A.px_s = amax(A.px_s, Du * px);
            }
        }
    };
    // hull rows' share of (A' y) on u_0 of this lane's axis
    auto info_hull_aty = [&]() -> double {
        double hyx = 0.0, hyy = 0.0;
#pragma unroll
        for (int r = 0; r < 2; r++) { hyx += HC_(H_HXS, q4 + 4 * r) * yh[r]; hyy += HC_(H_HYS, q4 + 4 * r) * yh[r]; }
        hyx = inst_sum(hyx); hyy = inst_sum(hyy);
        return a == 0 ? hyx : hyy;
    };
    // equality row 0, the hull rows, and the reductions over the four lanes of the instance
    auto info_finish = [&](auto sc_tag, InfoAcc& A) {
        constexpr bool SC = decltype(sc_tag)::value;
        if (bot) { // equality row 0
            const double ax = -X[OFF], z = Z0, r = ax - z;
            A.rp_u = amax(A.rp_u, r); A.z_u = amax(A.z_u, z); A.ax_u = amax(A.ax_u, ax);
            if (SC) { A.rp_s = amax(A.rp_s, E0 * r); A.z_s = amax(A.z_s, E0 * z); A.ax_s = amax(A.ax_s, E0 * ax); }
        }
        { // hull rows
            const double u0x = __shfl_sync(FULL, U[OFF], 16 + s8), u0y = __shfl_sync(FULL, U[OFF], 24 + s8);
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int h = q4 + 4 * r;
                const double ax = HC_(H_HXS, h) * u0x + HC_(H_HYS, h) * u0y;
                const double z = zh[r], einv = HC_(H_EHI, h), rr = ax - z;
                A.rp_u = amax(A.rp_u, einv * rr); A.z_u = amax(A.z_u, einv * z); A.ax_u = amax(A.ax_u, einv * ax);
                if (SC) { A.rp_s = amax(A.rp_s, rr); A.z_s = amax(A.z_s, z); A.ax_s = amax(A.ax_s, ax); }
            }
        }
        pri_res = inst_amax(A.rp_u); nz_u = inst_amax(A.z_u); nAx_u = inst_amax(A.ax_u);
        dua_res = cinv * inst_amax(A.rd_u); nq_u = cinv * inst_amax(A.q_u);
        nAty_u = cinv * inst_amax(A.aty_u); nPx_u = cinv * inst_amax(A.px_u);
        if (SC) {
            rp_s = inst_amax(A.rp_s); nz_s = inst_amax(A.z_s); nAx_s = inst_amax(A.ax_s);
            rd_s = inst_amax(A.rd_s); nq_s = inst_amax(A.q_s); nAty_s = inst_amax(A.aty_s); nPx_s = inst_amax(A.px_s);
        }
    };
    // X, U, Y go to three arrays of the dump area (the other three receive 1 / D and 1 / E for the epilogue, requested
    // asynchronously before this is called) and the slots are walked by a real loop.
    constexpr int oD = PHASE2 ? SD::oDump2 : SD::oDump1;
#define DX_(t) sl[oD + (0 * T + (t)) * 32]
#define DU_(t) sl[oD + (1 * T + (t)) * 32]
#define DY_(t) sl[oD + (2 * T + (t)) * 32]
#define DIX_(t) sl[oD + (3 * T + (t)) * 32]
#define DIU_(t) sl[oD + (4 * T + (t)) * 32]
#define DIE_(t) sl[oD + (5 * T + (t)) * 32]
    auto update_info_rolled = [&](auto sc_tag) {
        InfoAcc A;
        const double Xpart = sx16(X[T - 1]);
        const double Upart = sx16(U[T >= 2 ? T - 2 : 0]);
        const double hy_m = info_hull_aty();
#pragma unroll
        for (int t = 0; t < T; t++) { DX_(t) = X[t]; DU_(t) = U[t]; DY_(t) = Y[t]; }
#pragma unroll 1
        for (int t = 0; t < T; t++) {
            const int tn = t < T - 1 ? t + 1 : t, tp = t > 0 ? t - 1 : 0;
            const double xn = DX_(tn), un = DU_(tn), ua = DU_(tp), ypv = DY_(tp);
            const double Xn = (t < T - 1) ? xn : Xpart;
            const double Un = (t < T - 1) ? un : Upart;
            const double Ua = (t >= 1) ? ua : 0.0;
            const double Yp = (t == OFF && bot) ? Y0 : (t == 0 ? 0.0 : ypv);
            info_slot(sc_tag, t, DX_(t), Xn, Ua, DU_(t), Un, DY_(t), Yp, hy_m, A);
        }
        info_finish(sc_tag, A);
    };


    // is_primal_infeasible (OSQP auxil.c) on delta_y of the last iteration, recomputed from X+ and dX
    // (X~ = X+ + (1/alpha - 1) dX; delta Yh_k = alpha rho_e E_k^2 (A X~)_k on the equality rows)
    auto primal_infeasible = [&](double eps) -> int {
        const double arho = alpha * kRhoEqOverIneq * rho;
        double dY[T];
        {
            const double xtl = fma(k1, tX[T - 1], X[T - 1]);
            const double xpart = sx16(xtl);
#pragma unroll
            for (int t = 0; t < T; t++) {
                const double xt = fma(k1, tX[t], X[t]), ut = fma(k1, tU[t], U[t]);
                const double xn_ = (t < T - 1) ? fma(k1, tX[t < T - 1 ? t + 1 : t], X[t < T - 1 ? t + 1 : t]) : xpart;
                dY[t] = arho * (S_(A_EE, t) * fma(b_, ut, cx0 * xt + cx1 * xn_));
            }
        }
        double nrm = 0.0, lhs = 0.0;
#pragma unroll
        for (int t = 0; t < T; t++) nrm = amax(nrm, dY[t]);
        if (bot) { nrm = amax(nrm, dY0); lhs += (-dcm_a) * dY0; }
        double dyp[2];
#pragma unroll
        for (int r = 0; r < 2; r++) { // project delta_y of the hull rows on the polar of the recession cone
            const int h = q4 + 4 * r;
            double dy = dyh[r];
            const double e = HC_(H_EH, h), uh = HC_(H_UP, h), lh = HC_(H_LO, h);
            const bool u_inf = uh > infty * kMinScaling, l_inf = lh < -infty * kMinScaling;
            if (u_inf) dy = l_inf ? 0.0 : fmin(dy, 0.0);
            else if (l_inf) dy = fmax(dy, 0.0);
            if (h >= nh) dy = 0.0;
            dyp[r] = dy;
            nrm = amax(nrm, e * dy);
            lhs += uh * fmax(dy, 0.0) + lh * fmin(dy, 0.0);
        }
        nrm = inst_max(nrm);
        lhs = inst_sum(lhs);
        int res = 0;
        double hdx = 0.0, hdy = 0.0;
#pragma unroll
        for (int r = 0; r < 2; r++) { hdx += HC_(H_HXS, q4 + 4 * r) * dyp[r]; hdy += HC_(H_HYS, q4 + 4 * r) * dyp[r]; }
        hdx = inst_sum(hdx); hdy = inst_sum(hdy);
        double m = 0.0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            double yp;
            if (t == OFF) yp = bot ? dY0 : (t == 0 ? 0.0 : dY[t > 0 ? t - 1 : 0]);
            else yp = (t == 0) ? 0.0 : dY[t > 0 ? t - 1 : 0];
            m = amax(m, cy0 * dY[t] + cy1 * yp);
            double aty = b_ * dY[t];
            if (t == OFF) { if (bot) aty += (a == 0 ? hdx : hdy); }
            if (!(bot && t == T - 1)) m = amax(m, aty);
        }
        m = inst_max(m);
        if (nrm > 1.0 / infty && lhs < -eps * nrm) res = m < eps * nrm;
        return res;
    };
    // is_dual_infeasible on delta_x of the last iteration (dX, dU in tX / tU)
    auto dual_infeasible = [&](double eps) -> int {
        double nrm = 0.0, qdx = 0.0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            nrm = amax(nrm, tX[t]);
            qdx = fma(S_(A_CQ, t), tX[t], qdx);
            if (!(bot && t == T - 1)) nrm = amax(nrm, tU[t]);
        }
        if (bot) qdx = fma(cqu0, tU[OFF], qdx);
        nrm = inst_max(nrm);
        qdx = inst_sum(qdx);
        const double Xpart = sx16(tX[T - 1]);
        const double Upart = sx16(tU[T >= 2 ? T - 2 : 0]);
        double m = 0.0;
        int bad = 0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            m = amax(m, cQ * tX[t]);
            if (!(bot && (t == T - 1 || t < OFF))) {
                const double ua = (t >= 1) ? tU[t >= 1 ? t - 1 : 0] : 0.0;
                const double ubn = (t < T - 1) ? tU[t < T - 1 ? t + 1 : t] : Upart;
                const double sc = (t == 0) ? (bot ? 2.0 : 1.0) : 2.0;
                m = amax(m, cR * ((sc * tU[t] - ua) - ubn));
            }
            const double xn_ = (t < T - 1) ? tX[t < T - 1 ? t + 1 : t] : Xpart;
            const double ax = fma(b_, tU[t], cx0 * tX[t] + cx1 * xn_);
            if (((t >= OFF) || !bot) && (ax > eps * nrm || ax < -eps * nrm)) bad = 1; // both bounds finite on equality rows
        }
        if (bot) { const double ax = -tX[OFF]; if (ax > eps * nrm || ax < -eps * nrm) bad = 1; }
        {
            const double u0x = __shfl_sync(FULL, tU[OFF], 16 + s8), u0y = __shfl_sync(FULL, tU[OFF], 24 + s8);
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int h = q4 + 4 * r;
                const double ax = (HC_(H_HXS, h) * u0x + HC_(H_HYS, h) * u0y) * HC_(H_EHI, h);
                const double uh = HC_(H_UP, h), lh = HC_(H_LO, h);
                if (h < nh && ((uh < infty * kMinScaling && ax > eps * nrm) || (lh > -infty * kMinScaling && ax < -eps * nrm))) bad = 1;
            }
        }
        m = inst_max(m);
        bad = (int)inst_max((double)bad);
        int res = 0;
        if (nrm > 1.0 / infty && qdx < -c * eps * nrm) {
            if (m < c * eps * nrm) res = !bad;
        }
        return res;
    };
    // check_termination (OSQP auxil.c); called by the whole warp, returns the status to stop with or 0
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
        if (__any_sync(FULL, need_pi)) { const int r = primal_infeasible(epi); if (need_pi) prim_inf = r; }
        if (__any_sync(FULL, need_di)) { const int r = dual_infeasible(edi); if (need_di) dual_inf = r; }
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
    // phase 1 stops after this many iterations (first termination check / rho update / max_iter)
    int phase1_len = chk > 0 ? chk : max_iter;
    if (adaptive && rint < phase1_len) phase1_len = rint;
    if (phase1_len > max_iter) phase1_len = max_iter;
    int iter = PHASE2 ? phase1_len : 0;
    int to_check = chk ? chk - (iter % chk) : 0;
    int to_adapt = adaptive ? rint - (iter % rint) : 0;
    // the adaptive-rho step of osqp_solve (compute_rho_estimate + adapt_rho) from the scaled norms of update_info
    auto adapt_rho = [&]() {
        const double pr = rp_s / (fmax(nz_s, nAx_s) + 1e-10);
        const double dr = rd_s / (fmax(fmax(nq_s, nAty_s), nPx_s) + 1e-10);
        double rn = rho * sqrt(pr / (dr + 1e-10));
        rn = fmin(fmax(rn, kRhoMin), kRhoMax);
        if (rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance) {
            rho = rn;
            rho_updates++;
        }
    };
    // 1 / D and 1 / E of the lane's slots for the epilogue, by cp.async (no register is involved) into the dump area: phase 1 asks
    // for them when its iterations are over, phase 2 at once
    auto request_scalings = [&]() {
        const unsigned s_dix = (unsigned)__cvta_generic_to_shared(&DIX_(0));
        const unsigned s_diu = (unsigned)__cvta_generic_to_shared(&DIU_(0));
        const unsigned s_die = (unsigned)__cvta_generic_to_shared(&DIE_(0));
        const double* gx = d.Dinv + xrow0_;
        const double* gu = d.Dinv + urow0_;
        const double* ge = d.Einv + erow0_;
        const ptrdiff_t sBv = sB_;
        if (!skip) {
#pragma unroll 1
            for (int t = 0; t < T; t++) {
                if ((t >= OFF) || !bot) {
                    if (LISTED) { // through L2 (see ldc)
                        DIX_(t) = __ldcg(gx + t * sBv); DIU_(t) = __ldcg(gu + t * sBv); DIE_(t) = __ldcg(ge + t * sBv);
                    } else {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s_dix + t * 256), "l"(gx + t * sBv) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s_diu + t * 256), "l"(gu + t * sBv) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s_die + t * 256), "l"(ge + t * sBv) : "memory");
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    if constexpr (!PHASE2) {
        // ---- phase 1 is ONE pass of osqp_solve's loop, written straight: factor, phase1_len iterations, one update_info, the
        // checks that fall on that iteration.  Everything after the iterations walks the slots with real loops over a copy of the
        // iterate in shared memory (update_info_rolled, the epilogue), so that the once-per-warp code is a few hundred
        // instructions instead of several thousand fetched from L2.
        if (__any_sync(FULL, !done)) {
        build_factor();
        Yext = Y0 - (kRhoEqOverIneq * rho * EE0) * Z0;
        hull_partials();
        const int run = phase1_len < 1 ? 1 : phase1_len;
        const bool act = !done;
        const double rho_inv = 1.0 / rho;
        // (the shared-memory constants are loop-invariant, and ptxas hoists ~20 of them into registers at the price of spilling
        // parts of the iterate to local memory inside the loop: the base pointers get a zero it cannot fold)
#pragma unroll 1
        for (int r = 0; r < run; r++) {
            const int z = r >> 30;
            sl += z; sb += z; smr += z; sh += z;
            iteration(act, rho_inv);
        }
        iter = run;
        if (act) iter_done = iter;
        const bool check_now = chk && run == chk;
        const bool adapt_now = adaptive && run == rint;
        const bool at_max = iter >= max_iter;
        request_scalings();
        if (adapt_now) update_info_rolled(std::true_type{}); else update_info_rolled(std::false_type{});
#pragma unroll 1
        for (int approx = 0; approx < 2; approx++) {
            if (approx == 0) {
                if (check_now || at_max) {
                    const int stop = check_termination(0);
                    if (act && !done && stop) { status = stop; done = true; }
                }
                if (adapt_now && act && !done) adapt_rho();
            } else if (at_max) { // osqp_solve leaves the loop: approximate criteria, else MAX_ITER_REACHED
                const int stop = check_termination(1);
                if (act && !done) { status = stop ? stop : WC_OSQP_MAX_ITER_REACHED; done = true; }
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        }
    } else {
    request_scalings();
    bool refactor = true;
    bool warp_busy = __any_sync(FULL, !done);
    // every rare block below has ONE call site (update_info, check_termination, build_factor are large once inlined:
    // the instruction footprint of the kernel decides how often the hot loop misses the instruction cache)
#pragma unroll 1
    while (warp_busy) {
        if (refactor) { // numeric factorisation: at entry, and again after a rho update (unchanged rho recomputes the same factors)
            build_factor();
            Yext = Y0 - (kRhoEqOverIneq * rho * EE0) * Z0;
            hull_partials();
            refactor = false;
        }
        // iterations up to the next termination check / rho update / max_iter
        int run = max_iter - iter;
        if (chk && to_check < run) run = to_check;
        if (adaptive && to_adapt < run) run = to_adapt;
        if (run < 1) run = 1;
        const bool act = !done;
        const double rho_inv = 1.0 / rho;
#pragma unroll 1
        for (int r = 0; r < run; r++) iteration(act, rho_inv);
        iter += run;
        if (act) iter_done = iter;
        bool check_now = false, adapt_now = false;
        if (chk) { to_check -= run; if (to_check == 0) { check_now = true; to_check = chk; } }
        if (adaptive) { to_adapt -= run; if (to_adapt == 0) { adapt_now = true; to_adapt = rint; } }
        const bool at_max = iter >= max_iter;
        // ---- termination / adaptive rho (instance-uniform decisions, warp-uniform control flow)
        if (adapt_now) update_info_rolled(std::true_type{}); else update_info_rolled(std::false_type{});
#pragma unroll 1
        for (int approx = 0; approx < 2; approx++) {
            if (approx == 0) {
                if (check_now || at_max) {
                    const int stop = check_termination(0);
                    if (act && !done && stop) { status = stop; done = true; }
                }
                if (adapt_now && act && !done) adapt_rho();
            } else if (at_max) { // osqp_solve leaves the loop: approximate criteria, else MAX_ITER_REACHED
                const int stop = check_termination(1);
                if (act && !done) { status = stop ? stop : WC_OSQP_MAX_ITER_REACHED; done = true; }
            }
        }
        if (adapt_now) refactor = true;
        warp_busy = __any_sync(FULL, !done);
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    }
    __syncwarp();

    // ------------------------------------------------------------------ epilogue: store_solution + controller checks
    const bool unfinished = !skip && !done; // phase 1 only
    const bool has_sol = !(status == WC_OSQP_PRIMAL_INFEASIBLE || status == WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                           status == WC_OSQP_DUAL_INFEASIBLE || status == WC_OSQP_DUAL_INFEASIBLE_INACCURATE ||
                           status == WC_OSQP_UNSOLVED || status == WC_OSQP_NON_CVX) || unfinished;
    if (!skip) { // back to OSQP's scaled workspace
        { // from the copy of the iterate and the scaling factors in shared memory (update_info_rolled, request_scalings)
            double* px = d.x + xrow0_;
            double* pu = d.x + urow0_;
            double* pe = d.yeq + erow0_;
            const ptrdiff_t sBv = sB_;
#pragma unroll 1
            for (int t = 0; t < T; t++) {
                const bool real = (t >= OFF) || !bot;
                const bool shadow = bot && t == T - 1;
                if (real) {
                    px[t * sBv] = has_sol ? DX_(t) * DIX_(t) : 0.0;
                    if (!shadow) {
                        pu[t * sBv] = has_sol ? DU_(t) * DIU_(t) : 0.0;
                        pe[t * sBv] = has_sol ? DY_(t) * DIE_(t) : 0.0;
                    }
                }
            }
        }
        if (bot) {
            d.yeq[(size_t)(a * nR) * B + i] = has_sol ? Y0 * ldc(d.Einv + ((size_t)(a * nR) * B + i)) : 0.0;
            d.z0[(size_t)a * B + i] = has_sol ? Z0 * E0 : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int h = q4 + 4 * r;
            d.yh[(size_t)h * B + i] = has_sol ? yh[r] : 0.0;
            d.zh[(size_t)h * B + i] = has_sol ? zh[r] : 0.0;
        }
    }
    // desired ZMP (= unscaled u_0) of both axes on every lane; the hull margin split over the four lanes
    const double zx = __shfl_sync(FULL, U[OFF], 16 + s8), zy = __shfl_sync(FULL, U[OFF], 24 + s8);
    const double upx = __shfl_sync(FULL, upa, s8), upy = __shfl_sync(FULL, upa, 8 + s8);
    // ConvexHullProjectionConstraint::computeMargin (Controller.cpp:513-517): margin < -tol  <=>  outside the polygon AND
    // farther than tol from its boundary.  A ZMP on an active hull row violates it by OSQP's tolerance (~1e-4 m); the hull
    // is the convex hull of foot rectangles (interior angles >= 90 degrees), so the distance to the polygon is at most
    // sqrt(2) x the largest normalised row violation: only a violation that could reach the tolerance takes the exact
    // point-to-segment distances.
    double vmax = 0.0;
    int outside = 0;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int h = q4 + 4 * r;
        if (h < nh) {
            const double a0 = __ldg(d.hullA + ((size_t)(2 * h) * B + i)), a1 = __ldg(d.hullA + ((size_t)(2 * h + 1) * B + i));
            const double viol = (a0 * zx + a1 * zy) - __ldg(d.hullb + ((size_t)h * B + i));
            if (!(viol <= 0.0)) { outside = 1; vmax = fmax(vmax, viol / sqrt(a0 * a0 + a1 * a1)); }
        }
    }
    vmax = inst_max(vmax);
    outside = (int)inst_max((double)outside);
    if (lane < 8 && valid && !pending) {
        if (unfinished) {
            const int idx = atomicAdd(d.cont_count, 1);
            d.cont_list[idx] = i;
            d.rho[i] = rho;
            d.rho_updates[i] = rho_updates;
            if (init_state == 2) d.inited[i] = 1;
        } else {
            int flags = kflags;
            double out[2] = {upx, upy};
            if (!skip) {
                d.rho[i] = rho;
                d.rho_updates[i] = rho_updates;
                if (init_state == 2) d.inited[i] = 1;
                if (osqp_eigen_solve_ok(status, st.osqp_eigen_checks_status)) {
                    out[0] = has_sol ? zx : NAN; out[1] = has_sol ? zy : NAN;
                    d.uprev[i] = out[0]; d.uprev[B + i] = out[1];
                    const bool inside = !outside;
                    double best = INFINITY;
                    const bool exact = !inside && !(4.0 * vmax < d.hull_tol);
                    if (exact) {
                        for (int h = 0; h < nh; h++) {
                            const int h1 = (h + 1) % nh;
                            const double p0x = d.hullv[(size_t)(2 * h) * B + i], p0y = d.hullv[(size_t)(2 * h + 1) * B + i];
                            const double p1x = d.hullv[(size_t)(2 * h1) * B + i], p1y = d.hullv[(size_t)(2 * h1 + 1) * B + i];
                            const double dx = p1x - p0x, dy = p1y - p0y;
                            const double len2 = dx * dx + dy * dy;
                            double tt = len2 > 0 ? ((out[0] - p0x) * dx + (out[1] - p0y) * dy) / len2 : 0.0;
                            tt = tt < 0 ? 0 : (tt > 1 ? 1 : tt);
                            const double ex = out[0] - (p0x + tt * dx), ey = out[1] - (p0y + tt * dy);
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
            // MODE 2 leaves the marker in place: a MODE 0 CTA that has not started yet must still see the instance as not its own
            // (K1 rewrites tick_flags on the next tick, phase 2 strips the bit)
            d.tick_flags[i] = LISTED ? (flags | kMpcFlagSetupPending) : flags;
        }
    }
    __syncwarp(); // the next group of this CTA reuses the shared-memory working set
    } // groups
#undef kx0_
#undef sB_
#undef xrow0_
#undef urow0_
#undef erow0_
#undef DX_
#undef DU_
#undef DY_
#undef DIX_
#undef DIU_
#undef DIE_
#undef SV_
#undef LBXV_
#undef LBUV_
#undef S_
#undef LBX_
#undef LBU_
#undef HC_
}

// ------------------------------------------------------------------------------------------------
template <int NT>
static cudaError_t launch_split_t(const MpcDev& d, const MpcTickIO& io, cudaStream_t s, cudaEvent_t* mid, cudaStream_t side,
                                  cudaEvent_t ev_fork, cudaEvent_t ev_join)
{
    constexpr size_t smem = (size_t)SplitDims<NT>::total * sizeof(double) * WC_SPLIT_WARPS;
    constexpr size_t smem2 = (size_t)SplitDims<NT>::total2 * sizeof(double) * WC_SPLIT_WARPS;
    auto k1 = mpc_split_kernel<0, NT>;
    auto k2 = mpc_split_kernel<1, NT>;
    auto k1l = mpc_split_kernel<2, NT>;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k1l, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int blocks = ((d.B + 7) / 8 + WC_SPLIT_WARPS - 1) / WC_SPLIT_WARPS;
    const int cap = (WC_SPLIT_MINB / WC_SPLIT_WARPS) * 148;
    if (d.setup_overlap) {
        // osqp_setup of the listed instances (~2 % per tick in a walking gait) and their phase 1 on the side stream, next to the
        // phase 1 of everyone else; both join before phase 2
        if ((e = cudaEventRecord(ev_fork, s)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(side, ev_fork, 0)) != cudaSuccess) return e;
        if ((e = launch_mpc_chain_setup(d, side, false)) != cudaSuccess) return e;
        k1l<<<blocks < cap ? blocks : cap, 32 * WC_SPLIT_WARPS, smem, side>>>(d, io); // grid-stride over the setup list
        if ((e = cudaEventRecord(ev_join, side)) != cudaSuccess) return e;
    } else {
        if ((e = launch_mpc_chain_setup(d, s, false)) != cudaSuccess) return e;
    }
    if (mid) cudaEventRecord(mid[0], s); // per-kernel timing (wc_mpc_enable_profiling)
    k1<<<blocks, 32 * WC_SPLIT_WARPS, smem, s>>>(d, io);
    if (d.setup_overlap && (e = cudaStreamWaitEvent(s, ev_join, 0)) != cudaSuccess) return e;
    if (mid) cudaEventRecord(mid[1], s);
    k2<<<blocks < cap ? blocks : cap, 32 * WC_SPLIT_WARPS, smem2, s>>>(d, io); // grid-stride over the continuation list
    return cudaGetLastError();
}

bool mpc_split_supported(int N) { return N == 25 || N == 10; }

cudaError_t launch_mpc_split(const MpcDev& d, const MpcTickIO& io, cudaStream_t s, cudaEvent_t* mid, cudaStream_t side, cudaEvent_t ev_fork,
                             cudaEvent_t ev_join)
{
    if (d.setup_overlap && (!side || !ev_fork || !ev_join)) return cudaErrorInvalidValue;
    switch (d.N) {
    case 25: return launch_split_t<25>(d, io, s, mid, side, ev_fork, ev_join);
    case 10: return launch_split_t<10>(d, io, s, mid, side, ev_fork, ev_join);
    default: return cudaErrorInvalidConfiguration;
    }
}

} // namespace wcb
