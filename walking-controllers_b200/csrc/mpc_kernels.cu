// mpc_kernels.cu — see mpc_kernels.cuh for the mapping to the reference.
#include "mpc_kernels.cuh"

#include <cmath>

namespace wcb {

__device__ __forceinline__ double limit_scaling(double v)
{
    v = v < kMinScaling ? 1.0 : v;
    v = v > kMaxScaling ? kMaxScaling : v;
    return v;
}

// =====================================================================================
// K1: per-instance problem build.  One thread per instance: contact-pair change detection, convex
// hull of the support polygon, gradient ring update.  (OSQP's setup-time Ruiz equilibration of a new
// solver runs cooperatively at the head of K2, where the instance's data is staged in shared memory.)
//   WalkingController::setConvexHullConstraint  (Controller.cpp:364-435)
//   buildConvexHull -> iDynTree ConvexHullProjectionConstraint (external; restated)
//   MPCSolver::MPCSolver / setGradient          (MPCSolver.cpp:16-52,183-266)
//   osqp_setup: scale_data + cold start          (external; restated from OSQP 0.6)
// =====================================================================================
__device__ double cross2(const double* o, const double* a, const double* b)
{
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0]);
}

__device__ int build_hull(const double (*pts)[2], int npts, double* A, double* b, double* v)
{
    double p[16][2], h[32][2];
    int k = 0;
    for (int i = 0; i < npts; i++) { p[i][0] = pts[i][0]; p[i][1] = pts[i][1]; }
    for (int i = 1; i < npts; i++) {
        const double x = p[i][0], y = p[i][1];
        int j = i - 1;
        while (j >= 0 && (p[j][0] > x || (p[j][0] == x && p[j][1] > y))) {
            p[j + 1][0] = p[j][0]; p[j + 1][1] = p[j][1]; j--;
        }
        p[j + 1][0] = x; p[j + 1][1] = y;
    }
    for (int i = 0; i < npts; i++) {
        while (k >= 2 && cross2(h[k - 2], h[k - 1], p[i]) <= 0) k--;
        h[k][0] = p[i][0]; h[k][1] = p[i][1]; k++;
    }
    for (int i = npts - 2, j = k + 1; i >= 0; i--) {
        while (k >= j && cross2(h[k - 2], h[k - 1], p[i]) <= 0) k--;
        h[k][0] = p[i][0]; h[k][1] = p[i][1]; k++;
    }
    k--;
    if (k > kMaxHull) k = kMaxHull;
    for (int i = 0; i < k; i++) {
        const double* p0 = h[i];
        const double* p1 = h[(i + 1) % k];
        A[2 * i] = p1[1] - p0[1];
        A[2 * i + 1] = -(p1[0] - p0[0]);
        b[i] = A[2 * i] * p0[0] + A[2 * i + 1] * p0[1];
        v[2 * i] = p0[0]; v[2 * i + 1] = p0[1];
    }
    return k;
}

__device__ void foot_points(const MpcDev& d, const double* H, size_t B, int i, double (*out)[2])
{
    const double vx[4] = {d.foot_front, -d.foot_back, -d.foot_back, d.foot_front};
    const double vy[4] = {d.foot_left, d.foot_left, -d.foot_right, -d.foot_right};
    const double r00 = H[0 * B + i], r01 = H[1 * B + i], px = H[3 * B + i];
    const double r10 = H[4 * B + i], r11 = H[5 * B + i], py = H[7 * B + i];
    for (int k = 0; k < 4; k++) {
        out[k][0] = r00 * vx[k] + r01 * vy[k] + px;
        out[k][1] = r10 * vx[k] + r11 * vy[k] + py;
    }
}

__global__ void __launch_bounds__(128) mpc_prepare_kernel(MpcDev d, MpcTickIO io)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.B) return;
    if (i == 0 && d.cont_count) *d.cont_count = 0; // consumed by phase 2 of the previous tick
    const size_t B = (size_t)d.B;
    const int N = d.N, nR = N + 1;
    const uint8_t cur = (uint8_t)((io.l_contact[i] ? 1 : 0) | (io.r_contact[i] ? 2 : 0));
    int flags = 0;
    // "None foot is in contact" (Controller.cpp:404-408): the reference returns false on the tick of the transition AND,
    // because m_feetStatus then equals (false, false), keeps solving with the stale hull on the following no-contact ticks.
    // Here the instance is flagged on every tick without contact and not solved: a ring slot that was never refreshed and
    // a stale hull must not produce an output that looks valid.  inited = 0 forces a full gradient refresh (and a new OSQP
    // workspace) when contact returns, as the reference's new MPCSolver after the next phase change does.
    if (cur == 0) {
        d.feet[i] = cur;
        d.inited[i] = 0;
        d.tick_flags[i] = WC_MPC_FLAG_NO_CONTACT;
        return;
    }
    if (cur != d.feet[i]) {
        d.feet[i] = cur;
        double pts[8][2], A[16], b[8], v[16];
        int npts = 0;
        if (cur == 3) { foot_points(d, io.lf_H, B, i, pts); foot_points(d, io.rf_H, B, i, pts + 4); npts = 8; }
        else if (cur == 1) { foot_points(d, io.lf_H, B, i, pts); npts = 4; }
        else { foot_points(d, io.rf_H, B, i, pts); npts = 4; }
        const int nh = build_hull(pts, npts, A, b, v);
        d.nhull[i] = nh;
        for (int h = 0; h < kMaxHull; h++) {
            const bool on = h < nh;
            d.hullA[(2 * h) * B + i] = on ? A[2 * h] : 0.0;
            d.hullA[(2 * h + 1) * B + i] = on ? A[2 * h + 1] : 0.0;
            d.hullb[h * B + i] = on ? b[h] : 0.0;
            d.hullv[(2 * h) * B + i] = on ? v[2 * h] : 0.0;
            d.hullv[(2 * h + 1) * B + i] = on ? v[2 * h + 1] : 0.0;
        }
        d.have_solver[i] = 1;
        d.inited[i] = 0;
        flags |= WC_MPC_FLAG_NEW_SOLVER;
    }
    if (!d.have_solver[i]) { d.tick_flags[i] = flags | WC_MPC_FLAG_NO_SOLVER; return; }
    const bool inited = d.inited[i] != 0;
    if (!inited && d.setup_list) { // osqp_setup pending
        d.setup_list[atomicAdd(d.setup_count, 1)] = i;
        if (d.setup_overlap) flags |= kMpcFlagSetupPending;
    }
    const bool reset = io.reset ? (io.reset[i] != 0) : false;
    const double Q[2] = {d.Qd[i], d.Qd[B + i]};
    const int head = mpc_ring_head(d);

    // ---- MPCSolver::setGradient (state block; the input block -R u_prev is formed in K2)
    // reference sample k of this tick's window: the caller's array, or the device-side ring
    const double* rbase = io.ref_is_ring ? d.ref_ring : io.ref;
    auto ref_row = [&](int k) { return io.ref_is_ring ? (head + k) % nR : k; };
    if (!inited || reset) {
        for (int k = 0; k <= N; k++) {
            const int slot = (head + k) % nR;
            const int rr = ref_row(k);
            const double rx = rbase[(size_t)(2 * rr) * B + i], ry = rbase[(size_t)(2 * rr + 1) * B + i];
            d.grad[(size_t)(slot)*B + i] = d.coupled ? -(Q[0] * rx + d.Qoff[0] * ry) : -Q[0] * rx;
            d.grad[(size_t)(nR + slot) * B + i] = d.coupled ? -(d.Qoff[1] * rx + Q[1] * ry) : -Q[1] * ry;
        }
    } else {
        const int slot = (head + N) % nR;
        const int rr = ref_row(N);
        const double rx = rbase[(size_t)(2 * rr) * B + i], ry = rbase[(size_t)(2 * rr + 1) * B + i];
        d.grad[(size_t)(slot)*B + i] = d.coupled ? -(Q[0] * rx + d.Qoff[0] * ry) : -Q[0] * rx;
        d.grad[(size_t)(nR + slot) * B + i] = d.coupled ? -(d.Qoff[1] * rx + Q[1] * ry) : -Q[1] * ry;
    }

    d.tick_flags[i] = flags;
}

// the window as a ring over stages: one new sample per tick (slot of stage N), or the whole window after a replan
__global__ void __launch_bounds__(128) mpc_ref_ring_kernel(MpcDev d, const double* newest, const double* window)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.B) return;
    const size_t B = (size_t)d.B;
    const int N = d.N, nR = N + 1, head = mpc_ring_head(d);
    if (window) {
        for (int k = 0; k <= N; k++) {
            const int slot = (head + k) % nR;
            d.ref_ring[(size_t)(2 * slot) * B + i] = window[(size_t)(2 * k) * B + i];
            d.ref_ring[(size_t)(2 * slot + 1) * B + i] = window[(size_t)(2 * k + 1) * B + i];
        }
    } else {
        const int slot = (head + N) % nR;
        d.ref_ring[(size_t)(2 * slot) * B + i] = newest[i];
        d.ref_ring[(size_t)(2 * slot + 1) * B + i] = newest[B + i];
    }
}

void launch_mpc_ref_ring_update(const MpcDev& d, const double* newest, const double* window, cudaStream_t s)
{
    const int threads = 128;
    mpc_ref_ring_kernel<<<(d.B + threads - 1) / threads, threads, 0, s>>>(d, newest, window);
}

void launch_mpc_prepare(const MpcDev& d, const MpcTickIO& io, cudaStream_t s)
{
    const int threads = 128;
    const int blocks = (d.B + threads - 1) / threads;
    mpc_prepare_kernel<<<blocks, threads, 0, s>>>(d, io);
}

// =====================================================================================
// K2: batched OSQP-equivalent ADMM.  G lanes per instance, 32/G instances per warp, every warp
// independent (only __syncwarp).  Per-axis vectors are stored "split": index a*nA + k for the state
// x_k (k = 0..N) and a*nA + nR + k for the input u_k (k = 0..N-1), nA = 2N+1, nR = N+1, so that all
// cooperative accesses are stride-1 in the stage index k.  Shared memory per instance (doubles):
//   X  [2nA] iterate          QB [2nA] scaled q          W  [2nA] rhs / x~ / delta_x
//   L1 [2nA] L2 [2nA] DI [2nA]  pentadiagonal LDL^T (chain order x_N,u_{N-1},...,x_1 | u_0,x_0)
//   DD [2nA] Ruiz D            YE [2nR] y eq   DYE [2nR] delta_y eq   CM/CA/CB [2nR] scaled A   EE [2nR]
//   HX,HY,UH,LH,ZH,YH,DYH,EH [8]  hull rows     CF [16] corner factor   RS [8] scratch
// =====================================================================================
struct Smem {
    double *X, *QB, *W, *L1, *L2, *DI, *DD, *YE, *DYE, *CM, *CA, *CB, *EE;
    double *HX, *HY, *UH, *LH, *ZH, *YH, *DYH, *EH, *CF, *RS;
    double *ZP, *RT, *RED;   // segment-parallel solve only: work vector [2nA], separator right-hand-side parts [2][16][4], reduced factors [2][16][8]
    double *L1C, *L2C, *DIC; // coupled weights only: entry (a, other axis) of the 2x2 blocks of L1, L2, D^-1, at the split index of (a, position)
};

__device__ __forceinline__ Smem carve(double* base, int N, int coupled, int par)
{
    const int nA = 2 * N + 1, nR = N + 1;
    Smem s;
    double* p = base;
    s.X = p; p += 2 * nA; s.QB = p; p += 2 * nA; s.W = p; p += 2 * nA;
    s.L1 = p; p += 2 * nA; s.L2 = p; p += 2 * nA; s.DI = p; p += 2 * nA; s.DD = p; p += 2 * nA;
    s.YE = p; p += 2 * nR; s.DYE = p; p += 2 * nR; s.CM = p; p += 2 * nR; s.CA = p; p += 2 * nR;
    s.CB = p; p += 2 * nR; s.EE = p; p += 2 * nR;
    s.HX = p; p += 8; s.HY = p; p += 8; s.UH = p; p += 8; s.LH = p; p += 8;
    s.ZH = p; p += 8; s.YH = p; p += 8; s.DYH = p; p += 8; s.EH = p; p += 8;
    s.CF = p; p += 16; s.RS = p; p += 8;
    s.L1C = s.L2C = s.DIC = nullptr;
    if (coupled) { s.L1C = p; p += 2 * nA; s.L2C = p; p += 2 * nA; s.DIC = p; p += 2 * nA; }
    s.ZP = s.RT = s.RED = nullptr;
    if (par) { s.ZP = p; p += 2 * nA; s.RT = p; p += 2 * 16 * 4; s.RED = p; }
    return s;
}

template <int G>
__device__ __forceinline__ double group_max(double v, unsigned mask)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o, 32));
    return v;
}
template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned mask)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, 32);
    return v;
}

// Pentadiagonal LDL^T of one axis chain (x_N, u_{N-1}, x_{N-1}, ..., x_1) + Schur complement on its
// (u_0, x_0) corner.  Rows are stored at the split index of their variable.  out3 = (S11, S12, S22).
__device__ void factor_axis(const Smem& s, int a, int N, double cQ, double cR, double sigma, double rho_e,
                            double hull_diag, double* out3)
{
    const int nA = 2 * N + 1, nR = N + 1;
    const double* __restrict__ CM = s.CM + a * nR;
    const double* __restrict__ CA = s.CA + a * nR;
    const double* __restrict__ CB = s.CB + a * nR;
    const double* __restrict__ DX = s.DD + a * nA;
    const double* __restrict__ DU = s.DD + a * nA + nR;
    double* __restrict__ L1X = s.L1 + a * nA; double* __restrict__ L1U = s.L1 + a * nA + nR;
    double* __restrict__ L2X = s.L2 + a * nA; double* __restrict__ L2U = s.L2 + a * nA + nR;
    double* __restrict__ DIX = s.DI + a * nA; double* __restrict__ DIU = s.DI + a * nA + nR;
    // pivots / sub-diagonals of the previous two chain positions
    double Dm1 = 1.0, Dm2 = 1.0, l1m1 = 0.0, e1m1 = 0.0, e2m2 = 0.0, e2m1 = 0.0;
    for (int k = N; k >= 1; k--) {
        { // x_k : K entries  d, e1 (with u_{k-1}), e2 (with x_{k-1})
            const double dx = DX[k];
            const double dj = cQ * dx * dx + sigma + rho_e * (CM[k] * CM[k] + (k < N ? CA[k + 1] * CA[k + 1] : 0.0));
            const double e1j = rho_e * CM[k] * CB[k], e2j = rho_e * CM[k] * CA[k];
            const double l2 = e2m2 / Dm2;
            const double l1 = (e1m1 - l2 * Dm2 * l1m1) / Dm1;
            const double Dj = dj - l1 * l1 * Dm1 - l2 * l2 * Dm2;
            L1X[k] = l1; L2X[k] = l2; DIX[k] = 1.0 / Dj;
            e2m2 = e2m1; e2m1 = e2j; e1m1 = e1j; Dm2 = Dm1; Dm1 = Dj; l1m1 = l1;
        }
        if (k >= 2) { // u_{k-1} : d, e1 (with x_{k-1}), e2 (with u_{k-2})
            const int ku = k - 1;
            const double du = DU[ku];
            const double sc = (ku < N - 1) ? 2.0 : 1.0;
            const double dj = cR * sc * du * du + sigma + rho_e * CB[ku + 1] * CB[ku + 1];
            const double e1j = rho_e * CA[ku + 1] * CB[ku + 1], e2j = -cR * du * DU[ku - 1];
            const double l2 = e2m2 / Dm2;
            const double l1 = (e1m1 - l2 * Dm2 * l1m1) / Dm1;
            const double Dj = dj - l1 * l1 * Dm1 - l2 * l2 * Dm2;
            L1U[ku] = l1; L2U[ku] = l2; DIU[ku] = 1.0 / Dj;
            e2m2 = e2m1; e2m1 = e2j; e1m1 = e1j; Dm2 = Dm1; Dm1 = Dj; l1m1 = l1;
        }
    }
    // corner rows: u_0 then x_0  (Dm1 = pivot of x_1, Dm2 = pivot of u_1)
    {
        const double du = DU[0];
        const double sc = (0 < N - 1) ? 2.0 : 1.0;
        const double dj = cR * sc * du * du + sigma + rho_e * CB[1] * CB[1] + hull_diag;
        const double e1j = rho_e * CA[1] * CB[1]; // K[u_0, x_0]
        const double l2 = e2m2 / Dm2;
        const double l1 = (e1m1 - l2 * Dm2 * l1m1) / Dm1;
        L1U[0] = l1; L2U[0] = l2; DIU[0] = 0.0;
        out3[0] = dj - l1 * l1 * Dm1 - l2 * l2 * Dm2;   // S11
        const double l2x = e2m1 / Dm1;                  // L[x_0, x_1]
        out3[1] = e1j - l2x * Dm1 * l1;                 // S12
        L1X[0] = 0.0; L2X[0] = l2x; DIX[0] = 0.0;
        const double dx = DX[0];
        const double d0 = cQ * dx * dx + sigma + rho_e * (CM[0] * CM[0] + CA[1] * CA[1]);
        out3[2] = d0 - l2x * l2x * Dm1;                 // S22
    }
}

// dense 4x4 LDL^T of the corner, order [u0x, u0y, x0x, x0y]; CF: L (6 strict-lower, row-major) + 4 Dinv
__device__ void factor_corner(const double* Sx, const double* Sy, double kxy, double* CF)
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

__device__ __forceinline__ void solve_corner(const double* CF, double* r)
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

// forward substitution on one axis chain; leaves the reduced corner right-hand sides in W[u_0], W[x_0]
__device__ __forceinline__ void sweep_forward(const Smem& s, int a, int N)
{
    const int nA = 2 * N + 1, nR = N + 1;
    double* __restrict__ WX = s.W + a * nA; double* __restrict__ WU = s.W + a * nA + nR;
    const double* __restrict__ L1X = s.L1 + a * nA; const double* __restrict__ L1U = s.L1 + a * nA + nR;
    const double* __restrict__ L2X = s.L2 + a * nA; const double* __restrict__ L2U = s.L2 + a * nA + nR;
    double t1 = 0.0, t2 = 0.0; // t at the previous two chain positions
#pragma unroll 2
    for (int k = N; k >= 2; k--) {
        // the term with the older value first: one FMA on the loop-carried critical path
        const double tx = (WX[k] - L2X[k] * t2) - L1X[k] * t1;
        WX[k] = tx;
        const double tu = (WU[k - 1] - L2U[k - 1] * t1) - L1U[k - 1] * tx;
        WU[k - 1] = tu;
        t2 = tx; t1 = tu;
    }
    {
        const double tx = (WX[1] - L2X[1] * t2) - L1X[1] * t1; // x_1
        WX[1] = tx;
        t2 = t1; t1 = tx;
    }
    WU[0] = (WU[0] - L2U[0] * t2) - L1U[0] * t1;
    WX[0] = WX[0] - L2X[0] * t1;
}

// backward substitution, starting from the corner solution already stored in W[u_0], W[x_0]
__device__ __forceinline__ void sweep_backward(const Smem& s, int a, int N)
{
    const int nA = 2 * N + 1, nR = N + 1;
    double* __restrict__ WX = s.W + a * nA; double* __restrict__ WU = s.W + a * nA + nR;
    const double* __restrict__ L1X = s.L1 + a * nA; const double* __restrict__ L1U = s.L1 + a * nA + nR;
    const double* __restrict__ L2X = s.L2 + a * nA; const double* __restrict__ L2U = s.L2 + a * nA + nR;
    const double* __restrict__ DIX = s.DI + a * nA; const double* __restrict__ DIU = s.DI + a * nA + nR;
    double a1 = WU[0], a2 = WX[0]; // solution at chain positions j+1, j+2
#pragma unroll 2
    for (int k = 1; k < N; k++) {
        const double xx = (WX[k] * DIX[k] - L2X[k - 1] * a2) - L1U[k - 1] * a1; // x_k
        WX[k] = xx;
        const double xu = (WU[k] * DIU[k] - L2U[k - 1] * a1) - L1X[k] * xx;     // u_k
        WU[k] = xu;
        a2 = xx; a1 = xu;
    }
    WX[N] = (WX[N] * DIX[N] - L2X[N - 1] * a2) - L1U[N - 1] * a1;
}

// ---- arbitrary 2x2 weights: the two axis chains couple at every stage, so the reduced KKT matrix is block-pentadiagonal with 2x2
// blocks in the chain order x_N, u_{N-1}, x_{N-1}, ..., x_1, u_0, x_0 (2N+1 block positions; the hull rows on u_0 are part of that
// position's diagonal block, no separate corner).  Block LDL' and the two substitutions run on ONE lane: the rare path.
// Position p: even p = 2m <-> x_{N-m}, odd p = 2m+1 <-> u_{N-1-m}; entry (a, a) of a block lives in L1 / L2 / DI at the split
// index of (a, p), entry (a, other axis) in L1C / L2C / DIC.
__device__ __forceinline__ int blk_index(int a, int p, int N) { return a * (2 * N + 1) + ((p & 1) ? (N + 1) + (N - 1 - (p >> 1)) : N - (p >> 1)); }

struct M2 { double a, b, c, d; }; // [[a, b], [c, d]]
__device__ __forceinline__ M2 m2_mul(const M2& x, const M2& y) { return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d}; }
__device__ __forceinline__ M2 m2_mul_t(const M2& x, const M2& y) { return {x.a * y.a + x.b * y.b, x.a * y.c + x.b * y.d, x.c * y.a + x.d * y.b, x.c * y.c + x.d * y.d}; } // x y'
__device__ __forceinline__ M2 m2_sub(const M2& x, const M2& y) { return {x.a - y.a, x.b - y.b, x.c - y.c, x.d - y.d}; }
__device__ __forceinline__ M2 m2_inv_sym(const M2& x) { const double det = x.a * x.d - x.b * x.c, id = 1.0 / det; return {x.d * id, -x.b * id, -x.c * id, x.a * id}; }

__device__ bool factor_block(const Smem& s, int N, const double* cQ, const double* cR, double cqx, double cr01, double cr10, double sigma,
                             double rho_e, const double* hull_diag, double kxy)
{
    const int nA = 2 * N + 1, nR = N + 1;
    M2 Dm1 = {1, 0, 0, 1}, Dm2 = {1, 0, 0, 1}, Dim1 = {1, 0, 0, 1}, Dim2 = {1, 0, 0, 1}, L1m1 = {0, 0, 0, 0};
    bool pd = true; // every pivot block positive definite <=> the KKT matrix has OSQP's expected inertia (n positive pivots)
    for (int p = 0; p <= 2 * N; p++) {
        const bool isu = p & 1;
        const int k = isu ? N - 1 - (p >> 1) : N - (p >> 1);
        M2 Kpp, E1 = {0, 0, 0, 0}, E2 = {0, 0, 0, 0};
        double dd[2];
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const double* CM = s.CM + a * nR; const double* CA = s.CA + a * nR; const double* CB = s.CB + a * nR;
            const double* DX = s.DD + a * nA; const double* DU = DX + nR;
            double e1 = 0.0, e2 = 0.0;
            if (!isu) {
                dd[a] = cQ[a] * DX[k] * DX[k] + sigma + rho_e * (CM[k] * CM[k] + (k < N ? CA[k + 1] * CA[k + 1] : 0.0));
                if (k < N) { e1 = rho_e * CA[k + 1] * CB[k + 1]; e2 = rho_e * CM[k + 1] * CA[k + 1]; } // with u_k (p-1), x_{k+1} (p-2)
            } else {
                const double sc = (k < N - 1) ? 2.0 : 1.0;
                dd[a] = cR[a] * sc * DU[k] * DU[k] + sigma + rho_e * CB[k + 1] * CB[k + 1] + (k == 0 ? hull_diag[a] : 0.0);
                e1 = rho_e * CM[k + 1] * CB[k + 1];                     // with x_{k+1} (p-1)
                if (k < N - 1) e2 = -cR[a] * DU[k] * DU[k + 1];         // with u_{k+1} (p-2)
            }
            if (a == 0) { E1.a = e1; E2.a = e2; } else { E1.d = e1; E2.d = e2; }
        }
        double off;
        if (!isu) off = cqx * s.DD[k] * s.DD[nA + k];
        else {
            const double sc = (k < N - 1) ? 2.0 : 1.0;
            off = cr01 * sc * s.DD[nR + k] * s.DD[nA + nR + k] + (k == 0 ? kxy : 0.0);
            if (k < N - 1) { // rows (u_k^x, u_k^y) against columns (u_{k+1}^x, u_{k+1}^y): P(u_k^x, u_{k+1}^y) = -R(0,1), P(u_k^y, u_{k+1}^x) = -R(1,0)
                E2.b = -cr01 * s.DD[nR + k] * s.DD[nA + nR + k + 1];
                E2.c = -cr10 * s.DD[nA + nR + k] * s.DD[nR + k + 1];
            }
        }
        Kpp = {dd[0], off, off, dd[1]};
        M2 L2 = {0, 0, 0, 0}, L1 = {0, 0, 0, 0};
        if (p >= 2) L2 = m2_mul(E2, Dim2);
        if (p >= 1) {
            M2 t = E1;
            if (p >= 2) t = m2_sub(t, m2_mul_t(m2_mul(L2, Dm2), L1m1));
            L1 = m2_mul(t, Dim1);
        }
        M2 Dp = Kpp;
        if (p >= 1) Dp = m2_sub(Dp, m2_mul_t(m2_mul(L1, Dm1), L1));
        if (p >= 2) Dp = m2_sub(Dp, m2_mul_t(m2_mul(L2, Dm2), L2));
        Dp.b = Dp.c = 0.5 * (Dp.b + Dp.c);
        if (!(Dp.a > 0.0) || !(Dp.a * Dp.d - Dp.b * Dp.c > 0.0)) pd = false;
        const M2 Di = m2_inv_sym(Dp);
        const int i0 = blk_index(0, p, N), i1 = blk_index(1, p, N);
        s.L1[i0] = L1.a; s.L1C[i0] = L1.b; s.L1C[i1] = L1.c; s.L1[i1] = L1.d;
        s.L2[i0] = L2.a; s.L2C[i0] = L2.b; s.L2C[i1] = L2.c; s.L2[i1] = L2.d;
        s.DI[i0] = Di.a; s.DIC[i0] = Di.b; s.DIC[i1] = Di.c; s.DI[i1] = Di.d;
        Dm2 = Dm1; Dim2 = Dim1; Dm1 = Dp; Dim1 = Di; L1m1 = L1;
    }
    return pd;
}

// (P + sigma I + A' rho A) x~ = W in place, block chain
__device__ void solve_block(const Smem& s, int N)
{
    double t1x = 0, t1y = 0, t2x = 0, t2y = 0;
    for (int p = 0; p <= 2 * N; p++) { // L t = w
        const int i0 = blk_index(0, p, N), i1 = blk_index(1, p, N);
        const double tx = s.W[i0] - (s.L1[i0] * t1x + s.L1C[i0] * t1y) - (s.L2[i0] * t2x + s.L2C[i0] * t2y);
        const double ty = s.W[i1] - (s.L1C[i1] * t1x + s.L1[i1] * t1y) - (s.L2C[i1] * t2x + s.L2[i1] * t2y);
        s.W[i0] = tx; s.W[i1] = ty;
        t2x = t1x; t2y = t1y; t1x = tx; t1y = ty;
    }
    double x1x = 0, x1y = 0, x2x = 0, x2y = 0; // solution at p+1, p+2
    for (int p = 2 * N; p >= 0; p--) { // x_p = D_p^-1 t_p - L1_{p+1}' x_{p+1} - L2_{p+2}' x_{p+2}
        const int i0 = blk_index(0, p, N), i1 = blk_index(1, p, N);
        double xx = s.DI[i0] * s.W[i0] + s.DIC[i0] * s.W[i1];
        double xy = s.DIC[i1] * s.W[i0] + s.DI[i1] * s.W[i1];
        if (p + 1 <= 2 * N) {
            const int j0 = blk_index(0, p + 1, N), j1 = blk_index(1, p + 1, N);
            xx -= s.L1[j0] * x1x + s.L1C[j1] * x1y;
            xy -= s.L1C[j0] * x1x + s.L1[j1] * x1y;
        }
        if (p + 2 <= 2 * N) {
            const int j0 = blk_index(0, p + 2, N), j1 = blk_index(1, p + 2, N);
            xx -= s.L2[j0] * x2x + s.L2C[j1] * x2y;
            xy -= s.L2C[j0] * x2x + s.L2[j1] * x2y;
        }
        s.W[i0] = xx; s.W[i1] = xy;
        x2x = x1x; x2y = x1y; x1x = xx; x1y = xy;
    }
}

template <int G>
__global__ void __launch_bounds__(128) mpc_admm_kernel(MpcDev d, MpcTickIO io, int smem_stride)
{
    extern __shared__ double smem_all[];
    constexpr int IPW = 32 / G; // instances per warp
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int grp = lane / G;   // instance slot in the warp
    const int gl = lane % G;    // lane in group
    const int slot = warp_in_block * IPW + grp;
    const long long inst_ll = ((long long)blockIdx.x * warps_per_block + warp_in_block) * IPW + grp;
    const size_t B = (size_t)d.B;
    const int N = d.N, nA = 2 * N + 1, nR = N + 1;
    const bool valid = inst_ll < (long long)d.B;
    const int i = valid ? (int)inst_ll : d.B - 1;
    const unsigned FULL = 0xffffffffu;
    // lanes of this instance's group: group collectives run under group-uniform control flow
    const unsigned gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (grp * G));
    Smem s = carve(smem_all + (size_t)slot * smem_stride, N, d.coupled, d.par);

    const wc_osqp_settings& st = d.st;
    const double sigma = st.sigma, alpha = st.alpha, infty = st.infty;

    const int kflags = d.tick_flags[i];
    const bool skip = !valid || (kflags & (WC_MPC_FLAG_NO_CONTACT | WC_MPC_FLAG_NO_SOLVER)) != 0;
    const int nh = skip ? 0 : d.nhull[i];
    const double Q[2] = {d.Qd[i], d.Qd[B + i]};
    const double R[2] = {d.Rd[i], d.Rd[B + i]};
    // a new MPCSolver (support phase changed) has no OSQP workspace yet: osqp_setup runs below
    const bool setup = !skip && d.inited[i] == 0;
    double c = setup ? 1.0 : d.cscale[i];
    double rho = setup ? fmin(fmax(st.rho, kRhoMin), kRhoMax) : d.rho[i];
    const double up[2] = {d.uprev[i], d.uprev[B + i]};
    const int head = mpc_ring_head(d);
    const bool cpl = d.coupled != 0;
    const double qx = d.Qoff[0], r01 = d.Roff[0], r10 = d.Roff[1]; // the Hessian's cross-axis entries (0 unless cpl)

    // ------------------------------------------------------------------ stage in
    for (int j = gl; j < 2 * nA; j += G) {
        s.X[j] = setup ? 0.0 : d.x[(size_t)j * B + i];
        s.DD[j] = setup ? 1.0 : d.D[(size_t)j * B + i];
        s.QB[j] = 0.0;
    }
    for (int j = gl; j < 2 * nR; j += G) {
        s.YE[j] = setup ? 0.0 : d.yeq[(size_t)j * B + i];
        s.EE[j] = setup ? 1.0 : d.Eeq[(size_t)j * B + i];
        s.DYE[j] = 0.0;
    }
    for (int h = gl; h < kMaxHull; h += G) {
        s.YH[h] = setup ? 0.0 : d.yh[(size_t)h * B + i];
        s.ZH[h] = setup ? 0.0 : d.zh[(size_t)h * B + i];
        s.EH[h] = setup ? 1.0 : d.Eh[(size_t)h * B + i];
        s.DYH[h] = 0.0;
        // raw hull rows parked in HX/HY until the scaled coefficients replace them
        s.HX[h] = (h < nh) ? d.hullA[(size_t)(2 * h) * B + i] : 0.0;
        s.HY[h] = (h < nh) ? d.hullA[(size_t)(2 * h + 1) * B + i] : 0.0;
    }
    __syncwarp();
    // unscaled gradient (MPCSolver::setGradient): state block from the ring, input block G u_prev
    for (int k = gl; k < nR; k += G) {
        const int slot_k = (head + k) % nR;
        s.QB[k] = d.grad[(size_t)(slot_k)*B + i];
        s.QB[nA + k] = d.grad[(size_t)(nR + slot_k) * B + i];
    }
    if (gl == 0) { // G = -Theta' R~ e1 -> only the u_0 block (MPCSolver.cpp:244-245)
        s.QB[nR] = cpl ? -(R[0] * up[0] + d.Roff[0] * up[1]) : -(R[0] * up[0]);
        s.QB[nA + nR] = cpl ? -(d.Roff[1] * up[0] + R[1] * up[1]) : -(R[1] * up[1]);
    }
    __syncwarp();
    if (setup) {
        // ---- osqp_setup -> scale_data: Ruiz equilibration of [[P, A'],[A, 0]] with cost normalisation
        // (OSQP scaling.c).  P and A are never formed: their scaled entries follow from D, E, c in
        // closed form.  Group-uniform branch: only group-scope collectives / barriers inside.
        const double ad = fabs(d.a_dyn), bd = fabs(d.b_dyn);
        for (int it = 0; it < st.scaling; it++) {
            for (int k = gl; k < nR; k += G) {
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const double* DX = s.DD + a * nA; const double* DU = DX + nR; const double* Ea = s.EE + a * nR;
                    const double dx = DX[k];
                    const double* DXo = s.DD + (1 - a) * nA; const double* DUo = DXo + nR; // the other axis
                    const double rnx = (a == 0) ? r01 : r10, rpv = (a == 0) ? r10 : r01;  // P(u_k^a, u_{k+1}^o), P(u_k^a, u_{k-1}^o)
                    double nrm = fabs(c * Q[a] * dx * dx);
                    if (cpl) nrm = amax(nrm, c * qx * dx * DXo[k]);
                    nrm = fmax(nrm, Ea[k] * dx);
                    if (k < N) nrm = fmax(nrm, ad * Ea[k + 1] * dx);
                    s.W[a * nA + k] = 1.0 / sqrt(limit_scaling(nrm));
                    if (k < N) {
                        const double du = DU[k];
                        const double sc = (k < N - 1) ? 2.0 : 1.0;
                        double nu = fabs(c * sc * R[a] * du * du);
                        if (k >= 1) nu = amax(nu, c * R[a] * du * DU[k - 1]);
                        if (k < N - 1) nu = amax(nu, c * R[a] * du * DU[k + 1]);
                        if (cpl) {
                            nu = amax(nu, c * sc * r01 * du * DUo[k]);
                            if (k >= 1) nu = amax(nu, c * rpv * du * DUo[k - 1]);
                            if (k < N - 1) nu = amax(nu, c * rnx * du * DUo[k + 1]);
                        }
                        nu = fmax(nu, bd * Ea[k + 1] * du);
                        if (k == 0) {
                            const double* Hc = (a == 0) ? s.HX : s.HY;
                            for (int h = 0; h < nh; h++) nu = fmax(nu, fabs(s.EH[h] * Hc[h]) * du);
                        }
                        s.W[a * nA + nR + k] = 1.0 / sqrt(limit_scaling(nu));
                    }
                    double nr = Ea[k] * dx;
                    if (k >= 1) {
                        nr = fmax(nr, ad * Ea[k] * DX[k - 1]);
                        nr = fmax(nr, bd * Ea[k] * DU[k - 1]);
                    }
                    s.DYE[a * nR + k] = 1.0 / sqrt(limit_scaling(nr));
                }
            }
            for (int h = gl; h < nh; h += G) {
                const double nrm = fmax(fabs(s.EH[h] * s.HX[h]) * s.DD[nR], fabs(s.EH[h] * s.HY[h]) * s.DD[nA + nR]);
                s.DYH[h] = 1.0 / sqrt(limit_scaling(nrm));
            }
            __syncwarp(gmask);
            for (int j = gl; j < 2 * nA; j += G) s.DD[j] *= s.W[j];
            for (int j = gl; j < 2 * nR; j += G) s.EE[j] *= s.DYE[j];
            for (int h = gl; h < nh; h += G) s.EH[h] *= s.DYH[h];
            __syncwarp(gmask);
            // cost normalisation: c <- c / max(mean_j ||P_j||_inf, ||q||_inf)
            double sum = 0.0, qn = 0.0;
            for (int k = gl; k < nR; k += G) {
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const double* DX = s.DD + a * nA; const double* DU = DX + nR;
                    const double dx = DX[k];
                    const double* DXo = s.DD + (1 - a) * nA; const double* DUo = DXo + nR;
                    const double rnx = (a == 0) ? r01 : r10, rpv = (a == 0) ? r10 : r01;
                    sum += cpl ? fmax(fabs(c * Q[a] * dx * dx), fabs(c * qx * dx * DXo[k])) : fabs(c * Q[a] * dx * dx);
                    qn = amax(qn, c * dx * s.QB[a * nA + k]);
                    if (k < N) {
                        const double du = DU[k];
                        const double sc = (k < N - 1) ? 2.0 : 1.0;
                        double nu = fabs(c * sc * R[a] * du * du);
                        if (k >= 1) nu = amax(nu, c * R[a] * du * DU[k - 1]);
                        if (k < N - 1) nu = amax(nu, c * R[a] * du * DU[k + 1]);
                        if (cpl) {
                            nu = amax(nu, c * sc * r01 * du * DUo[k]);
                            if (k >= 1) nu = amax(nu, c * rpv * du * DUo[k - 1]);
                            if (k < N - 1) nu = amax(nu, c * rnx * du * DUo[k + 1]);
                        }
                        sum += nu;
                        qn = amax(qn, c * du * s.QB[a * nA + nR + k]);
                    }
                }
            }
            sum = group_sum<G>(sum, gmask);
            qn = group_max<G>(qn, gmask);
            double c_temp = sum / (double)(2 * nA);
            c_temp = fmax(c_temp, limit_scaling(qn));
            c_temp = limit_scaling(c_temp);
            c *= 1.0 / c_temp;
        }
        for (int j = gl; j < 2 * nR; j += G) s.DYE[j] = 0.0;
        for (int h = gl; h < kMaxHull; h += G) s.DYH[h] = 0.0;
        // publish the scaling of the new solver
        for (int j = gl; j < 2 * nA; j += G) d.D[(size_t)j * B + i] = s.DD[j];
        for (int j = gl; j < 2 * nR; j += G) d.Eeq[(size_t)j * B + i] = s.EE[j];
        for (int h = gl; h < kMaxHull; h += G) d.Eh[(size_t)h * B + i] = s.EH[h];
        if (gl == 0) d.cscale[i] = c; // (inited is set once the factorisation below has succeeded)
    }
    __syncwarp();
    const double cinv = 1.0 / c;
    // scaled problem data: q (osqp_update_lin_cost), A coefficients, bounds (osqp_update_bounds)
    for (int j = gl; j < 2 * nA; j += G) s.QB[j] = (s.QB[j] * s.DD[j]) * c;
    for (int k = gl; k < nR; k += G) {
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const double* DX = s.DD + a * nA; const double* DU = DX + nR;
            const double e = s.EE[a * nR + k];
            s.CM[a * nR + k] = -e * DX[k];
            s.CA[a * nR + k] = (k >= 1) ? d.a_dyn * e * DX[k - 1] : 0.0;
            s.CB[a * nR + k] = (k >= 1) ? d.b_dyn * e * DU[k - 1] : 0.0;
        }
    }
    for (int h = gl; h < kMaxHull; h += G) {
        const bool on = h < nh;
        const double e = s.EH[h];
        s.HX[h] = on ? e * s.HX[h] * s.DD[nR] : 0.0;
        s.HY[h] = on ? e * s.HY[h] * s.DD[nA + nR] : 0.0;
        s.UH[h] = on ? e * d.hullb[(size_t)h * B + i] : 0.0;
        s.LH[h] = on ? e * (-infty) : 0.0;
    }
    // equality row 0: l = u = -dcm (MPCSolver.cpp:143-146), scaled by E
    const double lb0[2] = {-io.dcm[i] * s.EE[0], -io.dcm[B + i] * s.EE[nR]};
    double z0prev[2] = {setup ? 0.0 : d.z0[i], setup ? 0.0 : d.z0[B + i]};
    __syncwarp();


    // ------------------------------------------------------------------ segment-parallel chain solve (d.par; long horizons)
    // Chain positions p = 0 .. 2N per axis (even p = 2m <-> x_{N-m}, odd p = 2m+1 <-> u_{N-1-m}); K is pentadiagonal in p.
    // Lane (axis ap, segment sg) owns the interior [p_lo, p_hi] of its segment; separator sg = positions p_hi+1, p_hi+2; the last
    // separator is the corner (u_0, x_0).  Interiors are eliminated in parallel (local pentadiagonal LDL'), what remains is a
    // block-tridiagonal system over the separators of an axis (2x2 blocks, eliminated serially by the axis' first lane) that
    // ends in the same (u_0, x_0) Schur complement the serial path hands to the dense 4x4 corner.
    const bool par = d.par != 0 && !cpl;
    constexpr int PS = G / 2;                      // segments per axis
    const int ap = gl / PS, sg = gl % PS;
    int p_lo = 0, p_hi = -1;
    {
        const int nInt = nA - 2 * PS, base = nInt / PS, rem = nInt % PS;
        p_lo = sg * (base + 2) + (sg < rem ? sg : rem);
        p_hi = p_lo + base + (sg < rem ? 1 : 0) - 1;
    }
    const double cQa = c * Q[ap < 2 ? ap : 0], cRa = c * R[ap < 2 ? ap : 0];
    auto kdiag = [&](int p, double rho_e, double hd) -> double {
        const double* CM = s.CM + ap * nR; const double* CA = s.CA + ap * nR; const double* CB = s.CB + ap * nR;
        const double* DX = s.DD + ap * nA; const double* DU = DX + nR;
        if (!(p & 1)) { const int k = N - (p >> 1); return cQa * DX[k] * DX[k] + sigma + rho_e * (CM[k] * CM[k] + (k < N ? CA[k + 1] * CA[k + 1] : 0.0)); }
        const int k = N - 1 - (p >> 1);
        const double sc = (k < N - 1) ? 2.0 : 1.0;
        return cRa * sc * DU[k] * DU[k] + sigma + rho_e * CB[k + 1] * CB[k + 1] + (k == 0 ? hd : 0.0);
    };
    auto kof1 = [&](int p, double rho_e) -> double { // K[p, p+1], p <= 2N-1
        const double* CM = s.CM + ap * nR; const double* CA = s.CA + ap * nR; const double* CB = s.CB + ap * nR;
        if (!(p & 1)) { const int k = N - (p >> 1); return rho_e * CM[k] * CB[k]; }
        const int k = N - 1 - (p >> 1); return rho_e * CA[k + 1] * CB[k + 1];
    };
    auto kof2 = [&](int p, double rho_e) -> double { // K[p, p+2], p <= 2N-2
        const double* CM = s.CM + ap * nR; const double* CA = s.CA + ap * nR;
        const double* DU = s.DD + ap * nA + nR;
        if (!(p & 1)) { const int k = N - (p >> 1); return rho_e * CM[k] * CA[k]; }
        const int k = N - 1 - (p >> 1); return -cRa * DU[k] * DU[k - 1];
    };
    // local solve K_II v = src over the lane's interior (vectors indexed by the split index): forward substitution src -> tmp,
    // backward substitution on tmp (written back only if STORE); bv = the solution at p_lo, p_lo + 1, p_hi - 1, p_hi.  The split
    // index of position p is c0 - (p >> 1) for even p (a state) and c1 - (p >> 1) for odd p (an input): walked in pairs.
    const int c0 = ap * nA + N, c1 = ap * nA + nR + N - 1;
    auto split_of = [&](int p) { return ((p & 1) ? c1 : c0) - (p >> 1); };
    auto local_solve = [&](const double* src, double* tmp, const bool STORE, double (&bv)[4]) {
        double t1 = 0.0, t2 = 0.0;
        auto fstep = [&](int j) {
            const double t = (src[j] - s.L2[j] * t2) - s.L1[j] * t1;
            tmp[j] = t; t2 = t1; t1 = t;
        };
        int p = p_lo;
        if (p & 1) { fstep(c1 - (p >> 1)); p++; }
        for (; p + 1 <= p_hi; p += 2) { const int hh = p >> 1; fstep(c0 - hh); fstep(c1 - hh); }
        if (p <= p_hi) fstep(c0 - (p >> 1));
        double x1 = 0.0, x2 = 0.0, l1n = 0.0, l2n = 0.0, l2nn = 0.0; // multipliers of the rows above: L1[p+1], L2[p+1], L2[p+2]
        auto bstep = [&](int j) {
            const double x = (tmp[j] * s.DI[j] - l2nn * x2) - l1n * x1;
            if (STORE) tmp[j] = x;
            x2 = x1; x1 = x;
            l2nn = l2n; l1n = s.L1[j]; l2n = s.L2[j];
        };
        bstep(split_of(p_hi)); bv[3] = x1;      // (segments have at least two interior unknowns)
        bstep(split_of(p_hi - 1)); bv[2] = x1;
        p = p_hi - 2;
        if (p >= p_lo && !(p & 1)) { bstep(c0 - (p >> 1)); p--; }
        for (; p - 1 >= p_lo; p -= 2) { const int hh = p >> 1; bstep(c1 - hh); bstep(c0 - hh); } // p odd: u_k, then the state with the same shift
        if (p >= p_lo) bstep(c1 - (p >> 1));
        bv[0] = x1; bv[1] = x2; // the last two values produced: positions p_lo and p_lo + 1
    };
    auto factor_par = [&](double rho_now) {
        const double rho_e = kRhoEqOverIneq * rho_now;
        double hd = 0.0;
        for (int h = 0; h < nh; h++) { const double v = (ap == 0) ? s.HX[h] : s.HY[h]; hd += rho_now * v * v; }
        // ---- local LDL' of the interior: L1[p] = L(p, p-1), L2[p] = L(p, p-2), DI[p] = 1 / D_p
        {
            double Dm1 = 1.0, Dm2 = 1.0, l1m1 = 0.0;
            for (int p = p_lo; p <= p_hi; p++) {
                const int j = split_of(p);
                const double e1 = p - 1 >= p_lo ? kof1(p - 1, rho_e) : 0.0, e2 = p - 2 >= p_lo ? kof2(p - 2, rho_e) : 0.0;
                const double l2 = e2 / Dm2;
                const double l1 = (e1 - l2 * Dm2 * l1m1) / Dm1;
                const double Dp = kdiag(p, rho_e, hd) - l1 * l1 * Dm1 - l2 * l2 * Dm2;
                s.L1[j] = l1; s.L2[j] = l2; s.DI[j] = 1.0 / Dp;
                Dm2 = Dm1; Dm1 = Dp; l1m1 = l1;
            }
        }
        // ---- G = K_SI K_II^-1 K_IS over the four neighbouring separator unknowns [q0, q1 | r0, r1] (q: p_lo-2, p_lo-1; r: p_hi+1, p_hi+2)
        // couplings: (q0, p_lo): k2(q0); (q1, p_lo): k1(q1); (q1, p_lo+1): k2(q1); (p_hi, r0): k1(p_hi); (p_hi-1, r0): k2(p_hi-1); (p_hi, r1): k2(p_hi)
        const bool hasL = sg > 0;
        const double cq0 = hasL ? kof2(p_lo - 2, rho_e) : 0.0, cq1a = hasL ? kof1(p_lo - 1, rho_e) : 0.0, cq1b = hasL ? kof2(p_lo - 1, rho_e) : 0.0;
        const double cr0a = kof1(p_hi, rho_e), cr0b = kof2(p_hi - 1, rho_e), cr1 = kof2(p_hi, rho_e);
        double Gm[4][4];
        double* Zv = s.ZP;
#pragma unroll 1
        for (int col = hasL ? 0 : 2; col < 4; col++) {
            for (int p = p_lo; p <= p_hi; p++) Zv[split_of(p)] = 0.0;
            if (col == 0) Zv[split_of(p_lo)] = cq0;
            else if (col == 1) { Zv[split_of(p_lo)] = cq1a; Zv[split_of(p_lo + 1)] += cq1b; }
            else if (col == 2) { Zv[split_of(p_hi)] = cr0a; Zv[split_of(p_hi - 1)] += cr0b; }
            else Zv[split_of(p_hi)] = cr1;
            double bv[4];
            local_solve(Zv, Zv, false, bv);
            const double v0 = bv[0], v1 = bv[1], vm = bv[2], ve = bv[3];
            Gm[0][col] = cq0 * v0; Gm[1][col] = cq1a * v0 + cq1b * v1; Gm[2][col] = cr0a * ve + cr0b * vm; Gm[3][col] = cr1 * ve;
        }
        // publish the lane's contribution (W is free between solves): [RR (3), RL (4), LL (3)] at W[(ap * PS + sg) * 10 ...]
        double* Gs = s.W + (ap * PS + sg) * 10;
        Gs[0] = Gm[2][2]; Gs[1] = Gm[2][3]; Gs[2] = Gm[3][3];
        Gs[3] = hasL ? Gm[2][0] : 0.0; Gs[4] = hasL ? Gm[2][1] : 0.0; Gs[5] = hasL ? Gm[3][0] : 0.0; Gs[6] = hasL ? Gm[3][1] : 0.0;
        Gs[7] = hasL ? Gm[0][0] : 0.0; Gs[8] = hasL ? Gm[0][1] : 0.0; Gs[9] = hasL ? Gm[1][1] : 0.0;
        __syncwarp(gmask);
        // ---- reduced block-tridiagonal system over the separators of the axis: serial block LDL' on the axis' first lane
        if (sg == 0) {
            const int nInt = nA - 2 * PS, base = nInt / PS, rem = nInt % PS;
            M2 Ap = {1, 0, 0, 1}, ApInv = {1, 0, 0, 1};
            for (int t = 0; t < PS; t++) {
                const int hi = t * (base + 2) + (t < rem ? t : rem) + base + (t < rem ? 1 : 0) - 1; // p_hi of segment t
                const int r0 = hi + 1, r1 = hi + 2;
                const double* Gt = s.W + (ap * PS + t) * 10;
                M2 A = {kdiag(r0, rho_e, hd) - Gt[0], kof1(r0, rho_e) - Gt[1], kof1(r0, rho_e) - Gt[1], kdiag(r1, rho_e, hd) - Gt[2]};
                if (t + 1 < PS) { const double* Gn = s.W + (ap * PS + t + 1) * 10; A.a -= Gn[7]; A.b -= Gn[8]; A.c -= Gn[8]; A.d -= Gn[9]; }
                M2 LB = {0, 0, 0, 0};
                if (t >= 1) {
                    const M2 Bt = {-Gt[3], -Gt[4], -Gt[5], -Gt[6]}; // rows (r0, r1) of separator t, columns (q0, q1) = separator t-1
                    LB = m2_mul(Bt, ApInv);
                    A = m2_sub(A, m2_mul_t(m2_mul(LB, Ap), LB));
                    A.b = A.c = 0.5 * (A.b + A.c);
                }
                Ap = A;
                double* Rd = s.RED + (ap * 16 + t) * 8;
                Rd[0] = LB.a; Rd[1] = LB.b; Rd[2] = LB.c; Rd[3] = LB.d;
                if (t + 1 < PS) {
                    ApInv = m2_inv_sym(Ap);
                    Rd[4] = ApInv.a; Rd[5] = ApInv.b; Rd[6] = ApInv.c; Rd[7] = ApInv.d;
                } else { // the corner (u_0, x_0): Schur complement for the dense 4x4
                    s.RS[3 * ap] = Ap.a; s.RS[3 * ap + 1] = Ap.b; s.RS[3 * ap + 2] = Ap.d;
                }
            }
        }
        __syncwarp(gmask);
        if (gl == 0) {
            double kxy = 0.0;
            for (int h = 0; h < nh; h++) kxy += rho_now * s.HX[h] * s.HY[h];
            factor_corner(s.RS, s.RS + 3, kxy, s.CF);
        }
        __syncwarp(gmask);
    };
    // one reduced-KKT solve in place on W (group-uniform; `act` freezes finished instances)
    auto solve_par = [&](bool act, double rho_now) {
        const double rho_e = kRhoEqOverIneq * rho_now;
        const bool hasL = sg > 0;
        const double cq0 = hasL ? kof2(p_lo - 2, rho_e) : 0.0, cq1a = hasL ? kof1(p_lo - 1, rho_e) : 0.0, cq1b = hasL ? kof2(p_lo - 1, rho_e) : 0.0;
        const double cr0a = kof1(p_hi, rho_e), cr0b = kof2(p_hi - 1, rho_e), cr1 = kof2(p_hi, rho_e);
        if (act) { // pass 1: boundary values of K_II^-1 r_I -> the separators' right-hand sides
            double bv[4];
            local_solve(s.W, s.ZP, false, bv);
            const double v0 = bv[0], v1 = bv[1], vm = bv[2], ve = bv[3];
            double* rt = s.RT + (ap * 16 + sg) * 4;
            rt[0] = cq0 * v0; rt[1] = cq1a * v0 + cq1b * v1; rt[2] = cr0a * ve + cr0b * vm; rt[3] = cr1 * ve;
        }
        __syncwarp(gmask);
        const int nInt = nA - 2 * PS, base = nInt / PS, rem = nInt % PS;
        auto sep_r0 = [&](int t) { return t * (base + 2) + (t < rem ? t : rem) + base + (t < rem ? 1 : 0); };
        if (act && sg == 0) { // reduced forward substitution along the separators of the axis
            double px = 0.0, py = 0.0;
            for (int t = 0; t < PS; t++) {
                const int j0 = split_of(sep_r0(t)), j1 = split_of(sep_r0(t) + 1);
                const double* rt = s.RT + (ap * 16 + t) * 4;
                double r0 = s.W[j0] - rt[2], r1 = s.W[j1] - rt[3];
                if (t + 1 < PS) { const double* rn = s.RT + (ap * 16 + t + 1) * 4; r0 -= rn[0]; r1 -= rn[1]; }
                const double* Rd = s.RED + (ap * 16 + t) * 8;
                r0 -= Rd[0] * px + Rd[1] * py; r1 -= Rd[2] * px + Rd[3] * py;
                s.W[j0] = r0; s.W[j1] = r1; px = r0; py = r1;
            }
        }
        __syncwarp(gmask);
        if (act && gl == 0) {
            double r[4] = {s.W[nR], s.W[nA + nR], s.W[0], s.W[nA]};
            solve_corner(s.CF, r);
            s.W[nR] = r[0]; s.W[nA + nR] = r[1]; s.W[0] = r[2]; s.W[nA] = r[3];
        }
        __syncwarp(gmask);
        if (act && sg == 0) { // reduced backward substitution
            double nx = s.W[split_of(sep_r0(PS - 1))], ny = s.W[split_of(sep_r0(PS - 1) + 1)];
            for (int t = PS - 2; t >= 0; t--) {
                const int j0 = split_of(sep_r0(t)), j1 = split_of(sep_r0(t) + 1);
                const double* Rd = s.RED + (ap * 16 + t) * 8; const double* Rn = s.RED + (ap * 16 + t + 1) * 8;
                const double r0 = s.W[j0], r1 = s.W[j1];
                const double x0 = (Rd[4] * r0 + Rd[5] * r1) - (Rn[0] * nx + Rn[2] * ny); // A^-1 r - LB_{t+1}' x_{t+1}
                const double x1 = (Rd[6] * r0 + Rd[7] * r1) - (Rn[1] * nx + Rn[3] * ny);
                s.W[j0] = x0; s.W[j1] = x1; nx = x0; ny = x1;
            }
        }
        __syncwarp(gmask);
        if (act) { // pass 2: x_I = K_II^-1 (r_I - K_IS x_S)
            if (hasL) {
                const double xq0 = s.W[split_of(p_lo - 2)], xq1 = s.W[split_of(p_lo - 1)];
                s.W[split_of(p_lo)] -= cq0 * xq0 + cq1a * xq1;
                s.W[split_of(p_lo + 1)] -= cq1b * xq1;
            }
            const double xr0 = s.W[split_of(p_hi + 1)], xr1 = s.W[split_of(p_hi + 2)];
            s.W[split_of(p_hi)] -= cr0a * xr0 + cr1 * xr1;
            s.W[split_of(p_hi - 1)] -= cr0b * xr0;
            double bv[4];
            local_solve(s.W, s.W, true, bv);
        }
        __syncwarp(gmask);
    };

    // ------------------------------------------------------------------ factorisation
    // returns false when the reduced KKT matrix is not positive definite (arbitrary weight triplets can make the problem
    // non-convex: osqp_setup then fails on the inertia of the KKT factorisation and there is no solver)
    auto factor = [&](double rho_now) -> bool {
        const double rho_e = kRhoEqOverIneq * rho_now;
        if (cpl) {
            int ok = 1;
            if (gl == 0) {
                double hd[2] = {0.0, 0.0}, kxy = 0.0;
                for (int h = 0; h < nh; h++) { hd[0] += rho_now * s.HX[h] * s.HX[h]; hd[1] += rho_now * s.HY[h] * s.HY[h]; kxy += rho_now * s.HX[h] * s.HY[h]; }
                const double cQ2[2] = {c * Q[0], c * Q[1]}, cR2[2] = {c * R[0], c * R[1]};
                ok = factor_block(s, N, cQ2, cR2, c * qx, c * r01, c * r10, sigma, rho_e, hd, kxy) ? 1 : 0;
            }
            __syncwarp();
            return __shfl_sync(gmask, ok, grp * G) != 0;
        }
        if (par) { factor_par(rho_now); __syncwarp(); return true; }
        if (gl < 2) {
            double hd = 0.0;
            for (int h = 0; h < nh; h++) { const double v = (gl == 0) ? s.HX[h] : s.HY[h]; hd += rho_now * v * v; }
            factor_axis(s, gl, N, c * Q[gl], c * R[gl], sigma, rho_e, hd, s.RS + 3 * gl);
        }
        __syncwarp();
        if (gl == 0) {
            double kxy = 0.0;
            for (int h = 0; h < nh; h++) kxy += rho_now * s.HX[h] * s.HY[h];
            factor_corner(s.RS, s.RS + 3, kxy, s.CF);
        }
        __syncwarp();
        return true;
    };
    // (called from ONE site, the top of the ADMM loop, so that it is inlined: a second call site makes it a real function, and
    // every local it captures by reference -- the shared-memory pointers, rho, c, the segment bounds -- then lives in local
    // memory for the whole kernel)
    bool setup_failed = false;

    // ------------------------------------------------------------------ ADMM (osqp_solve)
    int status = WC_OSQP_UNSOLVED;
    int iter_done = 0, rho_updates = 0;
    bool done = skip;
    bool factor_pending = true, first_factor = true;
    double pri_res = 0.0, dua_res = 0.0;
    // norms kept from the last update_info (needed by check_termination and compute_rho_estimate)
    double nz_u = 0, nAx_u = 0, nq_u = 0, nAty_u = 0, nPx_u = 0;
    double rp_s = 0, rd_s = 0, nz_s = 0, nAx_s = 0, nq_s = 0, nAty_s = 0, nPx_s = 0;
    int prim_inf = 0, dual_inf = 0;

    // (A' y)_j and (P x)_j for the variables of stage k, axis a; V = vector playing the role of x, Y of y
    // returns through references: x-variable (px_x, aty_x), u-variable (px_u, aty_u; only if k < N)
    auto stage_products = [&](int a, int k, const double* V, const double* Yeq, const double* Yh,
                              double& px_x, double& aty_x, double& px_u, double& aty_u) {
        const double* DX = s.DD + a * nA; const double* DU = DX + nR;
        const double* VX = V + a * nA; const double* VU = VX + nR;
        const double* Ya = Yeq + a * nR;
        const double* CMa = s.CM + a * nR; const double* CAa = s.CA + a * nR; const double* CBa = s.CB + a * nR;
        const double* DXo = s.DD + (1 - a) * nA; const double* DUo = DXo + nR;
        const double* VXo = V + (1 - a) * nA; const double* VUo = VXo + nR;
        px_x = (c * Q[a]) * DX[k] * DX[k] * VX[k];
        if (cpl) px_x += (c * qx) * DX[k] * (DXo[k] * VXo[k]);
        aty_x = CMa[k] * Ya[k] + (k < N ? CAa[k + 1] * Ya[k + 1] : 0.0);
        px_u = 0.0; aty_u = 0.0;
        if (k < N) {
            const double sc = (k < N - 1) ? 2.0 : 1.0;
            double t = sc * DU[k] * VU[k];
            if (k >= 1) t -= DU[k - 1] * VU[k - 1];
            if (k < N - 1) t -= DU[k + 1] * VU[k + 1];
            px_u = (c * R[a]) * DU[k] * t;
            if (cpl) {
                const double rnx = (a == 0) ? r01 : r10, rpv = (a == 0) ? r10 : r01;
                double tc = sc * r01 * (DUo[k] * VUo[k]);
                if (k >= 1) tc -= rpv * (DUo[k - 1] * VUo[k - 1]);
                if (k < N - 1) tc -= rnx * (DUo[k + 1] * VUo[k + 1]);
                px_u += c * DU[k] * tc;
            }
            aty_u = CBa[k + 1] * Ya[k + 1];
            if (k == 0) {
                const double* Hc = (a == 0) ? s.HX : s.HY;
                for (int h = 0; h < nh; h++) aty_u += Hc[h] * Yh[h];
            }
        }
    };

    // evaluates residual information for the CURRENT (x, z, y) (update_info + the norms of check_termination)
    auto update_info = [&]() {
        double m_rp_u = 0, m_rp_s = 0, m_z_u = 0, m_z_s = 0, m_ax_u = 0, m_ax_s = 0;
        double m_rd_u = 0, m_rd_s = 0, m_q_u = 0, m_q_s = 0, m_aty_u = 0, m_aty_s = 0, m_px_u = 0, m_px_s = 0;
        for (int k = gl; k < nR; k += G) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const double* XX = s.X + a * nA; const double* XU = XX + nR;
                const int j = a * nR + k;
                double ax = s.CM[j] * XX[k];
                if (k >= 1) ax += s.CA[j] * XX[k - 1] + s.CB[j] * XU[k - 1];
                const double z = (k == 0) ? lb0[a] : 0.0;
                const double einv = 1.0 / s.EE[j];
                const double r = ax - z;
                m_rp_s = amax(m_rp_s, r); m_rp_u = amax(m_rp_u, einv * r);
                m_z_s = amax(m_z_s, z); m_z_u = amax(m_z_u, einv * z);
                m_ax_s = amax(m_ax_s, ax); m_ax_u = amax(m_ax_u, einv * ax);
                double px_x, aty_x, px_u, aty_u;
                stage_products(a, k, s.X, s.YE, s.YH, px_x, aty_x, px_u, aty_u);
                {
                    const double q = s.QB[a * nA + k], dinv = 1.0 / s.DD[a * nA + k];
                    const double rr = q + px_x + aty_x;
                    m_rd_s = amax(m_rd_s, rr); m_rd_u = amax(m_rd_u, dinv * rr);
                    m_q_s = amax(m_q_s, q); m_q_u = amax(m_q_u, dinv * q);
                    m_aty_s = amax(m_aty_s, aty_x); m_aty_u = amax(m_aty_u, dinv * aty_x);
                    m_px_s = amax(m_px_s, px_x); m_px_u = amax(m_px_u, dinv * px_x);
                }
                if (k < N) {
                    const double q = s.QB[a * nA + nR + k], dinv = 1.0 / s.DD[a * nA + nR + k];
                    const double rr = q + px_u + aty_u;
                    m_rd_s = amax(m_rd_s, rr); m_rd_u = amax(m_rd_u, dinv * rr);
                    m_q_s = amax(m_q_s, q); m_q_u = amax(m_q_u, dinv * q);
                    m_aty_s = amax(m_aty_s, aty_u); m_aty_u = amax(m_aty_u, dinv * aty_u);
                    m_px_s = amax(m_px_s, px_u); m_px_u = amax(m_px_u, dinv * px_u);
                }
            }
        }
        for (int h = gl; h < nh; h += G) {
            const double ax = s.HX[h] * s.X[nR] + s.HY[h] * s.X[nA + nR];
            const double z = s.ZH[h];
            const double einv = 1.0 / s.EH[h];
            const double r = ax - z;
            m_rp_s = amax(m_rp_s, r); m_rp_u = amax(m_rp_u, einv * r);
            m_z_s = amax(m_z_s, z); m_z_u = amax(m_z_u, einv * z);
            m_ax_s = amax(m_ax_s, ax); m_ax_u = amax(m_ax_u, einv * ax);
        }
        const unsigned mk = gmask;
        pri_res = group_max<G>(m_rp_u, mk); rp_s = group_max<G>(m_rp_s, mk);
        nz_u = group_max<G>(m_z_u, mk); nz_s = group_max<G>(m_z_s, mk);
        nAx_u = group_max<G>(m_ax_u, mk); nAx_s = group_max<G>(m_ax_s, mk);
        dua_res = cinv * group_max<G>(m_rd_u, mk); rd_s = group_max<G>(m_rd_s, mk);
        nq_u = cinv * group_max<G>(m_q_u, mk); nq_s = group_max<G>(m_q_s, mk);
        nAty_u = cinv * group_max<G>(m_aty_u, mk); nAty_s = group_max<G>(m_aty_s, mk);
        nPx_u = cinv * group_max<G>(m_px_u, mk); nPx_s = group_max<G>(m_px_s, mk);
    };

    // is_primal_infeasible / is_dual_infeasible (OSQP auxil.c) on delta_y (DYE, DYH) and delta_x (W)
    auto primal_infeasible = [&](double eps) -> int {
        // project delta_y on the polar of the recession cone: hull rows have l = -inf -> max(dy, 0)
        double nrm = 0.0, lhs = 0.0;
        for (int j = gl; j < 2 * nR; j += G) {
            const double dy = s.DYE[j];
            nrm = amax(nrm, s.EE[j] * dy);
            const double bnd = (j == 0) ? lb0[0] : ((j == nR) ? lb0[1] : 0.0);
            lhs += bnd * fmax(dy, 0.0) + bnd * fmin(dy, 0.0);
        }
        for (int h = gl; h < nh; h += G) {
            double dy = s.DYH[h];
            const bool u_inf = s.UH[h] > infty * kMinScaling, l_inf = s.LH[h] < -infty * kMinScaling;
            if (u_inf) dy = l_inf ? 0.0 : fmin(dy, 0.0);
            else if (l_inf) dy = fmax(dy, 0.0);
            s.DYH[h] = dy;
            nrm = amax(nrm, s.EH[h] * dy);
            lhs += s.UH[h] * fmax(dy, 0.0) + s.LH[h] * fmin(dy, 0.0);
        }
        nrm = group_max<G>(nrm, gmask);
        lhs = group_sum<G>(lhs, gmask);
        __syncwarp(gmask);
        int res = 0;
        if (nrm > 1.0 / infty && lhs < -eps * nrm) {
            double m = 0.0;
            for (int k = gl; k < nR; k += G) {
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    double px_x, aty_x, px_u, aty_u;
                    stage_products(a, k, s.X, s.DYE, s.DYH, px_x, aty_x, px_u, aty_u);
                    m = amax(m, aty_x / s.DD[a * nA + k]);
                    if (k < N) m = amax(m, aty_u / s.DD[a * nA + nR + k]);
                }
            }
            m = group_max<G>(m, gmask);
            res = m < eps * nrm;
        }
        return res;
    };
    auto dual_infeasible = [&](double eps) -> int {
        double nrm = 0.0, qdx = 0.0;
        for (int j = gl; j < 2 * nA; j += G) {
            nrm = amax(nrm, s.DD[j] * s.W[j]);
            qdx += s.QB[j] * s.W[j];
        }
        nrm = group_max<G>(nrm, gmask);
        qdx = group_sum<G>(qdx, gmask);
        int res = 0;
        if (nrm > 1.0 / infty && qdx < -c * eps * nrm) {
            double m = 0.0;
            for (int k = gl; k < nR; k += G) {
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    double px_x, aty_x, px_u, aty_u;
                    stage_products(a, k, s.W, s.YE, s.YH, px_x, aty_x, px_u, aty_u);
                    m = amax(m, px_x / s.DD[a * nA + k]);
                    if (k < N) m = amax(m, px_u / s.DD[a * nA + nR + k]);
                }
            }
            m = group_max<G>(m, gmask);
            if (m < c * eps * nrm) {
                int bad = 0;
                for (int k = gl; k < nR; k += G) {
#pragma unroll
                    for (int a = 0; a < 2; a++) {
                        const double* WX = s.W + a * nA; const double* WU = WX + nR;
                        const int j = a * nR + k;
                        double ax = s.CM[j] * WX[k];
                        if (k >= 1) ax += s.CA[j] * WX[k - 1] + s.CB[j] * WU[k - 1];
                        ax /= s.EE[j];
                        if (ax > eps * nrm || ax < -eps * nrm) bad = 1; // both bounds finite on equality rows
                    }
                }
                for (int h = gl; h < nh; h += G) {
                    double ax = s.HX[h] * s.W[nR] + s.HY[h] * s.W[nA + nR];
                    ax /= s.EH[h];
                    if ((s.UH[h] < infty * kMinScaling && ax > eps * nrm) ||
                        (s.LH[h] > -infty * kMinScaling && ax < -eps * nrm)) bad = 1;
                }
                bad = (int)group_max<G>((double)bad, gmask);
                res = !bad;
            }
        }
        return res;
    };
    // check_termination (OSQP auxil.c); returns 1 when the algorithm stops, sets status
    auto check_termination = [&](int approximate) -> int {
        double eps_abs = st.eps_abs, eps_rel = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
        if (pri_res > infty || dua_res > infty) { status = WC_OSQP_NON_CVX; return 1; }
        if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
        int prc = 0, drc = 0;
        prim_inf = 0; dual_inf = 0;
        const double eps_prim = eps_abs + eps_rel * fmax(nz_u, nAx_u);
        if (pri_res < eps_prim) prc = 1; else prim_inf = primal_infeasible(epi);
        const double eps_dual = eps_abs + eps_rel * fmax(fmax(nq_u, nAty_u), nPx_u);
        if (dua_res < eps_dual) drc = 1; else dual_inf = dual_infeasible(edi);
        if (prc && drc) { status = approximate ? WC_OSQP_SOLVED_INACCURATE : WC_OSQP_SOLVED; return 1; }
        if (prim_inf) { status = approximate ? WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE : WC_OSQP_PRIMAL_INFEASIBLE; return 1; }
        if (dual_inf) { status = approximate ? WC_OSQP_DUAL_INFEASIBLE_INACCURATE : WC_OSQP_DUAL_INFEASIBLE; return 1; }
        return 0;
    };

    const int max_iter = st.max_iter;
    const int chk = st.check_termination;
    const int rint = d.rho_interval;
    int iter = 0;
    int to_check = chk, to_adapt = rint; // countdowns instead of modulo in the hot loop
    bool warp_busy = __any_sync(FULL, !done);
    while (warp_busy) {
        if (factor_pending) { // warp-uniform: the tick's factorisation, then once more after every rho update
            const bool ok = factor(rho);
            factor_pending = false;
            if (first_factor) {
                first_factor = false;
                setup_failed = !ok && setup;
                if (setup && !setup_failed && gl == 0) d.inited[i] = 1;
                if (setup_failed) done = true; // no solver: status stays UNSOLVED, the controller keeps its previous output
            }
        }
        iter++;
        const bool act = !done;
        const double rho_e = kRhoEqOverIneq * rho;
        const double rho_e_inv = 1.0 / rho_e, rho_inv = 1.0 / rho;
        bool check_now = false, adapt_now = false;
        if (chk) { if (--to_check == 0) { check_now = true; to_check = chk; } }
        if (st.adaptive_rho && rint) { if (--to_adapt == 0) { adapt_now = true; to_adapt = rint; } }
        const bool keep_delta = check_now || iter >= max_iter; // delta_x / delta_y feed the infeasibility tests only
        // ---- update_xz_tilde: right-hand side  sigma x - q + A' (rho z - y)
        if (act) {
            for (int k = gl; k < nR; k += G) {
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const double* Ya = s.YE + a * nR;
                    const double* CMa = s.CM + a * nR; const double* CAa = s.CA + a * nR; const double* CBa = s.CB + a * nR;
                    // w_k = rho_e (z_k - y_k / rho_e), z_k = 0 except equality row 0
                    const double zk = (k == 0) ? z0prev[a] : 0.0;
                    const double wk = rho_e * (zk - rho_e_inv * Ya[k]);
                    const double wk1 = (k < N) ? rho_e * (0.0 - rho_e_inv * Ya[k + 1]) : 0.0;
                    const int jx = a * nA + k;
                    double r = sigma * s.X[jx] - s.QB[jx];
                    r += CMa[k] * wk;
                    if (k < N) r += CAa[k + 1] * wk1;
                    s.W[jx] = r;
                    if (k < N) {
                        const int ju = a * nA + nR + k;
                        double ru = sigma * s.X[ju] - s.QB[ju];
                        ru += CBa[k + 1] * wk1;
                        if (k == 0) {
                            const double* Hc = (a == 0) ? s.HX : s.HY;
                            for (int h = 0; h < nh; h++) ru += Hc[h] * (rho * (s.ZH[h] - rho_inv * s.YH[h]));
                        }
                        s.W[ju] = ru;
                    }
                }
            }
        }
        __syncwarp();
        // ---- reduced KKT solve: the two axis chains run on two lanes (coupled weights: one block chain on one lane)
        if (cpl) {
            if (act && gl == 0) solve_block(s, N);
        } else if (par) {
            solve_par(act, rho);
        } else {
        if (act && gl < 2) sweep_forward(s, gl, N);
        __syncwarp();
        if (act && gl == 0) {
            double r[4] = {s.W[nR], s.W[nA + nR], s.W[0], s.W[nA]};
            solve_corner(s.CF, r);
            s.W[nR] = r[0]; s.W[nA + nR] = r[1]; s.W[0] = r[2]; s.W[nA] = r[3];
        }
        __syncwarp();
        if (act && gl < 2) sweep_backward(s, gl, N);
        }
        __syncwarp();
        // ---- update_z / update_y on the rows (z~ = A x~)
        if (act) {
            for (int k = gl; k < nR; k += G) {
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const double* WX = s.W + a * nA; const double* WU = WX + nR;
                    const int j = a * nR + k;
                    double zt = s.CM[j] * WX[k];
                    if (k >= 1) zt += s.CA[j] * WX[k - 1] + s.CB[j] * WU[k - 1];
                    const double zprev = (k == 0) ? z0prev[a] : 0.0;
                    const double znew = (k == 0) ? lb0[a] : 0.0; // projection on [l, u] with l == u
                    const double dy = rho_e * (alpha * zt + (1.0 - alpha) * zprev - znew);
                    s.YE[j] += dy;
                    if (keep_delta) s.DYE[j] = dy;
                }
            }
            for (int h = gl; h < nh; h += G) {
                const double zt = s.HX[h] * s.W[nR] + s.HY[h] * s.W[nA + nR];
                const double zp = s.ZH[h], y = s.YH[h];
                const double zr = alpha * zt + (1.0 - alpha) * zp;
                double zn = zr + rho_inv * y;
                zn = fmax(zn, s.LH[h]);
                zn = fmin(zn, s.UH[h]);
                const double dy = rho * (zr - zn);
                s.ZH[h] = zn;
                s.YH[h] = y + dy;
                if (keep_delta) s.DYH[h] = dy;
            }
        }
        __syncwarp();
        // ---- update_x
        if (act) {
            for (int j = gl; j < 2 * nA; j += G) {
                const double xp = s.X[j];
                const double xn = alpha * s.W[j] + (1.0 - alpha) * xp;
                s.X[j] = xn;
                if (keep_delta) s.W[j] = xn - xp;
            }
            z0prev[0] = lb0[0]; z0prev[1] = lb0[1];
            iter_done = iter;
        }
        __syncwarp();
        // ---- termination / adaptive rho (group-uniform control flow, warp-uniform barriers)
        if (check_now) {
            update_info();
            if (act) { if (check_termination(0)) done = true; }
        }
        if (adapt_now) {
            if (!check_now) update_info();
            if (act && !done) {
                // compute_rho_estimate + adapt_rho
                const double pr = rp_s / (fmax(nz_s, nAx_s) + 1e-10);
                const double dr = rd_s / (fmax(fmax(nq_s, nAty_s), nPx_s) + 1e-10);
                double rn = rho * sqrt(pr / (dr + 1e-10));
                rn = fmin(fmax(rn, kRhoMin), kRhoMax);
                if (rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance) {
                    rho = rn;
                    rho_updates++;
                }
            }
            __syncwarp();
            // warp-uniform; a no-op numerically for groups whose rho did not change.  Not after the last iteration: nothing is
            // solved any more and the segment-parallel factorisation stages its Schur blocks in W (delta_x of the final checks)
            if (iter < max_iter) factor_pending = true; // (at the top of the next iteration)
        }
        if (iter >= max_iter) { // warp-uniform: osqp_solve leaves the loop
            if (!check_now) {
                update_info();
                if (act && !done && check_termination(0)) done = true;
            }
            if (act && !done) {
                if (!check_termination(1)) status = WC_OSQP_MAX_ITER_REACHED;
                done = true;
            }
        }
        warp_busy = __any_sync(FULL, !done);
    }
    __syncwarp();

    // ------------------------------------------------------------------ epilogue: store_solution + controller checks
    const bool has_sol = !(status == WC_OSQP_PRIMAL_INFEASIBLE || status == WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                           status == WC_OSQP_DUAL_INFEASIBLE || status == WC_OSQP_DUAL_INFEASIBLE_INACCURATE ||
                           status == WC_OSQP_UNSOLVED || status == WC_OSQP_NON_CVX);
    if (!skip) {
        for (int j = gl; j < 2 * nA; j += G) d.x[(size_t)j * B + i] = has_sol ? s.X[j] : 0.0;
        for (int j = gl; j < 2 * nR; j += G) d.yeq[(size_t)j * B + i] = has_sol ? s.YE[j] : 0.0;
        for (int h = gl; h < kMaxHull; h += G) {
            d.yh[(size_t)h * B + i] = has_sol ? s.YH[h] : 0.0;
            d.zh[(size_t)h * B + i] = has_sol ? s.ZH[h] : 0.0;
        }
    }
    if (gl == 0 && valid) {
        int flags = kflags;
        double out[2] = {up[0], up[1]};
        if (!skip) {
            d.z0[i] = has_sol ? z0prev[0] : 0.0;
            d.z0[B + i] = has_sol ? z0prev[1] : 0.0;
            d.rho[i] = rho;
            d.rho_updates[i] = rho_updates;
            if (osqp_eigen_solve_ok(status, st.osqp_eigen_checks_status)) {
                out[0] = has_sol ? s.X[nR] * s.DD[nR] : NAN;
                out[1] = has_sol ? s.X[nA + nR] * s.DD[nA + nR] : NAN;
                d.uprev[i] = out[0]; d.uprev[B + i] = out[1];
                // ConvexHullProjectionConstraint::computeMargin (Controller.cpp:513-517)
                bool inside = true;
                double best = INFINITY;
                for (int h = 0; h < nh; h++) {
                    const double a0 = d.hullA[(size_t)(2 * h) * B + i], a1 = d.hullA[(size_t)(2 * h + 1) * B + i];
                    if (!(a0 * out[0] + a1 * out[1] <= d.hullb[(size_t)h * B + i])) inside = false;
                }
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

int mpc_admm_max_smem()
{
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return v;
}

template <int G>
static cudaError_t launch_admm_t(const MpcDev& d, const MpcTickIO& io, cudaStream_t s)
{
    constexpr int IPW = 32 / G;
    int stride = mpc_smem_doubles(d.N, d.coupled, d.par);
    // keep the per-instance stride == 8 (mod 16) doubles so that the two half-warps of a
    // cooperative access hit disjoint bank groups
    while ((stride % 16) != 8) stride++;
    const size_t per_inst = (size_t)stride * sizeof(double);
    const int max_smem = mpc_admm_max_smem();
    // aim for two resident CTAs per SM (228 KB total, 1 KB reserved per CTA)
    const size_t budget = (size_t)(max_smem > 116 * 1024 ? 113 * 1024 : max_smem);
    int warps = (int)(budget / (per_inst * IPW));
    if (warps < 1) {
        warps = 1;
        if (per_inst * IPW > (size_t)max_smem) return cudaErrorInvalidConfiguration;
    }
    if (warps > 4) warps = 4;
    const size_t smem = per_inst * IPW * warps;
    const int inst_per_block = warps * IPW;
    const int blocks = (d.B + inst_per_block - 1) / inst_per_block;
    cudaError_t e = cudaFuncSetAttribute(mpc_admm_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mpc_admm_kernel<G><<<blocks, warps * 32, smem, s>>>(d, io, stride);
    return cudaGetLastError();
}

cudaError_t launch_mpc_admm(const MpcDev& d, const MpcTickIO& io, int lanes, cudaStream_t s)
{
    switch (lanes) {
    case 2: return launch_admm_t<2>(d, io, s);
    case 4: return launch_admm_t<4>(d, io, s);
    case 8: return launch_admm_t<8>(d, io, s);
    case 16: return launch_admm_t<16>(d, io, s);
    case 32: return launch_admm_t<32>(d, io, s);
    default: return cudaErrorInvalidValue;
    }
}

// primal solution in the reference's variable order z = [x_0..x_N, u_0..u_{N-1}], unscaled (D x)
__global__ void mpc_export_primal_kernel(MpcDev d, double* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.B) return;
    const size_t B = (size_t)d.B;
    const int N = d.N, nA = 2 * N + 1, nR = N + 1;
    for (int a = 0; a < 2; a++) {
        for (int k = 0; k <= N; k++) {
            const size_t j = (size_t)(a * nA + k);
            out[(size_t)(2 * k + a) * B + i] = d.x[j * B + i] * d.D[j * B + i];
        }
        for (int k = 0; k < N; k++) {
            const size_t j = (size_t)(a * nA + nR + k);
            out[(size_t)(2 * (N + 1) + 2 * k + a) * B + i] = d.x[j * B + i] * d.D[j * B + i];
        }
    }
}

void launch_mpc_export_primal(const MpcDev& d, double* primal_out, cudaStream_t s)
{
    const int threads = 128;
    mpc_export_primal_kernel<<<(d.B + threads - 1) / threads, threads, 0, s>>>(d, primal_out);
}

} // namespace wcb
