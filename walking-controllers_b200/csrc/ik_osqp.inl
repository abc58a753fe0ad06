// ik_osqp.inl — WalkingQPIK_osqp on the GPU; included by ik_kernels.cu inside its namespaces so that it
// shares the kinematics (fk_stage) and the QP vectors (qp_vectors<COMC, 1>) with the qpOASES variant.
//
// Replaces, per instance and tick,
//   WalkingQPIK_osqp::{setHessianMatrix,setGradientVector,setLinearConstraintMatrix,setBounds,solve,isSolutionFeasible,
//                      getSolution,getLeftFootError,getRightFootError}
//                                              modules/Walking_module/src/WalkingQPInverseKinematics_osqp.cpp:135-454
//   OsqpEigen::Solver::{initSolver,updateHessianMatrix,updateGradient,updateLinearConstraintsMatrix,updateBounds,solve}
//   and OSQP 0.6 underneath (external; life cycle restated in oracle/qpik_osqp_ref.c, whose header lists the deviations).
//
// One warp per instance, lane j <-> QP variable j (29 <= 32), lane r <-> task row r (15 dense rows of A) and lane v <-> the
// empty "velocity limit" row v (23 rows with no entries: they only take part in the norms).  Every tick the problem data
// change, so every tick runs OSQP's scale_data (ten Ruiz passes, in place, same operation order as OSQP), rebuilds the
// reduced KKT matrix K = P + sigma I + rho_eq A'A (29 x 29, instead of the 67 x 67 quasi-definite KKT matrix: same x~, z~ = A x~)
// and its Cholesky factor L, inverts L once (column per lane) so that every K x~ = rhs is two triangular mat-vecs; the ADMM iterates (x, z, y; scaled) and rho persist in HBM between ticks as OSQP's workspace does.
// Shared memory per warp: P [29][33], A [15][33], K/L [29][33] (aliases the kinematics scratch), five 32-vectors = 20.6 KB.

template <int NJ, bool COMC>
struct IkOsqpLayout {
    static constexpr int NV = NJ + 6;
    static constexpr int NC = COMC ? 15 : 12;
    static constexpr int LD = 33;
    static constexpr int R1_KIN = WC_MAX_LINKS * 12 + WC_MAX_LINKS * 4;
    static constexpr int R1 = R1_KIN > NV * LD ? R1_KIN : NV * LD;
    static constexpr int O_WL = 0, O_MC = WC_MAX_LINKS * 12, O_KS = 0;
    static constexpr int O_PS = R1;
    static constexpr int O_AS = O_PS + NV * LD;
    static constexpr int O_DT = O_AS + NC * LD, O_ET = O_DT + 32, O_XT = O_ET + 32, O_WV = O_XT + 32, O_NR = O_WV + 32;
    static constexpr int TOTAL = ((O_NR + 32 + 15) / 16) * 16 + 8;
};

constexpr int kOsqpWarps = 4;
constexpr double kOsqpRhoMin = 1e-6, kOsqpRhoMax = 1e6, kOsqpRhoEq = 1e3, kOsqpRhoTol = 1e-4;
constexpr double kOsqpMinScaling = 1e-4, kOsqpMaxScaling = 1e4;

__device__ __forceinline__ double osqp_limit_scaling(double v)
{
    v = v < kOsqpMinScaling ? 1.0 : v;
    v = v > kOsqpMaxScaling ? kOsqpMaxScaling : v;
    return v;
}


// reductions with four independent chains: a warp has at most one other warp to hide the 8-cycle DFMA / DMNMX latency behind
template <int N, int SA>
__device__ __forceinline__ double osqp_dot(const double* a, const double* b)
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k + 3 < N; k += 4) {
        s0 += a[k * SA] * b[k]; s1 += a[(k + 1) * SA] * b[k + 1]; s2 += a[(k + 2) * SA] * b[k + 2]; s3 += a[(k + 3) * SA] * b[k + 3];
    }
#pragma unroll
    for (int k = N - N % 4; k < N; k++) s0 += a[k * SA] * b[k];
    return (s0 + s1) + (s2 + s3);
}
template <int N, int SA>
__device__ __forceinline__ double osqp_absmax(const double* a)
{
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
#pragma unroll
    for (int k = 0; k + 3 < N; k += 4) {
        m0 = fmax(m0, fabs(a[k * SA])); m1 = fmax(m1, fabs(a[(k + 1) * SA])); m2 = fmax(m2, fabs(a[(k + 2) * SA])); m3 = fmax(m3, fabs(a[(k + 3) * SA]));
    }
#pragma unroll
    for (int k = N - N % 4; k < N; k++) m0 = fmax(m0, fabs(a[k * SA]));
    return fmax(fmax(m0, m1), fmax(m2, m3));
}

template <int NJ, bool COMC>
__global__ void __launch_bounds__(kOsqpWarps * 32) ik_osqp_kernel(IkDev d, IkTickIO io, int mode)
{
    using LY = IkOsqpLayout<NJ, COMC>;
    constexpr int NV = LY::NV, NC = LY::NC, LD = LY::LD;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long inst = (long long)blockIdx.x * kOsqpWarps + warp;
    if (inst >= (long long)d.B) return; // whole warp exits together
    const int i = (int)inst;
    const size_t B = (size_t)d.B;
    double* sm = smem_all + (size_t)warp * LY::TOTAL;
    double* WL = sm + LY::O_WL; double* MC = sm + LY::O_MC; double* KS = sm + LY::O_KS;
    double* PS = sm + LY::O_PS; double* AS = sm + LY::O_AS;
    double* DT = sm + LY::O_DT; double* ET = sm + LY::O_ET; double* XT = sm + LY::O_XT; double* WV = sm + LY::O_WV;
    double* NR = sm + LY::O_NR;
    const IkModelDev& M = *d.model;
    const wc_ik_config& cfg = d.cfg;
    const wc_osqp_settings& st = d.osqp;
    const int j = lane;
    const bool isvar = j < NV, isjoint = j >= 6 && j < NV;
    const bool isrow = lane < NC;  // task row of this lane
    const bool isvel = lane < NJ;  // velocity-limit row of this lane

    // ---- kinematics: this lane's Jacobian columns and the replicated frame data
    double colLf[6] = {0, 0, 0, 0, 0, 0}, colRf[6] = {0, 0, 0, 0, 0, 0}, colN[3] = {0, 0, 0}, colC[3] = {0, 0, 0};
    double lfH[12], rfH[12], nR[9], com[3], pBv[3] = {0, 0, 0};
    if (mode != IK_MODE_EXT_KIN) {
        fk_stage<NV>(M, io, i, B, lane, WL, MC, colLf, colRf, colN, colC, lfH, rfH, nR, com, pBv);
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) { lfH[k] = io.lf_H[(size_t)k * B + i]; rfH[k] = io.rf_H[(size_t)k * B + i]; }
#pragma unroll
        for (int k = 0; k < 9; k++) nR[k] = io.neck_R[(size_t)k * B + i];
        if (isvar) {
#pragma unroll
            for (int r = 0; r < 6; r++) {
                colLf[r] = io.J_lf[(size_t)(r * NV + j) * B + i];
                colRf[r] = io.J_rf[(size_t)(r * NV + j) * B + i];
            }
#pragma unroll
            for (int r = 0; r < 3; r++) {
                colN[r] = io.J_neck[(size_t)((3 + r) * NV + j) * B + i];
                colC[r] = io.J_com[(size_t)(r * NV + j) * B + i];
            }
        }
    }
    if (!isvar) {
#pragma unroll
        for (int r = 0; r < 6; r++) { colLf[r] = 0.0; colRf[r] = 0.0; }
#pragma unroll
        for (int r = 0; r < 3; r++) { colN[r] = 0.0; colC[r] = 0.0; }
    }
    double bvec[16], gj, lam;
    qp_vectors<COMC, 1>(d, io, i, B, i, lane, lfH, rfH, nR, colN, colC, bvec, gj, lam);
    __syncwarp(); // the kinematics scratch (WL, MC) is dead from here on: K aliases it

    // ---- setHessianMatrix (:135-167): H = Lambda + Jn' Wn Jn (+ Jc' Wc Jc), column j by lane j, no FMA contraction so that
    // exact zeros (the sparsity pattern sparseView() keeps) are decided as on the CPU
    double WJ[3], WcJ[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        WJ[r] = __dadd_rn(__dadd_rn(__dmul_rn(cfg.neck_weight[3 * r], colN[0]), __dmul_rn(cfg.neck_weight[3 * r + 1], colN[1])),
                          __dmul_rn(cfg.neck_weight[3 * r + 2], colN[2]));
        WcJ[r] = __dadd_rn(__dadd_rn(__dmul_rn(cfg.com_weight[3 * r], colC[0]), __dmul_rn(cfg.com_weight[3 * r + 1], colC[1])),
                           __dmul_rn(cfg.com_weight[3 * r + 2], colC[2]));
    }
    uint32_t mask = 0;
    for (int r = 0; r < NV; r++) {
        const double n0 = __shfl_sync(FULL, colN[0], r), n1 = __shfl_sync(FULL, colN[1], r), n2 = __shfl_sync(FULL, colN[2], r);
        double s = (r == j && isjoint) ? lam : 0.0;
        s = __dadd_rn(s, __dmul_rn(n0, WJ[0])); s = __dadd_rn(s, __dmul_rn(n1, WJ[1])); s = __dadd_rn(s, __dmul_rn(n2, WJ[2]));
        if (!COMC) {
            const double c0 = __shfl_sync(FULL, colC[0], r), c1 = __shfl_sync(FULL, colC[1], r), c2 = __shfl_sync(FULL, colC[2], r);
            s = __dadd_rn(s, __dmul_rn(c0, WcJ[0])); s = __dadd_rn(s, __dmul_rn(c1, WcJ[1])); s = __dadd_rn(s, __dmul_rn(c2, WcJ[2]));
        }
        if (isvar) PS[r * LD + j] = s;
        if (isvar && r <= j && s != 0.0) mask |= 1u << r;
    }
    // ---- setLinearConstraintMatrix (:218-262): task rows = the feet (and CoM) Jacobians, velocity rows empty
    double acol[NC];
#pragma unroll
    for (int r = 0; r < 6; r++) { acol[r] = colLf[r]; acol[6 + r] = colRf[r]; }
    if (COMC) {
#pragma unroll
        for (int r = 0; r < 3; r++) acol[12 + r] = colC[r];
    }
#pragma unroll
    for (int r = 0; r < NC; r++) AS[r * LD + lane] = acol[r];
    // ---- which life-cycle branch (oracle/qpik_osqp_ref.c): first solve / same pattern / pattern changed
    const bool inited = d.oinit[i] != 0;
    const uint32_t old_mask = isvar ? d.omask[(size_t)j * B + i] : 0u;
    const bool same = inited && !__any_sync(FULL, isvar && mask != old_mask);
    __syncwarp();

    // ---- scale_data (OSQP scaling.c), in place; D of this lane's variable, E of this lane's task row (velocity rows: E = 1)
    double Dj = 1.0, Er = 1.0, c = 1.0, qj = isvar ? gj : 0.0;
    double c_pend = 1.0; // cost scaling of the previous pass, applied to P while the next pass reads it for its norms
    for (int it = 0; it < st.scaling; it++) {
        double dn = 0.0, en = 0.0;
        if (isvar) {
            double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
#pragma unroll
            for (int r = 0; r + 3 < NV; r += 4) {
                const double v0 = PS[r * LD + j] * c_pend, v1 = PS[(r + 1) * LD + j] * c_pend, v2 = PS[(r + 2) * LD + j] * c_pend,
                             v3 = PS[(r + 3) * LD + j] * c_pend;
                PS[r * LD + j] = v0; PS[(r + 1) * LD + j] = v1; PS[(r + 2) * LD + j] = v2; PS[(r + 3) * LD + j] = v3;
                m0 = fmax(m0, fabs(v0)); m1 = fmax(m1, fabs(v1)); m2 = fmax(m2, fabs(v2)); m3 = fmax(m3, fabs(v3));
            }
#pragma unroll
            for (int r = NV - NV % 4; r < NV; r++) { const double v = PS[r * LD + j] * c_pend; PS[r * LD + j] = v; m0 = fmax(m0, fabs(v)); }
            dn = fmax(fmax(m0, m1), fmax(m2, m3));
            double an0 = 0.0, an1 = 0.0;
#pragma unroll
            for (int r = 0; r + 1 < NC; r += 2) { an0 = fmax(an0, fabs(acol[r])); an1 = fmax(an1, fabs(acol[r + 1])); }
            if (NC % 2) an0 = fmax(an0, fabs(acol[NC - 1]));
            dn = fmax(dn, fmax(an0, an1));
        }
        if (isrow) en = osqp_absmax<NV, 1>(AS + lane * LD);
        const double dt = 1.0 / sqrt(osqp_limit_scaling(dn)), et = 1.0 / sqrt(osqp_limit_scaling(en));
        DT[lane] = dt; ET[lane] = et;
        __syncwarp();
        double pn = 0.0;
        if (isvar) {
            double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
#pragma unroll
            for (int r = 0; r + 3 < NV; r += 4) {
                const double v0 = (PS[r * LD + j] * DT[r]) * dt, v1 = (PS[(r + 1) * LD + j] * DT[r + 1]) * dt,
                             v2 = (PS[(r + 2) * LD + j] * DT[r + 2]) * dt, v3 = (PS[(r + 3) * LD + j] * DT[r + 3]) * dt;
                PS[r * LD + j] = v0; PS[(r + 1) * LD + j] = v1; PS[(r + 2) * LD + j] = v2; PS[(r + 3) * LD + j] = v3;
                m0 = fmax(m0, fabs(v0)); m1 = fmax(m1, fabs(v1)); m2 = fmax(m2, fabs(v2)); m3 = fmax(m3, fabs(v3));
            }
#pragma unroll
            for (int r = NV - NV % 4; r < NV; r++) { const double v = (PS[r * LD + j] * DT[r]) * dt; PS[r * LD + j] = v; m0 = fmax(m0, fabs(v)); }
            pn = fmax(fmax(m0, m1), fmax(m2, m3));
#pragma unroll
            for (int r = 0; r < NC; r++) { acol[r] = (acol[r] * ET[r]) * dt; AS[r * LD + j] = acol[r]; }
            qj *= dt; Dj *= dt;
        }
        if (isrow) Er *= et;
        NR[lane] = pn;
        __syncwarp();
        double c_temp = 0.0;
        for (int k = 0; k < NV; k++) c_temp += NR[k];
        c_temp /= (double)NV;
        double qn = warp_max(fabs(qj));
        qn = osqp_limit_scaling(qn);
        c_temp = qn > c_temp ? qn : c_temp;
        c_temp = 1.0 / osqp_limit_scaling(c_temp);
        c_pend = c_temp;
        qj *= c_temp; c *= c_temp;
        __syncwarp();
    }
    if (isvar && st.scaling > 0)
        for (int r = 0; r < NV; r++) PS[r * LD + j] *= c_pend;
    __syncwarp();
    const double cinv = 1.0 / c, Dinv = 1.0 / Dj, Einv = 1.0 / Er;
    // bounds (setBounds :264-347, setVelocityBounds :23-51), scaled by E
    double b_r = 0.0;
#pragma unroll
    for (int r = 0; r < NC; r++) b_r = lane == r ? bvec[r] : b_r;
    const double lt = b_r * Er;                                   // task rows: l = u
    const double lv = isvel ? cfg.min_vel[isvel ? lane : 0] : 0.0; // velocity rows: E = 1
    const double uv = isvel ? cfg.max_vel[isvel ? lane : 0] : 0.0;
    const double infty = st.infty;
    // constraint type of the velocity rows (set_rho_vec / update_rho_vec): -1 loose, 1 equality, 0 inequality
    const int vtype = (lv < -infty * kOsqpMinScaling && uv > infty * kOsqpMinScaling) ? -1 : ((uv - lv < kOsqpRhoTol) ? 1 : 0);

    // ---- iterates and rho: cold start, carried over, or re-scaled after a re-setup
    double rho, x = 0.0, yt = 0.0, zt = 0.0, yv = 0.0, zv = 0.0;
    int setups = inited ? d.osetups[i] : 0;
    if (same) {
        rho = d.orho[i];
        if (isvar) x = d.ox[(size_t)j * B + i];
        if (isrow) { yt = d.oy[(size_t)lane * B + i]; zt = d.oz[(size_t)lane * B + i]; }
        if (isvel) { yv = d.oy[(size_t)(NC + lane) * B + i]; zv = d.oz[(size_t)(NC + lane) * B + i]; }
    } else {
        rho = fmin(fmax(st.rho, kOsqpRhoMin), kOsqpRhoMax);
        setups += 1;
        if (inited) { // osqp_warm_start with the iterates of the cleared solver
            if (isvar) x = d.ox[(size_t)j * B + i] * Dinv;
            if (isrow) yt = (d.oy[(size_t)lane * B + i] * Einv) * c;
            if (isvel) yv = d.oy[(size_t)(NC + lane) * B + i] * c;
            XT[lane] = x;
            __syncwarp();
            if (isrow) zt = osqp_dot<NV, 1>(AS + lane * LD, XT);
            zv = 0.0;
            __syncwarp();
        }
    }

    // ---- reduced KKT matrix and its Cholesky factor (row i of L on lane i)
    double dinv = 1.0;
    bool spd = true;
    auto factor = [&](double rho_now) {
        const double rho_e = kOsqpRhoEq * rho_now;
        if (isvar) {
            for (int r = 0; r < NV; r++) {
                double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
                for (int k = 0; k < NC; k += 3) { c0 += AS[k * LD + r] * acol[k]; c1 += AS[(k + 1) * LD + r] * acol[k + 1]; c2 += AS[(k + 2) * LD + r] * acol[k + 2]; }
                const double acc = (c0 + c1) + c2;
                KS[r * LD + j] = (PS[r * LD + j] + (r == j ? st.sigma : 0.0)) + rho_e * acc;
            }
        }
        __syncwarp();
        // left-looking Cholesky: column k needs rows i >= k minus the dot product of rows i and k over the finished columns
        // (row i on lane i: conflict-free; row k: broadcast) — reads only, one store per entry, one barrier per column
        for (int k = 0; k < NV; k++) {
            double sv = 0.0;
            if (isvar && lane >= k) {
                const double* ri = KS + lane * LD; const double* rk = KS + k * LD;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int p = 0;
                for (; p + 3 < k; p += 4) { a0 += ri[p] * rk[p]; a1 += ri[p + 1] * rk[p + 1]; a2 += ri[p + 2] * rk[p + 2]; a3 += ri[p + 3] * rk[p + 3]; }
                for (; p < k; p++) a0 += ri[p] * rk[p];
                sv = ri[k] - ((a0 + a1) + (a2 + a3));
            }
            const double pk = __shfl_sync(FULL, sv, k);
            if (!(pk > 0.0)) spd = false;
            const double piv = sqrt(pk), pinv = 1.0 / piv;
            if (isvar && lane >= k) KS[lane * LD + k] = lane == k ? piv : sv * pinv;
            __syncwarp();
        }
        dinv = isvar ? 1.0 / KS[lane * LD + lane] : 1.0;
        // X = L^-1 (lower triangular), column c by lane c, stored transposed in the free upper triangle of the same buffer
        // (X[i][c] at KS[c * LD + i], i > c; X[c][c] = dinv stays in a register): the 25+ solves of the ADMM loop then are two
        // triangular mat-vecs with independent lanes instead of 2 x 29 shuffle-serialised substitution steps
        DT[lane] = dinv;
        __syncwarp();
        if (isvar) {
            for (int r = 1; r < NV; r++) {
                double a0 = 0.0, a1 = 0.0;
                int p = 0;
                for (; p + 1 < r; p += 2) {
                    const double x0 = p > lane ? KS[lane * LD + p] : (p == lane ? dinv : 0.0);
                    const double x1 = p + 1 > lane ? KS[lane * LD + p + 1] : (p + 1 == lane ? dinv : 0.0);
                    a0 += KS[r * LD + p] * x0; a1 += KS[r * LD + p + 1] * x1;
                }
                if (p < r) a0 += KS[r * LD + p] * (p > lane ? KS[lane * LD + p] : (p == lane ? dinv : 0.0));
                if (r > lane) KS[lane * LD + r] = -(a0 + a1) * DT[r];
            }
        }
        __syncwarp();
    };
    factor(rho);

    // ---- residuals and norms of update_info / compute_rho_estimate for the current (x, z, y)
    double pri_res = 0, dua_res = 0, nz_u = 0, nAx_u = 0, nq_u = 0, nAty_u = 0, nPx_u = 0;
    double rp_s = 0, rd_s = 0, nz_s = 0, nAx_s = 0, nq_s = 0, nAty_s = 0, nPx_s = 0;
    auto update_info = [&]() {
        XT[lane] = x; WV[lane] = isrow ? yt : 0.0;
        __syncwarp();
        double ax = 0.0;
        if (isrow) ax = osqp_dot<NV, 1>(AS + lane * LD, XT);
        const double rp_t = ax - zt, rp_v = 0.0 - zv;
        rp_s = warp_max(fmax(isrow ? fabs(rp_t) : 0.0, isvel ? fabs(rp_v) : 0.0));
        pri_res = warp_max(fmax(isrow ? fabs(Einv * rp_t) : 0.0, isvel ? fabs(rp_v) : 0.0));
        nz_s = warp_max(fmax(isrow ? fabs(zt) : 0.0, isvel ? fabs(zv) : 0.0));
        nz_u = warp_max(fmax(isrow ? fabs(Einv * zt) : 0.0, isvel ? fabs(zv) : 0.0));
        nAx_s = warp_max(isrow ? fabs(ax) : 0.0);
        nAx_u = warp_max(isrow ? fabs(Einv * ax) : 0.0);
        double px = 0.0, aty = 0.0;
        if (isvar) {
            px = osqp_dot<NV, LD>(PS + j, XT);
            double y0 = 0.0, y1 = 0.0, y2 = 0.0;
#pragma unroll
            for (int r = 0; r < NC; r += 3) { y0 += acol[r] * WV[r]; y1 += acol[r + 1] * WV[r + 1]; y2 += acol[r + 2] * WV[r + 2]; }
            aty = (y0 + y1) + y2;
        }
        const double rd = (qj + px) + aty;
        rd_s = warp_max(fabs(rd)); dua_res = cinv * warp_max(fabs(Dinv * rd));
        nq_s = warp_max(fabs(qj)); nq_u = warp_max(fabs(Dinv * qj));
        nAty_s = warp_max(fabs(aty)); nAty_u = warp_max(fabs(Dinv * aty));
        nPx_s = warp_max(fabs(px)); nPx_u = warp_max(fabs(Dinv * px));
        __syncwarp();
    };
    double dx = 0.0, dyt = 0.0, dyv = 0.0; // delta_x, delta_y of the last iteration
    auto primal_infeasible = [&](double eps) -> bool {
        double pv = dyv; // projection on the polar of the recession cone: only rows with an infinite bound change
        if (uv > infty * kOsqpMinScaling) pv = (lv < -infty * kOsqpMinScaling) ? 0.0 : fmin(pv, 0.0);
        else if (lv < -infty * kOsqpMinScaling) pv = fmax(pv, 0.0);
        if (!isvel) pv = 0.0;
        dyv = pv;
        const double pt = isrow ? dyt : 0.0;
        const double nrm = warp_max(fmax(fabs(Er * pt), fabs(pv)));
        if (!(nrm > 1.0 / infty)) return false;
        double lhs = (isrow ? lt * fmax(pt, 0.0) + lt * fmin(pt, 0.0) : 0.0) + (isvel ? uv * fmax(pv, 0.0) + lv * fmin(pv, 0.0) : 0.0);
        lhs = warp_sum(lhs);
        if (!(lhs < -eps * nrm)) return false;
        WV[lane] = pt;
        __syncwarp();
        double a = 0.0;
#pragma unroll
        for (int r = 0; r < NC; r++) a += acol[r] * WV[r];
        const double m = warp_max(fabs(a * Dinv));
        __syncwarp();
        return m < eps * nrm;
    };
    auto dual_infeasible = [&](double eps) -> bool {
        const double nrm = warp_max(fabs(Dj * dx));
        if (!(nrm > 1.0 / infty)) return false;
        const double qdx = warp_sum(qj * dx);
        if (!(qdx < -c * eps * nrm)) return false;
        XT[lane] = dx;
        __syncwarp();
        double pdx = 0.0, adx = 0.0;
        if (isvar) pdx = osqp_dot<NV, LD>(PS + j, XT);
        if (isrow) adx = osqp_dot<NV, 1>(AS + lane * LD, XT);
        const double m = warp_max(fabs(pdx * Dinv));
        __syncwarp();
        if (!(m < c * eps * nrm)) return false;
        adx *= Einv;
        // task rows have finite bounds on both sides; the empty velocity rows have A dx = 0 and never object
        const bool bad = isrow && (adx > eps * nrm || adx < -eps * nrm);
        return !__any_sync(FULL, bad);
    };
    auto check_termination = [&](int approximate) -> int {
        double eps_abs = st.eps_abs, eps_rel = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
        if (pri_res > infty || dua_res > infty) return WC_OSQP_NON_CVX;
        if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
        bool prc = false, drc = false, pinf = false, dinf = false;
        const double eps_prim = eps_abs + eps_rel * fmax(nz_u, nAx_u);
        if (pri_res < eps_prim) prc = true; else pinf = primal_infeasible(epi);
        const double eps_dual = eps_abs + eps_rel * (cinv * fmax(fmax(nq_u, nAty_u), nPx_u));
        if (dua_res < eps_dual) drc = true; else dinf = dual_infeasible(edi);
        if (prc && drc) return approximate ? WC_OSQP_SOLVED_INACCURATE : WC_OSQP_SOLVED;
        if (pinf) return approximate ? WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE : WC_OSQP_PRIMAL_INFEASIBLE;
        if (dinf) return approximate ? WC_OSQP_DUAL_INFEASIBLE_INACCURATE : WC_OSQP_DUAL_INFEASIBLE;
        return 0;
    };

    // ---- osqp_solve
    const double sigma = st.sigma, alpha = st.alpha;
    const int chk = st.check_termination, max_iter = st.max_iter;
    const int rint = st.adaptive_rho ? (st.adaptive_rho_interval ? st.adaptive_rho_interval : (chk ? 4 * chk : 100)) : 0;
    int status = WC_OSQP_UNSOLVED, iter = 0, info_iter = 0;
    bool checked_last = false;
    if (!spd) status = WC_OSQP_NON_CVX;
    while (spd && iter < max_iter) {
        iter++;
        const double rho_e = kOsqpRhoEq * rho, rho_e_inv = 1.0 / rho_e;
        const double rho_v = vtype == -1 ? kOsqpRhoMin : (vtype == 1 ? rho_e : rho), rho_v_inv = 1.0 / rho_v;
        const double x_prev = x, zt_prev = zt, zv_prev = zv;
        // right-hand side sigma x - q + A' rho (z - y / rho), then K x~ = rhs by the two triangular sweeps
        WV[lane] = isrow ? rho_e * (zt_prev - rho_e_inv * yt) : 0.0;
        __syncwarp();
        double t = sigma * x_prev - qj;
        {
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
            for (int r = 0; r < NC; r += 3) { r0 += acol[r] * WV[r]; r1 += acol[r + 1] * WV[r + 1]; r2 += acol[r + 2] * WV[r + 2]; }
            t += (r0 + r1) + r2;
        }
        if (!isvar) t = 0.0;
        XT[lane] = t;
        __syncwarp();
        { // L^-1 rhs, then L^-T of that: uniform, fully unrolled, predicated on the triangle (no per-lane trip counts)
            double u0 = dinv * t, u1 = 0.0, u2 = 0.0, u3 = 0.0;
            const double* xc = KS + lane; // X[lane][k] at KS[k * LD + lane], k < lane
#pragma unroll
            for (int k = 0; k + 3 < NV - 1; k += 4) {
                u0 += (k < lane ? xc[k * LD] : 0.0) * XT[k]; u1 += (k + 1 < lane ? xc[(k + 1) * LD] : 0.0) * XT[k + 1];
                u2 += (k + 2 < lane ? xc[(k + 2) * LD] : 0.0) * XT[k + 2]; u3 += (k + 3 < lane ? xc[(k + 3) * LD] : 0.0) * XT[k + 3];
            }
#pragma unroll
            for (int k = (NV - 1) - (NV - 1) % 4; k < NV - 1; k++) u0 += (k < lane ? xc[k * LD] : 0.0) * XT[k];
            const double tf = isvar ? (u0 + u1) + (u2 + u3) : 0.0;
            ET[lane] = tf;
            __syncwarp();
            double v0 = dinv * tf, v1 = 0.0, v2 = 0.0, v3 = 0.0;
            const double* xr = KS + (isvar ? lane : 0) * LD; // X[k][lane] at KS[lane * LD + k], k > lane
#pragma unroll
            for (int k = 1; k + 3 < NV; k += 4) {
                v0 += (k > lane ? xr[k] : 0.0) * ET[k]; v1 += (k + 1 > lane ? xr[k + 1] : 0.0) * ET[k + 1];
                v2 += (k + 2 > lane ? xr[k + 2] : 0.0) * ET[k + 2]; v3 += (k + 3 > lane ? xr[k + 3] : 0.0) * ET[k + 3];
            }
#pragma unroll
            for (int k = 1 + ((NV - 1) / 4) * 4; k < NV; k++) v0 += (k > lane ? xr[k] : 0.0) * ET[k];
            t = isvar ? (v0 + v1) + (v2 + v3) : 0.0;
        }
        __syncwarp();
        XT[lane] = t; // x~
        __syncwarp();
        double zt_t = 0.0; // z~ = A x~ (zero on the velocity rows)
        if (isrow) zt_t = osqp_dot<NV, 1>(AS + lane * LD, XT);
        // update_x, update_z, update_y
        x = alpha * t + (1.0 - alpha) * x_prev;
        dx = x - x_prev;
        if (isrow) {
            const double zr = alpha * zt_t + (1.0 - alpha) * zt_prev;
            double zn = zr + rho_e_inv * yt;
            zn = zn > lt ? zn : lt; zn = zn < lt ? zn : lt;
            zt = zn;
            dyt = rho_e * (zr - zn);
            yt += dyt;
        }
        if (isvel) {
            const double zr = alpha * 0.0 + (1.0 - alpha) * zv_prev;
            double zn = zr + rho_v_inv * yv;
            zn = zn > lv ? zn : lv; zn = zn < uv ? zn : uv;
            zv = zn;
            dyv = rho_v * (zr - zn);
            yv += dyv;
        }
        __syncwarp();
        checked_last = chk && (iter % chk == 0);
        if (checked_last) {
            update_info(); info_iter = iter;
            const int s = check_termination(0);
            if (s) { status = s; break; }
        }
        if (rint && (iter % rint == 0)) {
            if (!checked_last) { update_info(); info_iter = iter; }
            const double pr = rp_s / (fmax(nz_s, nAx_s) + 1e-10);
            const double dr = rd_s / (fmax(fmax(nq_s, nAty_s), nPx_s) + 1e-10);
            double rn = rho * sqrt(pr / (dr + 1e-10));
            rn = fmin(fmax(rn, kOsqpRhoMin), kOsqpRhoMax);
            if (rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance) {
                rho = rn;
                factor(rho);
                if (!spd) { status = WC_OSQP_NON_CVX; break; }
            }
        }
    }
    if (spd && status == WC_OSQP_UNSOLVED) {
        if (!checked_last) { update_info(); info_iter = iter; const int s = check_termination(0); if (s) status = s; }
        if (status == WC_OSQP_UNSOLVED) { const int s = check_termination(1); status = s ? s : WC_OSQP_MAX_ITER_REACHED; }
    }

    // ---- store_solution, isSolutionFeasible (:395-408), getSolution, get*FootError
    const bool has_sol = !(status == WC_OSQP_PRIMAL_INFEASIBLE || status == WC_OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                           status == WC_OSQP_DUAL_INFEASIBLE || status == WC_OSQP_DUAL_INFEASIBLE_INACCURATE ||
                           status == WC_OSQP_UNSOLVED || status == WC_OSQP_NON_CVX);
    const double xs = has_sol ? x * Dj : nan("");
    if (!has_sol) { x = 0.0; yt = zt = yv = zv = 0.0; }
    double ferr[12], axc[3] = {0, 0, 0};
#pragma unroll
    for (int r = 0; r < 6; r++) {
        ferr[r] = bvec[r] - warp_sum(isvar ? colLf[r] * xs : 0.0);
        ferr[6 + r] = bvec[6 + r] - warp_sum(isvar ? colRf[r] * xs : 0.0);
    }
    if (COMC) {
#pragma unroll
        for (int r = 0; r < 3; r++) axc[r] = warp_sum(isvar ? colC[r] * xs : 0.0);
    }
    // bit 0: m_optimizerSolver->solve() false (..._osqp.cpp:377): a non-convex problem, and with osqp-eigen >= 0.4
    // (osqp_eigen_checks_status) every status but OSQP_SOLVED -- the same rule as the MPC kernels
    int flags = (status == WC_OSQP_NON_CVX || status == WC_OSQP_UNSOLVED || (st.osqp_eigen_checks_status && status != WC_OSQP_SOLVED)) ? 1 : 0;
    {
        double mu = -INFINITY, ml = -INFINITY;
#pragma unroll
        for (int r = 0; r < 12; r++) { const double a = bvec[r] - ferr[r]; mu = fmax(mu, a - bvec[r]); ml = fmax(ml, a - bvec[r]); }
        if (COMC) {
#pragma unroll
            for (int r = 0; r < 3; r++) { mu = fmax(mu, axc[r] - bvec[12 + r]); ml = fmax(ml, axc[r] - bvec[12 + r]); }
        }
        mu = fmax(mu, warp_max(isvel ? 0.0 - uv : -INFINITY));
        ml = fmax(ml, warp_max(isvel ? 0.0 - lv : -INFINITY));
        if (!((mu < 1.0) || (ml > -1.0))) flags |= 2;
    }
    if (isjoint && io.dq) io.dq[(size_t)(j - 6) * B + i] = xs;
    {
        double el = 0.0, er = 0.0;
#pragma unroll
        for (int r = 0; r < 6; r++) { el = lane == r ? ferr[r] : el; er = lane == r ? ferr[6 + r] : er; }
        if (lane < 6) {
            if (io.lf_err) io.lf_err[(size_t)lane * B + i] = el;
            if (io.rf_err) io.rf_err[(size_t)lane * B + i] = er;
        }
    }
    if (isvar) { d.ox[(size_t)j * B + i] = x; d.omask[(size_t)j * B + i] = mask; }
    if (isrow) { d.oy[(size_t)lane * B + i] = yt; d.oz[(size_t)lane * B + i] = zt; }
    if (isvel) { d.oy[(size_t)(NC + lane) * B + i] = yv; d.oz[(size_t)(NC + lane) * B + i] = zv; }
    if (lane == 0) {
        d.orho[i] = rho; d.oinit[i] = 1; d.osetups[i] = setups;
        if (d.oflags) d.oflags[i] = flags;
        if (io.status) io.status[i] = status;
        if (io.nwsr) io.nwsr[i] = info_iter; // OSQP's info->iter takes the place of qpOASES' nWSR
    }
}

template <int NJ, bool COMC>
cudaError_t launch_osqp_t(const IkDev& d, const IkTickIO& io, int mode, cudaStream_t s)
{
    using LY = IkOsqpLayout<NJ, COMC>;
    const size_t smem = (size_t)LY::TOTAL * sizeof(double) * kOsqpWarps;
    cudaError_t e = cudaFuncSetAttribute(ik_osqp_kernel<NJ, COMC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ik_osqp_kernel<NJ, COMC><<<(d.B + kOsqpWarps - 1) / kOsqpWarps, kOsqpWarps * 32, smem, s>>>(d, io, mode);
    return cudaGetLastError();
}
