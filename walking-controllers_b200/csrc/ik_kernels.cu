// ik_kernels.cu — see ik_kernels.cuh for the mapping to the reference.
#include "ik_kernels.cuh"
#include "ik_math.cuh"

#include <cfloat>
#include <cmath>

namespace wcb {

namespace {

constexpr double kQpInf = 1e20;      // qpOASES INFTY
constexpr double kViolTol = 1e-10;
constexpr double kRankTol = 1e-11;

constexpr int kWarpsPerBlock = 4;
// tuning switches (build variants, tools/sweep_ik_variants.sh): pointer-jumping FK, bounded sub-tree loop
#ifndef WC_IK_JUMP
#define WC_IK_JUMP 1
#endif
#ifndef WC_IK_SUBLOOP
#define WC_IK_SUBLOOP 1
#endif
#ifndef WC_IK_MODEL_SMEM
#define WC_IK_MODEL_SMEM 0 // fast kernel: 1 = model table staged in shared memory per CTA; 0 = read through L1 (measured 0.741 -> 0.731 ms)
#endif
#ifndef WC_IK_FAST_MINB
#define WC_IK_FAST_MINB 4 // resident CTAs per SM the fast kernel is compiled for (register cap 128)
#endif

// ---- K3: floating-base forward kinematics + the Jacobian column of this lane's QP variable ---------
// WalkingFK::{evaluateWorldToBaseTransformation,setInternalRobotState} + the getters of
// WalkingForwardKinematics.cpp:278-340,411-454 (iDynTree KinDynComputations restated; MIXED representation).
// Lane l owns link l (and l + 32) for the transforms, then QP variable j = lane for the columns.
// WL [links][12] world_H_link and MC [links][4] mass-weighted link CoMs stay in shared memory.
// JUMP: root_H_link by pointer jumping (ceil(log2(depth)) rounds in which EVERY link composes with its current
// ancestor) instead of one pass per tree level with a handful of active lanes: 3 rounds instead of 7 levels for
// the iCub tree.  The products associate differently (1e-16 relative), so the OSQP-based variant, whose life cycle
// is decided on exact zeros of the neck Jacobian, keeps the level order of the CPU restatement (JUMP = false).
template <int NV, bool JUMP = false>
__device__ __forceinline__ void fk_stage(const IkModelDev& M, const IkTickIO& io, int i, size_t B, int lane,
                                         double* WL, double* MC, double* colLf, double* colRf, double* colN,
                                         double* colC, double* lfH, double* rfH, double* nR, double* com, double* pBo,
                                         const double* base_H_s = nullptr)
{
    const int j = lane;
    const int left = io.left_is_fixed[i] ? 1 : 0;
    const int L = M.n_links;
    const int l = lane; // link of this lane (first pass), links >= 32 in a second pass
    for (int base = 0; base < L; base += 32) {
        const int lk = base + l;
        if (lk < L) {
            double R[9], p[3];
#pragma unroll
            for (int k = 0; k < 9; k++) R[k] = M.joint_R[lk][k];
#pragma unroll
            for (int k = 0; k < 3; k++) p[k] = M.joint_p[lk][k];
            const int dj = M.dof[lk];
            if (dj >= 0) {
                const double q = io.q_lin[(size_t)dj * B + i];
                double s, c;
                sincos(q, &s, &c);
                const double v = 1.0 - c, x = M.axis[lk][0], y = M.axis[lk][1], z = M.axis[lk][2];
                const double Rq[9] = {c + x * x * v, x * y * v - z * s, x * z * v + y * s,
                                      y * x * v + z * s, c + y * y * v, y * z * v - x * s,
                                      z * x * v - y * s, z * y * v + x * s, c + z * z * v};
                double Rn[9];
                mat3_mul(R, Rq, Rn);
#pragma unroll
                for (int k = 0; k < 9; k++) R[k] = Rn[k];
            }
            // local transform parked in WL; composed level by level below
#pragma unroll
            for (int k = 0; k < 9; k++) WL[lk * 12 + k] = R[k];
#pragma unroll
            for (int k = 0; k < 3; k++) WL[lk * 12 + 9 + k] = p[k];
        }
    }
    __syncwarp();
    const bool jump = JUMP && WC_IK_JUMP && L <= 32 && M.n_jump > 0;
    if (jump) {
        // T[l] = anc_H_l; every round T[l] <- T[anc] T[l], anc <- anc(anc) (table from the host), until anc = none
        double R[9], p[3];
        const int lk = l < L ? l : 0;
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = WL[lk * 12 + k];
#pragma unroll
        for (int k = 0; k < 3; k++) p[k] = WL[lk * 12 + 9 + k];
        for (int r = 0; r < M.n_jump; r++) {
            const int an = l < L ? (int)M.jump[r][l] : -1;
            if (an >= 0) xf_mul(WL + an * 12, WL + an * 12 + 9, R, p, R, p);
            __syncwarp();
            if (an >= 0) {
#pragma unroll
                for (int k = 0; k < 9; k++) WL[lk * 12 + k] = R[k];
#pragma unroll
                for (int k = 0; k < 3; k++) WL[lk * 12 + 9 + k] = p[k];
            }
            __syncwarp();
        }
    }
    // root_H_link by tree depth (parent[l] < l, depth known on the host)
    for (int dep = 1; dep <= (jump ? 0 : M.max_depth); dep++) {
        for (int base = 0; base < L; base += 32) {
            const int lk = base + l;
            if (lk < L && M.depth[lk] == dep) {
                const int pa = M.parent[lk];
                double Ro[9], po[3];
                xf_mul(WL + pa * 12, WL + pa * 12 + 9, WL + lk * 12, WL + lk * 12 + 9, Ro, po);
#pragma unroll
                for (int k = 0; k < 9; k++) WL[lk * 12 + k] = Ro[k];
#pragma unroll
                for (int k = 0; k < 3; k++) WL[lk * 12 + 9 + k] = po[k];
            }
        }
        __syncwarp();
    }
    // link 0 holds its "joint" transform (identity for a proper root); world_H_root from the stance sole
    const int bl = M.base_link[left];
    double wR[9], wp[3];
    {
        const int fidx = left ? 0 : 1;
        double Rs[9], ps[3];
        xf_mul(WL + bl * 12, WL + bl * 12 + 9, M.frame_R[fidx], M.frame_p[fidx], Rs, ps); // root_H_sole
        // inverse
        double Ri[9], pi[3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) Ri[3 * r + c] = Rs[3 * c + r];
#pragma unroll
        for (int r = 0; r < 3; r++) pi[r] = -(Ri[3 * r] * ps[0] + Ri[3 * r + 1] * ps[1] + Ri[3 * r + 2] * ps[2]);
        double bR[9], bp[3];
        // stance-sole transform: from HBM, or from the shared-memory copy the caller prefetched
        // (io.base_H == NULL: the desired transform of the stance foot, WalkingFK::evaluateWorldToBaseTransformation,
        // WalkingForwardKinematics.cpp:217-244 -- the caller then does not upload it a second time)
        const double* bH = base_H_s ? base_H_s : (io.base_H ? io.base_H : (left ? io.lf_H_des : io.rf_H_des)) + i;
        const size_t bs = base_H_s ? (size_t)1 : B;
        bR[0] = bH[0 * bs]; bR[1] = bH[1 * bs]; bR[2] = bH[2 * bs]; bp[0] = bH[3 * bs];
        bR[3] = bH[4 * bs]; bR[4] = bH[5 * bs]; bR[5] = bH[6 * bs]; bp[1] = bH[7 * bs];
        bR[6] = bH[8 * bs]; bR[7] = bH[9 * bs]; bR[8] = bH[10 * bs]; bp[2] = bH[11 * bs];
        xf_mul(bR, bp, Ri, pi, wR, wp);
    }
    __syncwarp();
    for (int base = 0; base < L; base += 32) {
        const int lk = base + l;
        if (lk < L) {
            double Ro[9], po[3];
            xf_mul(wR, wp, WL + lk * 12, WL + lk * 12 + 9, Ro, po);
#pragma unroll
            for (int k = 0; k < 9; k++) WL[lk * 12 + k] = Ro[k];
#pragma unroll
            for (int k = 0; k < 3; k++) WL[lk * 12 + 9 + k] = po[k];
            const double m = M.mass[lk];
#pragma unroll
            for (int r = 0; r < 3; r++)
                MC[lk * 4 + r] = m * (Ro[3 * r] * M.com[lk][0] + Ro[3 * r + 1] * M.com[lk][1] + Ro[3 * r + 2] * M.com[lk][2] + po[r]);
        }
    }
    __syncwarp();
    // task frames (replicated) and total first moment
    {
        double R[9], p[3];
        xf_mul(WL + M.frame_link[0] * 12, WL + M.frame_link[0] * 12 + 9, M.frame_R[0], M.frame_p[0], R, p);
        lfH[0] = R[0]; lfH[1] = R[1]; lfH[2] = R[2]; lfH[3] = p[0]; lfH[4] = R[3]; lfH[5] = R[4]; lfH[6] = R[5]; lfH[7] = p[1];
        lfH[8] = R[6]; lfH[9] = R[7]; lfH[10] = R[8]; lfH[11] = p[2];
        xf_mul(WL + M.frame_link[1] * 12, WL + M.frame_link[1] * 12 + 9, M.frame_R[1], M.frame_p[1], R, p);
        rfH[0] = R[0]; rfH[1] = R[1]; rfH[2] = R[2]; rfH[3] = p[0]; rfH[4] = R[3]; rfH[5] = R[4]; rfH[6] = R[5]; rfH[7] = p[1];
        rfH[8] = R[6]; rfH[9] = R[7]; rfH[10] = R[8]; rfH[11] = p[2];
        mat3_mul(WL + M.frame_link[2] * 12, M.frame_R[2], nR);
    }
    // total first moment: lane partials + butterfly (every lane ends with the same bits); the serial loop over all
    // links on every lane was 6 % of the fused tick's samples (profiles/r1o_hot_lines.md)
    double Mc[3] = {0, 0, 0};
    for (int k = lane; k < L; k += 32) { Mc[0] += MC[k * 4]; Mc[1] += MC[k * 4 + 1]; Mc[2] += MC[k * 4 + 2]; }
#pragma unroll
    for (int r = 0; r < 3; r++) Mc[r] = warp_sum(Mc[r]);
    const double Mtot = M.total_mass;
#pragma unroll
    for (int r = 0; r < 3; r++) com[r] = Mc[r] / Mtot;
    const double pB[3] = {WL[bl * 12 + 9], WL[bl * 12 + 10], WL[bl * 12 + 11]};
    const double pL[3] = {lfH[3], lfH[7], lfH[11]}, pR[3] = {rfH[3], rfH[7], rfH[11]};
    // ---- Jacobian column of this lane (MIXED representation, WalkingForwardKinematics.cpp:31,436-454)
    if (j < 3) {
        colLf[j] = 1.0; colRf[j] = 1.0; colC[j] = 1.0;
    } else if (j < 6) {
        const int k = j - 3;
        double e[3] = {0, 0, 0}; e[k] = 1.0;
        double r[3], o[3];
        r[0] = pL[0] - pB[0]; r[1] = pL[1] - pB[1]; r[2] = pL[2] - pB[2];
        cross3(e, r, o); colLf[0] = o[0]; colLf[1] = o[1]; colLf[2] = o[2]; colLf[3 + k] = 1.0;
        r[0] = pR[0] - pB[0]; r[1] = pR[1] - pB[1]; r[2] = pR[2] - pB[2];
        cross3(e, r, o); colRf[0] = o[0]; colRf[1] = o[1]; colRf[2] = o[2]; colRf[3 + k] = 1.0;
        colN[k] = 1.0;
        r[0] = com[0] - pB[0]; r[1] = com[1] - pB[1]; r[2] = com[2] - pB[2];
        cross3(e, r, o); colC[0] = o[0]; colC[1] = o[1]; colC[2] = o[2];
    } else if (j < NV) {
        const int dj = j - 6;
        const int c = M.link_of_dof[dj];
        const double* Rc = WL + c * 12;
        const double o[3] = {WL[c * 12 + 9], WL[c * 12 + 10], WL[c * 12 + 11]};
        double a[3];
#pragma unroll
        for (int r = 0; r < 3; r++) a[r] = Rc[3 * r] * M.axis[c][0] + Rc[3 * r + 1] * M.axis[c][1] + Rc[3 * r + 2] * M.axis[c][2];
        double r3[3], lin[3];
        const double sl = (double)M.s_frame[left][0][dj], sr = (double)M.s_frame[left][1][dj], sn = (double)M.s_frame[left][2][dj];
        r3[0] = pL[0] - o[0]; r3[1] = pL[1] - o[1]; r3[2] = pL[2] - o[2];
        cross3(a, r3, lin);
#pragma unroll
        for (int r = 0; r < 3; r++) { colLf[r] = sl * lin[r]; colLf[3 + r] = sl * a[r]; }
        r3[0] = pR[0] - o[0]; r3[1] = pR[1] - o[1]; r3[2] = pR[2] - o[2];
        cross3(a, r3, lin);
#pragma unroll
        for (int r = 0; r < 3; r++) { colRf[r] = sr * lin[r]; colRf[3 + r] = sr * a[r]; }
#pragma unroll
        for (int r = 0; r < 3; r++) colN[r] = sn * a[r];
        // CoM: first moment of the sub-tree hanging from this joint (or of its complement when the
        // base link sits inside that sub-tree)
        double S[3] = {0, 0, 0};
        const unsigned long long mask = M.sub_mask[c];
        // links of the sub-tree in ascending order, between its first and last member only (one contiguous run for
        // a depth-first numbered tree); the warp runs as many trips as its largest sub-tree needs
        const int k1 = 63 - __clzll((long long)mask);
        for (int k = WC_IK_SUBLOOP ? __ffsll((long long)mask) - 1 : 0; k <= (WC_IK_SUBLOOP ? k1 : L - 1); k++)
            if ((mask >> k) & 1ull) { S[0] += MC[k * 4]; S[1] += MC[k * 4 + 1]; S[2] += MC[k * 4 + 2]; }
        const double ms = M.sub_mass[c];
        double mom[3];
        if (!M.base_in_sub[left][dj]) {
#pragma unroll
            for (int r = 0; r < 3; r++) mom[r] = (S[r] - ms * o[r]) / Mtot;
        } else {
#pragma unroll
            for (int r = 0; r < 3; r++) mom[r] = -((Mc[r] - S[r]) - (Mtot - ms) * o[r]) / Mtot;
        }
        cross3(a, mom, lin);
#pragma unroll
        for (int r = 0; r < 3; r++) colC[r] = lin[r];
    }
    __syncwarp();
    pBo[0] = pB[0]; pBo[1] = pB[1]; pBo[2] = pB[2];
}

// ---- K3b: the vectors of the QP: equality right-hand side b (setBounds), gradient entry g_j of this
// lane's variable (setGradientVector), regularisation weight Lambda_jj (WalkingQPIK::initializeMatrices).
// VARIANT 1 = WalkingQPIK_osqp: the neck term of the gradient carries an extra k_attFoot factor (_osqp.cpp:183,193) and a
// foot whose desired twist has twist(0) == twist(1) == 0 is constrained to the bare twist (_osqp.cpp:286-306).
template <bool COMC, int VARIANT = 0>
__device__ __forceinline__ void qp_vectors(const IkDev& d, const IkTickIO& io, int i, size_t B, int ig, int lane,
                                           const double* lfH, const double* rfH, const double* nR,
                                           const double* colN, const double* colC, double* bvec, double& gj, double& lam)
{
    const wc_ik_config& cfg = d.cfg;
    const int j = lane;
    const bool isjoint = j >= 6 && j < d.nJ + 6;
    // io may point at a one-instance shared-memory copy (i = 0, B = 1); ig is the instance for the gain arrays
    const double kpos = d.kpos ? d.kpos[ig] : cfg.k_pos_foot;
    const double katt = d.katt ? d.katt[ig] : cfg.k_att_foot;
    // equality right-hand side (setBounds, _qpOASES.cpp:216-282)
    double Hd[12], tw[6], R[9], Rd[9], e[3];
#pragma unroll
    for (int f = 0; f < 2; f++) {
        const double* src = f == 0 ? io.lf_H_des : io.rf_H_des;
        const double* tws = f == 0 ? io.lf_twist : io.rf_twist;
        const double* Hc = f == 0 ? lfH : rfH;
#pragma unroll
        for (int k = 0; k < 12; k++) Hd[k] = src[(size_t)k * B + i];
#pragma unroll
        for (int k = 0; k < 6; k++) tw[k] = tws[(size_t)k * B + i];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) { R[3 * r + c] = Hc[4 * r + c]; Rd[3 * r + c] = Hd[4 * r + c]; }
        rot_error(R, Rd, e);
        const bool bare = VARIANT == 1 && tw[0] == tw[1] && tw[0] == 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            bvec[6 * f + r] = bare ? tw[r] : tw[r] - kpos * (Hc[4 * r + 3] - Hd[4 * r + 3]);
            bvec[6 * f + 3 + r] = bare ? tw[3 + r] : tw[3 + r] - katt * e[r];
        }
    }
    double cv[3];
#pragma unroll
    for (int r = 0; r < 3; r++) cv[r] = io.com_vel_des[(size_t)r * B + i];
    if (COMC) {
#pragma unroll
        for (int r = 0; r < 3; r++)
            bvec[12 + r] = cv[r] - cfg.k_pos_com * (io.com_meas[(size_t)r * B + i] - io.com_pos_des[(size_t)r * B + i]);
    } else {
        bvec[12] = bvec[13] = bvec[14] = 0.0;
    }
    // gradient (setGradientVector, :155-182)
    double Rdes[9], nd[9], en[3], We[3];
#pragma unroll
    for (int k = 0; k < 9; k++) nd[k] = io.neck_R_des[(size_t)k * B + i];
    mat3_mul(nd, cfg.additional_rotation, Rdes); // setDesiredNeckOrientation (QPIK.cpp:143-146)
    rot_error(nR, Rdes, en);
    // W * (-k_neck e_n): with colN it gives this lane's neck term; also returned for the base elimination
    const double kn = VARIANT == 1 ? katt : 1.0;
#pragma unroll
    for (int r = 0; r < 3; r++)
        We[r] = cfg.neck_weight[3 * r] * (kn * (-cfg.k_neck * en[0])) + cfg.neck_weight[3 * r + 1] * (kn * (-cfg.k_neck * en[1])) +
                cfg.neck_weight[3 * r + 2] * (kn * (-cfg.k_neck * en[2]));
    gj = -(colN[0] * We[0] + colN[1] * We[1] + colN[2] * We[2]);
    lam = 0.0;
    if (isjoint) {
        lam = cfg.joint_reg_weight[j - 6];
        gj -= lam * (cfg.joint_reg_gain[j - 6] * (cfg.joint_reg_rad[j - 6] - io.q_meas[(size_t)(j - 6) * B + i]));
    }
    if (!COMC) {
        double Wv[3];
#pragma unroll
        for (int r = 0; r < 3; r++) Wv[r] = cfg.com_weight[3 * r] * cv[0] + cfg.com_weight[3 * r + 1] * cv[1] + cfg.com_weight[3 * r + 2] * cv[2];
        gj -= colC[0] * Wv[0] + colC[1] * Wv[1] + colC[2] * Wv[2];
    }
}

// ---- Goldfarb-Idnani dual active set on the bounds, reduced space, warp-cooperative ---------------
// Same algorithm as oracle/qpik_ref.c step 5.  Lane c <-> component c of the reduced-space vectors,
// lane j <-> variable j for the primal point and the bound slacks; matrices live in shared memory
// (J by columns per lane for J'n, by rows per lane for J2 d2 and the Givens updates; the odd leading
// dimension keeps both access patterns conflict-free).  Called by the whole warp.
template <int NV, int NZ>
__device__ __noinline__ int gi_bounds(int lane, const double* ZS /*[NV][NZ+1]*/, const double* XP, const double* LB,
                                      const double* UB, const double* LS /*[NZ][NZ+1] chol*/, double* JS, double* RS,
                                      double* W, double* NP, double* DV, double* ZV, double* RV, double* U, int* act,
                                      unsigned char* is_act, int nwsr_max, int* nwsr_out)
{
    constexpr int LDZ = NZ + 1, LDL = NZ + 1;
    constexpr unsigned FULL = 0xffffffffu;
    int q = 0, nwsr = 0;
    // J = L^-T : lane c solves L' J[:,c] = e_c
    if (lane < NZ) {
        const int c = lane;
        for (int i = NZ - 1; i >= 0; i--) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = i + 1; k < NZ; k++) s -= LS[k * LDL + i] * JS[k * LDL + c];
            JS[i * LDL + c] = s / LS[i * LDL + i];
        }
    }
    for (int j = lane; j < 2 * NV; j += 32) is_act[j] = 0;
    __syncwarp();
    for (;;) {
        // most violated inactive bound: min over (slack, id), ties to the smaller id (the oracle's scan order)
        double best = INFINITY; int id = 0x7fffffff;
        if (lane < NV) {
            double xj = XP[lane];
            for (int c = 0; c < NZ; c++) xj += ZS[lane * LDZ + c] * W[c];
            if (LB[lane] > -kQpInf && !is_act[2 * lane]) { best = xj - LB[lane]; id = 2 * lane; }
            if (UB[lane] < kQpInf && !is_act[2 * lane + 1]) { const double s = UB[lane] - xj; if (s < best) { best = s; id = 2 * lane + 1; } }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, id, o);
            if (ob < best || (ob == best && oi < id)) { best = ob; id = oi; }
        }
        if (!(best < -kViolTol)) break; // optimal
        const int p = id, pv = p >> 1;
        const double sgn = (p & 1) ? -1.0 : 1.0;
        double sp = best;
        if (lane < NZ) NP[lane] = sgn * ZS[pv * LDZ + lane];
        __syncwarp();
        double up_ = 0.0;
        bool added = false;
        while (!added) {
            // d = J' np (lane c), z = J2 d2 (lane i), r = R^-1 d1 (lane 0)
            double dc = 0.0;
            if (lane < NZ) { for (int i = 0; i < NZ; i++) dc += JS[i * LDL + lane] * NP[i]; DV[lane] = dc; }
            __syncwarp();
            double zi = 0.0;
            if (lane < NZ) { for (int c = q; c < NZ; c++) zi += JS[lane * LDL + c] * DV[c]; ZV[lane] = zi; }
            if (lane == 0)
                for (int i = q - 1; i >= 0; i--) { double s = DV[i]; for (int k = i + 1; k < q; k++) s -= RS[i * LDL + k] * RV[k]; RV[i] = s / RS[i * LDL + i]; }
            __syncwarp();
            double zn = (lane < NZ) ? zi * NP[lane] : 0.0, dd = (lane < NZ) ? dc * dc : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { zn += __shfl_xor_sync(FULL, zn, o); dd += __shfl_xor_sync(FULL, dd, o); }
            const bool have_z = zn > 1e-12 * dd && zn > 0.0;
            const double t2 = have_z ? -sp / zn : INFINITY;
            double t1 = INFINITY; int l = 0x7fffffff;
            if (lane < q && RV[lane] > 0.0) { t1 = U[lane] / RV[lane]; l = lane; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ot = __shfl_xor_sync(FULL, t1, o);
                const int ol = __shfl_xor_sync(FULL, l, o);
                if (ot < t1 || (ot == t1 && ol < l)) { t1 = ot; l = ol; }
            }
            const double tt = t1 < t2 ? t1 : t2;
            if (!(tt < INFINITY)) { *nwsr_out = nwsr; return WC_QP_INFEASIBLE; }
            if (have_z && lane < NZ) W[lane] += tt * zi;
            if (lane < q) U[lane] -= tt * RV[lane];
            up_ += tt;
            __syncwarp();
            if (have_z && t2 <= t1) {
                // full step: add p.  Givens chain on d[q..NZ-1] (sequential, carried in a register),
                // each rotation applied to two columns of J by the lanes (rows)
                double carry = DV[NZ - 1];
                for (int j = NZ - 1; j > q; j--) {
                    const double a0 = DV[j - 1], a1 = carry;
                    const double hh = hypot(a0, a1);
                    if (hh == 0.0) { carry = 0.0; continue; }
                    const double cs = a0 / hh, sn = a1 / hh;
                    carry = hh;
                    if (lane < NZ) {
                        const double j0 = JS[lane * LDL + j - 1], j1 = JS[lane * LDL + j];
                        JS[lane * LDL + j - 1] = cs * j0 + sn * j1;
                        JS[lane * LDL + j] = -sn * j0 + cs * j1;
                    }
                }
                if (lane <= q) RS[lane * LDL + q] = (lane < q) ? DV[lane] : carry;
                if (lane == 0) { act[q] = p; U[q] = up_; is_act[p] = 1; }
                q++;
                added = true;
                __syncwarp();
            } else {
                // drop constraint l: delete column l of R, re-triangularise with Givens on rows
                if (lane == 0) {
                    is_act[act[l]] = 0;
                    for (int k = l; k < q - 1; k++) { act[k] = act[k + 1]; U[k] = U[k + 1]; }
                }
                if (lane < NZ)
                    for (int k = l; k < q - 1; k++) if (lane <= k + 1) RS[lane * LDL + k] = RS[lane * LDL + k + 1];
                q--;
                __syncwarp();
                for (int k = l; k < q; k++) {
                    const double a0 = RS[k * LDL + k], a1 = RS[(k + 1) * LDL + k];
                    const double hh = hypot(a0, a1);
                    __syncwarp();
                    if (hh != 0.0) {
                        const double cs = a0 / hh, sn = a1 / hh;
                        if (lane == k) { RS[k * LDL + k] = hh; RS[(k + 1) * LDL + k] = 0.0; }
                        if (lane > k && lane < q) {
                            const double r0 = RS[k * LDL + lane], r1 = RS[(k + 1) * LDL + lane];
                            RS[k * LDL + lane] = cs * r0 + sn * r1;
                            RS[(k + 1) * LDL + lane] = -sn * r0 + cs * r1;
                        }
                        if (lane < NZ) {
                            const double j0 = JS[lane * LDL + k], j1 = JS[lane * LDL + k + 1];
                            JS[lane * LDL + k] = cs * j0 + sn * j1;
                            JS[lane * LDL + k + 1] = -sn * j0 + cs * j1;
                        }
                    }
                    __syncwarp();
                }
                double xj = XP[pv];
                for (int c = 0; c < NZ; c++) xj += ZS[pv * LDZ + c] * W[c];
                sp = (p & 1) ? UB[pv] - xj : xj - LB[pv];
            }
            nwsr++;
            if (nwsr > nwsr_max) { *nwsr_out = nwsr; return WC_QP_MAX_NWSR_REACHED; }
        }
    }
    *nwsr_out = nwsr;
    return WC_QP_SUCCESSFUL_RETURN;
}

template <int NJ, bool COMC>
struct IkLayout {
    static constexpr int NV = NJ + 6;
    static constexpr int NC = COMC ? 15 : 12;
    static constexpr int NZ = NV - NC;
    static constexpr int LDA = 33;
    static constexpr int LDZ = NZ + 1;
    static constexpr int LDL = NZ + 1;
    // region 1: kinematics scratch, later Z and the Cholesky factor
    static constexpr int R1_KIN = WC_MAX_LINKS * 12 + WC_MAX_LINKS * 4;
    static constexpr int R1_QP = 32 * LDZ + NZ * LDL;
    static constexpr int R1 = R1_KIN > R1_QP ? R1_KIN : R1_QP;
    static constexpr int O_WL = 0, O_MC = WC_MAX_LINKS * 12;
    static constexpr int O_ZS = 0, O_LS = 32 * LDZ;
    static constexpr int O_AS = R1;
    static constexpr int O_JN = O_AS + 15 * LDA;
    static constexpr int O_JC = O_JN + 3 * 32;
    static constexpr int O_JS = O_JC + 3 * 32;
    static constexpr int O_RS = O_JS + NZ * LDL;
    static constexpr int O_XP = O_RS + NZ * LDL;
    static constexpr int O_LB = O_XP + 32, O_UB = O_LB + 32, O_TS = O_UB + 32, O_LAM = O_TS + 32;
    static constexpr int O_W = O_LAM + 32, O_HR = O_W + 32, O_NP = O_HR + 32, O_DV = O_NP + 32, O_ZV = O_DV + 32;
    static constexpr int O_RV = O_ZV + 32, O_U = O_RV + 32, O_P = O_U + 32, O_WP = O_P + 3 * 32;
    static constexpr int O_BV = O_WP + 3 * 32;  // equality right-hand side b[16]
    static constexpr int O_VS = O_BV + 16;      // Householder vector broadcast, double-buffered 2 x 32
    static constexpr int O_INT = O_VS + 64;     // act[32] ints + is_act[64] bytes = 24 doubles
    static constexpr int TOTAL = ((O_INT + 24 + 15) / 16) * 16 + 8;
};

template <int NJ, bool COMC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) ik_kernel(IkDev d, IkTickIO io, int mode, int from_list)
{
    using LY = IkLayout<NJ, COMC>;
    constexpr int NV = LY::NV, NC = LY::NC, NZ = LY::NZ, LDA = LY::LDA, LDZ = LY::LDZ, LDL = LY::LDL;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long inst = (long long)blockIdx.x * kWarpsPerBlock + warp;
    // from_list: the instances the fast path handed over (active velocity bounds, near-singular tasks)
    if (inst >= (from_list ? (long long)*d.slow_count : (long long)d.B)) return; // whole warp exits together
    const int i = from_list ? d.slow_list[inst] : (int)inst;
    const size_t B = (size_t)d.B;
    double* sm = smem_all + (size_t)warp * LY::TOTAL;
    double* WL = sm + LY::O_WL; double* MC = sm + LY::O_MC;
    double* ZS = sm + LY::O_ZS; double* LS = sm + LY::O_LS;
    double* AS = sm + LY::O_AS; double* JN = sm + LY::O_JN; double* JC = sm + LY::O_JC;
    double* JS = sm + LY::O_JS; double* RS = sm + LY::O_RS;
    double* XP = sm + LY::O_XP; double* LBs = sm + LY::O_LB; double* UBs = sm + LY::O_UB;
    double* TS = sm + LY::O_TS; double* LAM = sm + LY::O_LAM;
    double* Ws = sm + LY::O_W; double* HR = sm + LY::O_HR; double* NPs = sm + LY::O_NP; double* DVs = sm + LY::O_DV;
    double* ZVs = sm + LY::O_ZV; double* RVs = sm + LY::O_RV; double* Us = sm + LY::O_U;
    double* Ps = sm + LY::O_P; double* WPs = sm + LY::O_WP;
    double* BV = sm + LY::O_BV; double* VS = sm + LY::O_VS;
    int* acts = reinterpret_cast<int*>(sm + LY::O_INT);
    unsigned char* is_act = reinterpret_cast<unsigned char*>(sm + LY::O_INT + 16);

    const IkModelDev& M = *d.model;
    const wc_ik_config& cfg = d.cfg;
    const int j = lane;            // QP variable of this lane
    const bool isvar = j < NV;
    const bool isjoint = j >= 6 && j < NV;

    // this lane's Jacobian columns and the replicated frame data
    double colLf[6] = {0, 0, 0, 0, 0, 0}, colRf[6] = {0, 0, 0, 0, 0, 0}, colN[3] = {0, 0, 0}, colC[3] = {0, 0, 0};
    double lfH[12], rfH[12], nR[9], com[3], pBv[3] = {0, 0, 0};

    if (mode != IK_MODE_EXT_KIN) {
        fk_stage<NV, true>(M, io, i, B, lane, WL, MC, colLf, colRf, colN, colC, lfH, rfH, nR, com, pBv);
        if (mode == IK_MODE_FK_ONLY) {
            if (lane < 12) {
                if (io.o_lf_H) io.o_lf_H[(size_t)lane * B + i] = lfH[lane];
                if (io.o_rf_H) io.o_rf_H[(size_t)lane * B + i] = rfH[lane];
            }
            if (lane < 9 && io.o_neck_R) io.o_neck_R[(size_t)lane * B + i] = nR[lane];
            if (lane < 3 && io.o_com) io.o_com[(size_t)lane * B + i] = com[lane];
            if (isvar) {
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    if (io.o_J_lf) io.o_J_lf[(size_t)(r * NV + j) * B + i] = colLf[r];
                    if (io.o_J_rf) io.o_J_rf[(size_t)(r * NV + j) * B + i] = colRf[r];
                }
                if (io.o_J_neck) {
                    // linear rows of the neck Jacobian are not used by the QP; emit them for the WalkingFK getter
                    double lin[3] = {0, 0, 0};
                    const double pN[3] = {0, 0, 0};
                    (void)pN;
                    // recompute: point = neck frame origin
                    double Rn[9], pn[3];
                    xf_mul(WL + M.frame_link[2] * 12, WL + M.frame_link[2] * 12 + 9, M.frame_R[2], M.frame_p[2], Rn, pn);
                    if (j < 3) lin[j] = 1.0;
                    else if (j < 6) {
                        double e[3] = {0, 0, 0}; e[j - 3] = 1.0;
                        const double r[3] = {pn[0] - pBv[0], pn[1] - pBv[1], pn[2] - pBv[2]};
                        cross3(e, r, lin);
                    } else {
                        const int c = M.link_of_dof[j - 6];
                        const double r[3] = {pn[0] - WL[c * 12 + 9], pn[1] - WL[c * 12 + 10], pn[2] - WL[c * 12 + 11]};
                        cross3(colN, r, lin); // colN already carries the path sign
                    }
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        io.o_J_neck[(size_t)(r * NV + j) * B + i] = lin[r];
                        io.o_J_neck[(size_t)((3 + r) * NV + j) * B + i] = colN[r];
                    }
                }
                if (io.o_J_com)
#pragma unroll
                    for (int r = 0; r < 3; r++) io.o_J_com[(size_t)(r * NV + j) * B + i] = colC[r];
            }
            return;
        }
    } else {
        // caller-supplied kinematics (WalkingQPIK drop-in: set*Jacobian + setRobotState)
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
                colN[r] = io.J_neck[(size_t)((3 + r) * NV + j) * B + i]; // setNeckJacobian keeps rows 3..5
                colC[r] = io.J_com[(size_t)(r * NV + j) * B + i];
            }
        }
        com[0] = com[1] = com[2] = 0.0;
    }

    // ===================================================================== K3b: QP data (no dense H)
    double gj, lam;
    {
        double bvec[15];
        qp_vectors<COMC>(d, io, i, B, i, lane, lfH, rfH, nR, colN, colC, bvec, gj, lam);
        TS[lane] = isvar ? gj : 0.0; // g parked in TS
        if (lane < 15) {
#pragma unroll
            for (int k = 0; k < 15; k++) if (lane == k) BV[k] = bvec[k];
        }
        LAM[lane] = lam;
        LBs[lane] = isjoint ? cfg.min_vel[j - 6] : -DBL_MAX; // setVelocityBounds (:23-51)
        UBs[lane] = isjoint ? cfg.max_vel[j - 6] : DBL_MAX;
    }
    // constraint matrix rows to shared memory (setLinearConstraintMatrix, :184-214): A[r][j]
#pragma unroll
    for (int r = 0; r < 6; r++) { AS[r * LDA + lane] = isvar ? colLf[r] : 0.0; AS[(6 + r) * LDA + lane] = isvar ? colRf[r] : 0.0; }
#pragma unroll
    for (int r = 0; r < 3; r++) {
        if (COMC) AS[(12 + r) * LDA + lane] = isvar ? colC[r] : 0.0;
        JN[r * 32 + lane] = isvar ? colN[r] : 0.0;
        JC[r * 32 + lane] = isvar ? colC[r] : 0.0;
    }
    __syncwarp();

    // ===================================================================== K4: exact QP
    // ---- Householder QR of A' with one column per lane: T = column `lane` of A' (lanes < NC),
    //      G = column `lane` of Q' (starts as the identity); reflectors are broadcast by shuffles.
    double T[NV], G[NV];
#pragma unroll
    for (int r = 0; r < NV; r++) {
        T[r] = (lane < NC) ? AS[lane * LDA + r] : 0.0;
        G[r] = (r == lane) ? 1.0 : 0.0;
    }
    double amax = 0.0;
#pragma unroll
    for (int r = 0; r < NV; r++) amax = fmax(amax, fabs(T[r]));
    amax = warp_max(amax);
    int status = WC_QP_SUCCESSFUL_RETURN;
    bool rank_def = false;
#pragma unroll 1
    for (int k = 0; k < NC; k++) { // rolled on purpose: the body is ~29 x 8 instructions, 15 copies thrash the I-cache
        // Reflector of column k, built by the whole warp: the owner publishes its column (T is register
        // resident and k is a run-time index, so one predicated store per row is far cheaper than selecting
        // T[k..] out of the registers), lane r takes row r, the norm is a warp reduction.
        double* vs = VS + (k & 1) * 32;
        if (lane == k) {
#pragma unroll
            for (int r = 0; r < NV; r++) vs[r] = T[r];
        }
        __syncwarp();
        const double xr = (lane >= k && lane < NV) ? vs[lane] : 0.0;
        const double sig = warp_sum(xr * xr);
        const double tk = __shfl_sync(FULL, xr, k);
        const double nrm_b = sqrt(sig);
        if (nrm_b <= kRankTol * (amax > 1.0 ? amax : 1.0)) { rank_def = true; break; }
        const double alpha = tk > 0 ? -nrm_b : nrm_b;
        const double beta = 1.0 / (sig - tk * alpha);
        __syncwarp();
        vs[lane] = (lane == k) ? tk - alpha : xr;
        __syncwarp();
        double tauT = 0.0, tauG = 0.0;
#pragma unroll
        for (int r = 0; r < NV; r++) {
            const double v = vs[r];
            tauT += v * T[r];
            tauG += v * G[r];
        }
        tauT *= beta; tauG *= beta;
        // the owner's own column is reflected too: it becomes (.., alpha, ~0, ..) up to rounding; the entries
        // below the diagonal are never read again
        if (!(lane >= k && lane < NC)) tauT = 0.0;
#pragma unroll
        for (int r = 0; r < NV; r++) {
            const double v = vs[r];
            T[r] -= tauT * v;
            G[r] -= tauG * v;
        }
    }
    double nu = 0.0;
    int nwsr = 0;
    if (rank_def) {
        status = WC_QP_INFEASIBLE; // linearly dependent equality rows
    } else {
        // ---- R' y = b   (lane c holds column c of R)
        double s = (lane < NC) ? BV[lane] : 0.0;
        double tdiag = 1.0;
#pragma unroll
        for (int k = 0; k < NC; k++) tdiag = (lane == k) ? T[k] : tdiag;
        const double tdinv = 1.0 / tdiag; // one division per lane instead of one per substitution step
        double y[NC];
#pragma unroll
        for (int k = 0; k < NC; k++) {
            y[k] = __shfl_sync(FULL, s * tdinv, k);
            if (lane > k) s -= T[k] * y[k];
        }
        // particular solution and null-space basis row of this lane
        double xp = 0.0;
#pragma unroll
        for (int k = 0; k < NC; k++) xp += G[k] * y[k];
        double Zr[NZ];
#pragma unroll
        for (int c = 0; c < NZ; c++) Zr[c] = G[NC + c];
        if (!isvar) { xp = 0.0;
#pragma unroll
            for (int c = 0; c < NZ; c++) Zr[c] = 0.0; }
        // ---- t = H xp + g with H = Lambda + Jn' W Jn (+ Jc' Wc Jc)
        double u3[3], wu[3];
#pragma unroll
        for (int r = 0; r < 3; r++) u3[r] = warp_sum(JN[r * 32 + lane] * xp);
#pragma unroll
        for (int r = 0; r < 3; r++) wu[r] = cfg.neck_weight[3 * r] * u3[0] + cfg.neck_weight[3 * r + 1] * u3[1] + cfg.neck_weight[3 * r + 2] * u3[2];
        double tj = LAM[lane] * xp + JN[lane] * wu[0] + JN[32 + lane] * wu[1] + JN[64 + lane] * wu[2] + TS[lane];
        if (!COMC) {
            double uc[3], wc3[3];
#pragma unroll
            for (int r = 0; r < 3; r++) uc[r] = warp_sum(JC[r * 32 + lane] * xp);
#pragma unroll
            for (int r = 0; r < 3; r++) wc3[r] = cfg.com_weight[3 * r] * uc[0] + cfg.com_weight[3 * r + 1] * uc[1] + cfg.com_weight[3 * r + 2] * uc[2];
            tj += JC[lane] * wc3[0] + JC[32 + lane] * wc3[1] + JC[64 + lane] * wc3[2];
        }
        __syncwarp(); // kinematics scratch (WL/MC) is dead: region 1 becomes ZS/LS
#pragma unroll
        for (int c = 0; c < NZ; c++) ZS[lane * LDZ + c] = Zr[c];
        ZS[lane * LDZ + NZ] = 0.0;
        XP[lane] = xp;
        TS[lane] = isvar ? tj : 0.0;
        __syncwarp();
        // ---- reduced Hessian row `lane` (lanes < NZ): M = Z' Lambda Z + (Jn Z)' W (Jn Z) [+ CoM term]
        double Mr[NZ];
#pragma unroll
        for (int b = 0; b < NZ; b++) Mr[b] = 0.0;
        double hr = 0.0;
        const int a = lane < NZ ? lane : 0;
        {
            double P[3] = {0, 0, 0}, PC[3] = {0, 0, 0};
            for (int jj = 0; jj < NV; jj++) {
                const double za = ZS[jj * LDZ + a];
                const double zl = za * LAM[jj];
#pragma unroll
                for (int b = 0; b < NZ; b++) Mr[b] += zl * ZS[jj * LDZ + b];
                hr += za * TS[jj];
#pragma unroll
                for (int r = 0; r < 3; r++) P[r] += JN[r * 32 + jj] * za;
                if (!COMC) {
#pragma unroll
                    for (int r = 0; r < 3; r++) PC[r] += JC[r * 32 + jj] * za;
                }
            }
            double WP[3];
#pragma unroll
            for (int r = 0; r < 3; r++) WP[r] = cfg.neck_weight[3 * r] * P[0] + cfg.neck_weight[3 * r + 1] * P[1] + cfg.neck_weight[3 * r + 2] * P[2];
            if (lane < NZ) {
#pragma unroll
                for (int r = 0; r < 3; r++) { Ps[r * 32 + lane] = P[r]; WPs[r * 32 + lane] = WP[r]; }
            }
            __syncwarp();
#pragma unroll
            for (int b = 0; b < NZ; b++)
                Mr[b] += P[0] * WPs[b] + P[1] * WPs[32 + b] + P[2] * WPs[64 + b];
            if (!COMC) {
                __syncwarp();
                double WC3[3];
#pragma unroll
                for (int r = 0; r < 3; r++) WC3[r] = cfg.com_weight[3 * r] * PC[0] + cfg.com_weight[3 * r + 1] * PC[1] + cfg.com_weight[3 * r + 2] * PC[2];
                if (lane < NZ) {
#pragma unroll
                    for (int r = 0; r < 3; r++) WPs[r * 32 + lane] = WC3[r];
                }
                __syncwarp();
#pragma unroll
                for (int b = 0; b < NZ; b++)
                    Mr[b] += PC[0] * WPs[b] + PC[1] * WPs[32 + b] + PC[2] * WPs[64 + b];
            }
        }
        // ---- Cholesky M = L L' (row per lane, right-looking, pivots broadcast)
        double mmax = (lane < NZ) ? 0.0 : 0.0;
#pragma unroll
        for (int b = 0; b < NZ; b++) if (lane == b) mmax = fabs(Mr[b]);
        mmax = warp_max(lane < NZ ? mmax : 0.0);
        bool not_pd = false;
        double ldinv = 1.0; // 1 / L[lane][lane]
#pragma unroll
        for (int jc = 0; jc < NZ; jc++) {
            const double djj = __shfl_sync(FULL, Mr[jc], jc);
            if (!(djj > 1e-13 * (mmax > 1.0 ? mmax : 1.0))) { not_pd = true; break; }
            const double piv = sqrt(djj);
            const double pinv = 1.0 / piv;
            const double lij = (lane == jc) ? piv : Mr[jc] * pinv;
            if (lane == jc) ldinv = pinv;
            if (lane >= jc) Mr[jc] = lij;
#pragma unroll
            for (int k = jc + 1; k < NZ; k++) {
                const double lkj = __shfl_sync(FULL, Mr[jc], k);
                if (lane >= k) Mr[k] -= lij * lkj;
            }
        }
        if (not_pd) {
            status = WC_QP_NOT_CONVEX;
        } else {
            // L to shared memory (needed transposed below and by the active-set path)
            if (lane < NZ) {
#pragma unroll
                for (int b = 0; b < NZ; b++) LS[lane * LDL + b] = (b <= lane) ? Mr[b] : 0.0;
            }
            __syncwarp();
            // w = -M^-1 hr
            double acc = -hr;
            double wv[NZ];
#pragma unroll
            for (int k = 0; k < NZ; k++) {
                const double sk = __shfl_sync(FULL, acc * ldinv, k);
                if (lane > k) acc -= Mr[k] * sk;
                if (lane == k) acc = sk;
            }
#pragma unroll
            for (int k = NZ - 1; k >= 0; k--) {
                wv[k] = __shfl_sync(FULL, acc * ldinv, k);
                if (lane < k) acc -= LS[k * LDL + (lane < NZ ? lane : 0)] * wv[k];
            }
            nu = xp;
#pragma unroll
            for (int c = 0; c < NZ; c++) nu += Zr[c] * wv[c];
            // ---- velocity bounds: equality-constrained optimum feasible -> done (nWSR = 0)
            const bool viol = isjoint && ((LBs[lane] > -kQpInf && nu < LBs[lane] - kViolTol) ||
                                          (UBs[lane] < kQpInf && nu > UBs[lane] + kViolTol));
            if (__any_sync(FULL, viol)) {
#pragma unroll
                for (int c = 0; c < NZ; c++) if (lane == c) Ws[c] = wv[c];
                __syncwarp();
                int nw = 0;
                const int st = gi_bounds<NV, NZ>(lane, ZS, XP, LBs, UBs, LS, JS, RS, Ws, NPs, DVs, ZVs, RVs, Us, acts, is_act,
                                                 cfg.nwsr, &nw);
                __syncwarp();
                status = st; nwsr = nw;
                nu = xp;
#pragma unroll
                for (int c = 0; c < NZ; c++) nu += Zr[c] * Ws[c];
            }
        }
    }
    const bool ok = status == WC_QP_SUCCESSFUL_RETURN;
    if (!ok) nu = 0.0;
    if (!isvar) nu = 0.0;
    // ---- outputs: getSolution (:341-362), get{Left,Right}FootError (:364-401)
    if (isjoint && io.dq) io.dq[(size_t)(j - 6) * B + i] = nu;
    if (io.lf_err || io.rf_err) {
        double el[6], er[6];
#pragma unroll
        for (int r = 0; r < 6; r++) {
            el[r] = BV[r] - warp_sum(AS[r * LDA + lane] * nu);
            er[r] = BV[6 + r] - warp_sum(AS[(6 + r) * LDA + lane] * nu);
        }
#pragma unroll
        for (int r = 0; r < 6; r++) {
            if (lane == r && io.lf_err) io.lf_err[(size_t)r * B + i] = ok ? el[r] : 0.0;
            if (lane == r && io.rf_err) io.rf_err[(size_t)r * B + i] = ok ? er[r] : 0.0;
        }
    }
    if (lane == 0) {
        if (io.status) io.status[i] = status;
        if (io.nwsr) io.nwsr[i] = nwsr;
    }
}


// =====================================================================================================
// Fast path of the fused tick (CoM as constraint, all shipped configurations).  The kinematics is rooted
// at the stance foot, so the stance-foot rows of A are [X_s | 0] with X_s = [[I, -S(r_s)],[0, I]]: they
// pin the base twist, v_b = X_s^-1 b_s, in closed form.  What remains is a strictly convex QP in the
// joint velocities only,
//     min 1/2 dq' (Lambda + U' W U) dq + g_r' dq    s.t.  A_r dq = b_r,     A_r = [J_swing ; J_com] (9 x nJ),
// whose Hessian is diagonal plus rank 3 (U = neck angular Jacobian): H_r^-1 follows from the Woodbury
// identity and the multipliers from the 9x9 Schur complement A_r H_r^-1 A_r'.  This is the unique
// minimiser of the reference QP whenever no velocity bound is active (same point qpOASES returns with
// nWSR = 0); it needs ~5x fewer instructions than the general null-space factorisation of the 29 x 15
// problem.  Instances whose optimum violates a velocity bound, or whose Schur complement / weights are
// not safely positive definite, are appended to slow_list and solved by ik_kernel (null space +
// Goldfarb-Idnani) in a second, compacted launch: every status / nWSR decision stays with that kernel.
// =====================================================================================================
#ifndef WC_IK_FAST_WARPS
#define WC_IK_FAST_WARPS 4
#endif
constexpr int kFastWarps = WC_IK_FAST_WARPS;

template <int NJ>
struct IkFastLayout {
    static constexpr int NV = NJ + 6;
    static constexpr int KIN = WC_MAX_LINKS * 12 + WC_MAX_LINKS * 4;
    // rows of AS / UT / HS are read both lane-per-column and lane-per-ROW (G, S, rhs below): a row stride of 33
    // keeps the second pattern bank-conflict free (stride 32 put every row on the same bank pair, profiles/r1o)
    static constexpr int LDQ = 33;
    static constexpr int O_AS = 0;                  // a_j   [9][LDQ]
    static constexpr int O_UT = O_AS + 9 * LDQ;     // u~_j  [3][LDQ]
    static constexpr int O_HS = O_UT + 3 * LDQ;     // h_j   [9][LDQ]
    static constexpr int O_GR = O_HS + 9 * LDQ;     // g_r   [32]
    static constexpr int O_GK = O_GR + 32;          // G K   [27]
    static constexpr int O_SS = O_GK + 28;          // S     [81]
    static constexpr int QP = O_SS + 82;
    static constexpr int WORK = ((KIN > QP ? KIN : QP) + 15) / 16 * 16;
    // per-instance inputs prefetched by cp.async while the kinematics runs (one-instance SoA, stride 1)
    static constexpr int I_BASE = WORK, I_LFD = I_BASE + 12, I_RFD = I_LFD + 12, I_LTW = I_RFD + 12, I_RTW = I_LTW + 6;
    static constexpr int I_CPD = I_RTW + 6, I_CVD = I_CPD + 3, I_CM = I_CVD + 3, I_NRD = I_CM + 3, I_QM = I_NRD + 9;
    static constexpr int TOTAL = (I_QM + NJ + 15) / 16 * 16 + 8;
    static constexpr int MODEL_DOUBLES = (int)((sizeof(IkModelDev) + 7) / 8); // per-CTA copy of the model table
};

// SECOND = false: every instance, pass 0 only; an instance whose optimum violates a bound is appended to second_list.
// SECOND = true : the instances of second_list, pass 1 only (the selected bound travels with the list entry).  (Both passes in ONE launch were measured: the 10 % of
// warps that take the second pass hold their CTA's shared memory and registers 1.4x longer, 0.66 -> 0.83 ms.)
template <int NJ, bool SECOND>
__global__ void __launch_bounds__(kFastWarps * 32, WC_IK_FAST_MINB) ik_fast_kernel(IkDev d, IkTickIO io)
{
    using LY = IkFastLayout<NJ>;
    constexpr int NV = LY::NV;
    constexpr int LDQ = LY::LDQ;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (SECOND && (long long)blockIdx.x * kFastWarps >= (long long)d.slow_count[1]) return; // CTA beyond the list
    // model table: one shared-memory copy per CTA (every lane indexes it by its own link / joint)
#if WC_IK_MODEL_SMEM
    double* sm_model = smem_all + (size_t)kFastWarps * LY::TOTAL;
    {
        const double* src = reinterpret_cast<const double*>(d.model);
        for (int k = threadIdx.x; k < LY::MODEL_DOUBLES; k += kFastWarps * 32) sm_model[k] = src[k];
    }
    __syncthreads();
#endif
    // (a one-wave grid striding over the instances, to stage the model table once per CTA, was measured: the loop
    // costs more registers / spills than the copy's latency it saves, 0.71 -> 0.74 ms)
    const long long inst = (long long)blockIdx.x * kFastWarps + warp;
    if (inst >= (SECOND ? (long long)d.slow_count[1] : (long long)d.B)) return;
    const int i = SECOND ? d.second_list[inst] : (int)inst;
    const size_t B = (size_t)d.B;
    double* sm = smem_all + (size_t)warp * LY::TOTAL;
    // ---- prefetch the instance's task inputs (they are first needed after the kinematics)
    {
        auto pf = [&](int off, const double* src, int rows) {
            for (int r = lane; r < rows; r += 32) cp_async8(sm + off + r, src + (size_t)r * B + i);
        };
        if (io.base_H) pf(LY::I_BASE, io.base_H, 12);
        pf(LY::I_LFD, io.lf_H_des, 12); pf(LY::I_RFD, io.rf_H_des, 12);
        pf(LY::I_LTW, io.lf_twist, 6); pf(LY::I_RTW, io.rf_twist, 6); pf(LY::I_CPD, io.com_pos_des, 3);
        pf(LY::I_CVD, io.com_vel_des, 3); pf(LY::I_CM, io.com_meas, 3); pf(LY::I_NRD, io.neck_R_des, 9);
        pf(LY::I_QM, io.q_meas, NJ);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    double* WL = sm; double* MC = sm + WC_MAX_LINKS * 12;
    double* AS = sm + LY::O_AS; double* UT = sm + LY::O_UT; double* HS = sm + LY::O_HS;
    double* GR = sm + LY::O_GR; double* GK = sm + LY::O_GK; double* SS = sm + LY::O_SS;
#if WC_IK_MODEL_SMEM
    const IkModelDev& M = *reinterpret_cast<const IkModelDev*>(sm_model);
#else
    const IkModelDev& M = *d.model; // 5 KB table read through L1 (every CTA of the SM shares the lines)
#endif
    const wc_ik_config& cfg = d.cfg;
    const int j = lane;
    const bool isj = j >= 6 && j < NV;

    double colLf[6] = {0, 0, 0, 0, 0, 0}, colRf[6] = {0, 0, 0, 0, 0, 0}, colN[3] = {0, 0, 0}, colC[3] = {0, 0, 0};
    double lfH[12], rfH[12], nR[9], com[3], pB[3];
    // (the stance-sole transform is needed half way through the kinematics: wait for the prefetch there would
    // serialise, so base_H alone still comes from HBM inside fk_stage; everything else arrives meanwhile)
    fk_stage<NV, true>(M, io, i, B, lane, WL, MC, colLf, colRf, colN, colC, lfH, rfH, nR, com, pB);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    double bvec[15], gj, lam;
    {
        IkTickIO ios = io; // one-instance view of the prefetched inputs: element k at ptr[k] (B = 1, i = 0)
        ios.lf_H_des = sm + LY::I_LFD; ios.rf_H_des = sm + LY::I_RFD; ios.lf_twist = sm + LY::I_LTW; ios.rf_twist = sm + LY::I_RTW;
        ios.com_pos_des = sm + LY::I_CPD; ios.com_vel_des = sm + LY::I_CVD; ios.com_meas = sm + LY::I_CM;
        ios.neck_R_des = sm + LY::I_NRD; ios.q_meas = sm + LY::I_QM;
        qp_vectors<true>(d, ios, 0, 1, i, lane, lfH, rfH, nR, colN, colC, bvec, gj, lam);
    }
    const bool left = io.left_is_fixed[i] != 0;

    // ---- base twist from the stance-foot rows: v_lin + w x r_s = b_lin, w = b_ang
    double bs[6], bw[6], a[9], u[3];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        bs[r] = left ? bvec[r] : bvec[6 + r];
        bw[r] = left ? bvec[6 + r] : bvec[r];
        a[r] = isj ? (left ? colRf[r] : colLf[r]) : 0.0;
    }
#pragma unroll
    for (int r = 0; r < 3; r++) { a[6 + r] = isj ? colC[r] : 0.0; u[r] = isj ? colN[r] : 0.0; }
    const double pS[3] = {left ? lfH[3] : rfH[3], left ? lfH[7] : rfH[7], left ? lfH[11] : rfH[11]};
    const double pW[3] = {left ? rfH[3] : lfH[3], left ? rfH[7] : lfH[7], left ? rfH[11] : lfH[11]};
    const double om[3] = {bs[3], bs[4], bs[5]};
    double rS[3], rW[3], rC[3], t3[3], vlin[3];
#pragma unroll
    for (int r = 0; r < 3; r++) { rS[r] = pS[r] - pB[r]; rW[r] = pW[r] - pB[r]; rC[r] = com[r] - pB[r]; }
    cross3(om, rS, t3);
#pragma unroll
    for (int r = 0; r < 3; r++) vlin[r] = bs[r] - t3[r];
    double br[9];
    cross3(om, rW, t3);
#pragma unroll
    for (int r = 0; r < 3; r++) { br[r] = bw[r] - (vlin[r] + t3[r]); br[3 + r] = bw[3 + r] - om[r]; }
    cross3(om, rC, t3);
#pragma unroll
    for (int r = 0; r < 3; r++) br[6 + r] = bvec[12 + r] - (vlin[r] + t3[r]);
    // reduced gradient: g_j + (H_jb v_b)_j = g_j + u_j' W w
    const double* Wn = cfg.neck_weight;
    double Wom[3];
#pragma unroll
    for (int r = 0; r < 3; r++) Wom[r] = Wn[3 * r] * om[0] + Wn[3 * r + 1] * om[1] + Wn[3 * r + 2] * om[2];
    const double gr = isj ? gj + (u[0] * Wom[0] + u[1] * Wom[1] + u[2] * Wom[2]) : 0.0;
    bool fallback = __any_sync(FULL, isj && !(lam > 0.0));
    // Two passes at most.  Pass 0 is the equality-constrained optimum.  When it violates velocity bounds, the bound
    // Goldfarb-Idnani would add first (most negative slack, ties to the smaller id: gi_bounds) is fixed at its value
    // and the same Woodbury / Schur solve runs on the remaining joints: with an empty working set GI's first step is
    // a full step onto that bound, so if nothing is violated afterwards this IS its answer, with nWSR = 1 (98.5 % of
    // the instances that used to go to the general kernel).  Anything else still goes to the general kernel.
    bool fixedl = false;  // this lane's variable sits on its bound
    double beta = 0.0;    // ... at this value
    double gr_e = gr;     // gradient and right-hand side of the pass: + H_jp beta, - a_p beta
    double br_e[9];
#pragma unroll
    for (int r = 0; r < 9; r++) br_e[r] = br[r];
    double nu = 0.0;
    int nws = 0;
    // fix variable id >> 1 at its lower (id even) or upper (id odd) bound for the next pass
    auto fix_bound = [&](int id) {
        const double lb = isj ? cfg.min_vel[j - 6] : -DBL_MAX, ub = isj ? cfg.max_vel[j - 6] : DBL_MAX;
        const int pvar = id >> 1;
        fixedl = lane == pvar;
        beta = __shfl_sync(FULL, (id & 1) ? ub : lb, pvar);
        double up3[3], wup[3];
#pragma unroll
        for (int r = 0; r < 3; r++) up3[r] = __shfl_sync(FULL, u[r], pvar);
#pragma unroll
        for (int r = 0; r < 3; r++) wup[r] = Wn[3 * r] * up3[0] + Wn[3 * r + 1] * up3[1] + Wn[3 * r + 2] * up3[2];
        gr_e = (isj && !fixedl) ? gr + (u[0] * wup[0] + u[1] * wup[1] + u[2] * wup[2]) * beta : 0.0;
#pragma unroll
        for (int r = 0; r < 9; r++) br_e[r] = br[r] - __shfl_sync(FULL, a[r], pvar) * beta;
        nws = 1;
    };
    // the second launch starts from the bound the first launch selected (stored next to the list entry)
    if (SECOND) fix_bound(d.second_list[(size_t)d.B + inst]);
#pragma unroll 1
    for (int pass = SECOND ? 1 : 0;; pass++) {
    const double linv = (isj && lam > 0.0 && !fixedl) ? 1.0 / lam : 0.0;
    double ut[3], at[9];
#pragma unroll
    for (int r = 0; r < 3; r++) ut[r] = u[r] * linv;
#pragma unroll
    for (int r = 0; r < 9; r++) at[r] = a[r] * linv;
    // ---- Woodbury core: C = U Lambda^-1 U' (3x3), K = (I + W C)^-1 W
    double Cm[9];
    Cm[0] = warp_sum(ut[0] * u[0]); Cm[1] = warp_sum(ut[0] * u[1]); Cm[2] = warp_sum(ut[0] * u[2]);
    Cm[4] = warp_sum(ut[1] * u[1]); Cm[5] = warp_sum(ut[1] * u[2]); Cm[8] = warp_sum(ut[2] * u[2]);
    Cm[3] = Cm[1]; Cm[6] = Cm[2]; Cm[7] = Cm[5];
    double M3[9], K3[9];
    mat3_mul(Wn, Cm, M3);
    M3[0] += 1.0; M3[4] += 1.0; M3[8] += 1.0;
    {
        const double c00 = M3[4] * M3[8] - M3[5] * M3[7], c01 = M3[5] * M3[6] - M3[3] * M3[8], c02 = M3[3] * M3[7] - M3[4] * M3[6];
        const double det = M3[0] * c00 + M3[1] * c01 + M3[2] * c02;
        if (!(fabs(det) > 1e-12)) fallback = true;
        const double id = 1.0 / det;
        const double Mi[9] = {c00 * id, (M3[2] * M3[7] - M3[1] * M3[8]) * id, (M3[1] * M3[5] - M3[2] * M3[4]) * id,
                              c01 * id, (M3[0] * M3[8] - M3[2] * M3[6]) * id, (M3[2] * M3[3] - M3[0] * M3[5]) * id,
                              c02 * id, (M3[1] * M3[6] - M3[0] * M3[7]) * id, (M3[0] * M3[4] - M3[1] * M3[3]) * id};
        mat3_mul(Mi, Wn, K3);
    }
    __syncwarp(); // kinematics scratch is dead
#pragma unroll
    for (int r = 0; r < 9; r++) AS[r * LDQ + lane] = a[r];
#pragma unroll
    for (int r = 0; r < 3; r++) UT[r * LDQ + lane] = ut[r];
    GR[lane] = gr_e;
    __syncwarp();
    // ---- G = A_r Lambda^-1 U' (9x3), lane r < 9 owns row r, then the row of G K
    if (lane < 9) {
        double g0 = 0.0, g1 = 0.0, g2 = 0.0;
        for (int jj = 6; jj < NV; jj++) {
            const double av = AS[lane * LDQ + jj];
            g0 += av * UT[jj]; g1 += av * UT[LDQ + jj]; g2 += av * UT[2 * LDQ + jj];
        }
#pragma unroll
        for (int c = 0; c < 3; c++) GK[lane * 3 + c] = g0 * K3[c] + g1 * K3[3 + c] + g2 * K3[6 + c];
    }
    __syncwarp();
    // ---- h_j = column j of A_r H_r^-1
    double h[9];
#pragma unroll
    for (int r = 0; r < 9; r++) h[r] = at[r] - (GK[r * 3] * ut[0] + GK[r * 3 + 1] * ut[1] + GK[r * 3 + 2] * ut[2]);
#pragma unroll
    for (int r = 0; r < 9; r++) HS[r * LDQ + lane] = h[r];
    __syncwarp();
    // ---- S = A_r H_r^-1 A_r' (upper triangle: 45 entries on 32 lanes, two rounds) and the right-hand side
#pragma unroll
    for (int round = 0; round < 2; round++) {
        int e = lane + 32 * round;
        if (e < 45) {
            int r = 0;
            while (e >= 9 - r) { e -= 9 - r; r++; }
            const int c = r + e;
            double acc = 0.0;
            for (int jj = 6; jj < NV; jj++) acc += HS[r * LDQ + jj] * AS[c * LDQ + jj];
            SS[r * 9 + c] = acc; SS[c * 9 + r] = acc;
        }
    }
    double rhs = 0.0;
    if (lane < 9) {
        double acc = 0.0;
        for (int jj = 6; jj < NV; jj++) acc += HS[lane * LDQ + jj] * GR[jj];
        double brl = 0.0;
#pragma unroll
        for (int r = 0; r < 9; r++) brl = (lane == r) ? br_e[r] : brl;
        rhs = -(brl + acc);
    }
    __syncwarp();
    // ---- Cholesky S = L L' (row per lane, right-looking, pivots broadcast) and S mu = rhs
    double Mr[9];
#pragma unroll
    for (int c = 0; c < 9; c++) Mr[c] = (lane < 9) ? SS[lane * 9 + c] : 0.0;
    double mmax = 0.0;
#pragma unroll
    for (int c = 0; c < 9; c++) if (lane == c) mmax = fabs(Mr[c]);
    mmax = warp_max(mmax);
    double ldinv = 1.0;
#pragma unroll
    for (int jc = 0; jc < 9; jc++) {
        const double djj = __shfl_sync(FULL, Mr[jc], jc);
        if (!(djj > 1e-10 * (mmax > 1.0 ? mmax : 1.0))) { fallback = true; break; }
        const double piv = sqrt(djj);
        const double pinv = 1.0 / piv;
        const double lij = (lane == jc) ? piv : Mr[jc] * pinv;
        if (lane == jc) ldinv = pinv;
        if (lane >= jc) Mr[jc] = lij;
#pragma unroll
        for (int k = jc + 1; k < 9; k++) {
            const double lkj = __shfl_sync(FULL, Mr[jc], k);
            if (lane >= k) Mr[k] -= lij * lkj;
        }
    }
    if (fallback) break;
    {
        if (lane < 9) {
#pragma unroll
            for (int c = 0; c < 9; c++) SS[lane * 9 + c] = (c <= lane) ? Mr[c] : 0.0; // L (for the transposed solve)
        }
        __syncwarp();
        double acc = rhs;
        double mu[9];
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const double sk = __shfl_sync(FULL, acc * ldinv, k);
            if (lane > k) acc -= Mr[k] * sk;
            if (lane == k) acc = sk;
        }
#pragma unroll
        for (int k = 8; k >= 0; k--) {
            mu[k] = __shfl_sync(FULL, acc * ldinv, k);
            if (lane < k) acc -= SS[k * 9 + (lane < 9 ? lane : 0)] * mu[k];
        }
        // ---- dq = -H_r^-1 (g_r + A_r' mu)
        double t = gr_e;
#pragma unroll
        for (int r = 0; r < 9; r++) t += a[r] * mu[r];
        const double tt = t * linv;
        double s3[3], ks[3];
#pragma unroll
        for (int r = 0; r < 3; r++) s3[r] = warp_sum(u[r] * tt);
#pragma unroll
        for (int r = 0; r < 3; r++) ks[r] = K3[3 * r] * s3[0] + K3[3 * r + 1] * s3[1] + K3[3 * r + 2] * s3[2];
        nu = isj ? -(tt - (ut[0] * ks[0] + ut[1] * ks[1] + ut[2] * ks[2])) : 0.0;
        if (fixedl) nu = beta;
        // velocity bounds (setVelocityBounds, :23-51): same feasibility test as the general kernel
        const double lb = isj ? cfg.min_vel[j - 6] : -DBL_MAX, ub = isj ? cfg.max_vel[j - 6] : DBL_MAX;
        const bool viol = isj && !fixedl && ((lb > -kQpInf && nu < lb - kViolTol) || (ub < kQpInf && nu > ub + kViolTol));
        if (!__any_sync(FULL, viol)) break;       // optimal: nWSR = number of bounds fixed so far
        if (pass == 1) { fallback = true; break; } // a second working-set change: general kernel
        // most violated bound, GI's rule: min over (slack, id = 2 j + upper), ties to the smaller id
        double best = INFINITY; int id = 0x7fffffff;
        if (isj) {
            if (lb > -kQpInf) { best = nu - lb; id = 2 * j; }
            if (ub < kQpInf) { const double sl = ub - nu; if (sl < best) { best = sl; id = 2 * j + 1; } }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, id, o);
            if (ob < best || (ob == best && oi < id)) { best = ob; id = oi; }
        }
        if (!SECOND) { // hand over to the second launch, with the selected bound
            if (lane == 0) {
                const int idx = atomicAdd(d.slow_count + 1, 1);
                d.second_list[idx] = i; d.second_list[(size_t)d.B + idx] = id;
            }
            return;
        }
        fix_bound(id);
        __syncwarp(); // the next pass overwrites the shared-memory operands
    }
    } // passes
    if (fallback) {
        if (lane == 0) d.slow_list[atomicAdd(d.slow_count, 1)] = i;
        return;
    }
    // ---- outputs: getSolution (:341-362), get{Left,Right}FootError (:364-401) = b - J nu
    if (isj && io.dq) io.dq[(size_t)(j - 6) * B + i] = nu;
    if (io.lf_err || io.rf_err) {
        double ew[6], es[6];
#pragma unroll
        for (int r = 0; r < 6; r++) ew[r] = br[r] - warp_sum(a[r] * nu);
        cross3(om, rS, t3);
#pragma unroll
        for (int r = 0; r < 3; r++) { es[r] = bs[r] - (vlin[r] + t3[r]); es[3 + r] = 0.0; }
#pragma unroll
        for (int r = 0; r < 6; r++) {
            if (lane == r && io.lf_err) io.lf_err[(size_t)r * B + i] = left ? es[r] : ew[r];
            if (lane == r && io.rf_err) io.rf_err[(size_t)r * B + i] = left ? ew[r] : es[r];
        }
    }
    if (lane == 0) {
        if (io.status) io.status[i] = WC_QP_SUCCESSFUL_RETURN;
        if (io.nwsr) io.nwsr[i] = nws;
    }
}

#include "ik_osqp.inl"

template <int NJ, bool COMC>
cudaError_t launch_t(const IkDev& d, const IkTickIO& io, int mode, cudaStream_t s, cudaEvent_t mid)
{
    if (d.solver == WC_IK_SOLVER_OSQP) { // the OSQP-based variant has no split form: everything happens in the solve call
        if (mode == IK_MODE_TICK_KIN) return cudaSuccess;
        if (mode == IK_MODE_TICK_SOLVE) mode = IK_MODE_TICK;
    }
    if (d.solver == WC_IK_SOLVER_OSQP && mode != IK_MODE_FK_ONLY) return launch_osqp_t<NJ, COMC>(d, io, mode, s);
    using LY = IkLayout<NJ, COMC>;
    const size_t smem = (size_t)LY::TOTAL * sizeof(double) * kWarpsPerBlock;
    cudaError_t e = cudaFuncSetAttribute(ik_kernel<NJ, COMC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int blocks = (d.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
    // the tick as two calls: with the thread-per-instance fast path the kinematics part is a launch of its own; any other
    // configuration does the whole tick in the solve call (the kinematics call is then a no-op)
    const bool split_ok = COMC && d.slow_list && d.tpi_scratch && d.solver != WC_IK_SOLVER_OSQP;
    if (mode == IK_MODE_TICK_KIN) return split_ok ? launch_ik_tpi(d, io, s, 1) : cudaSuccess;
    if (mode == IK_MODE_TICK_SOLVE) {
        if (!split_ok) return launch_t<NJ, COMC>(d, io, IK_MODE_TICK, s, mid);
        e = cudaMemsetAsync(d.slow_count, 0, 2 * sizeof(int32_t), s);
        if (e != cudaSuccess) return e;
        e = launch_ik_tpi(d, io, s, 2);
        if (e != cudaSuccess) return e;
        if (mid) cudaEventRecord(mid, s);
        ik_kernel<NJ, COMC><<<blocks, kWarpsPerBlock * 32, smem, s>>>(d, io, IK_MODE_TICK, 1);
        return cudaGetLastError();
    }
    if (COMC && mode == IK_MODE_TICK && d.slow_list && (d.tpi_scratch || NJ == 23)) {
        // fast path for every instance, then the general kernel on the compacted remainder
        e = cudaMemsetAsync(d.slow_count, 0, 2 * sizeof(int32_t), s);
        if (e != cudaSuccess) return e;
        if (d.tpi_scratch) { // thread-per-instance fast path (ik_tpi.cu, any n_dofs), then the general kernel on slow_list
            e = launch_ik_tpi(d, io, s);
            if (e != cudaSuccess) return e;
            if (mid) cudaEventRecord(mid, s);
            ik_kernel<NJ, COMC><<<blocks, kWarpsPerBlock * 32, smem, s>>>(d, io, mode, 1);
            return cudaGetLastError();
        }
        if constexpr (NJ == 23) { // warp-per-instance fast kernels: instantiated for the 23-DoF iCub set only (WC_IK_PATH=warp)
            using LF = IkFastLayout<NJ>;
            const size_t smem_f = ((size_t)LF::TOTAL * kFastWarps + (WC_IK_MODEL_SMEM ? LF::MODEL_DOUBLES : 0)) * sizeof(double);
            e = cudaFuncSetAttribute(ik_fast_kernel<NJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(ik_fast_kernel<NJ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
            if (e != cudaSuccess) return e;
            ik_fast_kernel<NJ, false><<<(d.B + kFastWarps - 1) / kFastWarps, kFastWarps * 32, smem_f, s>>>(d, io);
            ik_fast_kernel<NJ, true><<<(d.B + kFastWarps - 1) / kFastWarps, kFastWarps * 32, smem_f, s>>>(d, io);
            if (mid) cudaEventRecord(mid, s); // per-kernel timing (wc_ik_enable_profiling)
            ik_kernel<NJ, COMC><<<blocks, kWarpsPerBlock * 32, smem, s>>>(d, io, mode, 1);
            return cudaGetLastError();
        }
    }
    ik_kernel<NJ, COMC><<<blocks, kWarpsPerBlock * 32, smem, s>>>(d, io, mode, 0);
    return cudaGetLastError();
}

} // namespace

// One translation unit per number of actuated joints (-DWC_IK_NJ=n, csrc/Makefile): lane j <-> QP variable j makes the variable
// count a compile-time size of the warp-per-instance kernels.  launch_ik (wc_ik_api.cu) dispatches on the run-time n_dofs.
#ifndef WC_IK_NJ
#define WC_IK_NJ 23
#endif
#define WC_IK_CAT2(a, b) a##b
#define WC_IK_CAT(a, b) WC_IK_CAT2(a, b)
cudaError_t WC_IK_CAT(launch_ik_nj, WC_IK_NJ)(const IkDev& d, const IkTickIO& io, int mode, cudaStream_t s, cudaEvent_t mid)
{
    if (d.nJ != WC_IK_NJ) return cudaErrorInvalidValue;
    if (d.cfg.use_com_as_constraint) return launch_t<WC_IK_NJ, true>(d, io, mode, s, mid);
    return launch_t<WC_IK_NJ, false>(d, io, mode, s, mid);
}

} // namespace wcb
