// ik_tpi.cu — THREAD-PER-INSTANCE fast path of the fused FK + QP-IK tick (K3/K4, CoM as constraint: all shipped configurations).
//
// Same mathematics as ik_fast_kernel (ik_kernels.cu; mapping to the reference in ik_kernels.cuh:
// WalkingFK WM/src/WalkingForwardKinematics.cpp:217-276,411-454, WalkingQPIK_qpOASES::{setHessianMatrix, setGradientVector,
// setLinearConstraintMatrix, setBounds, solve, getSolution, get*FootError} WM/src/WalkingQPInverseKinematics_qpOASES.cpp:135-401):
// the stance-foot rows pin the base twist in closed form, the remaining QP in the joint velocities has the Hessian
// Lambda + U' W U (diagonal + rank 3) -> Woodbury inverse, 9x9 Schur complement, Cholesky; an optimum that violates a velocity bound is
// re-solved once with the bound Goldfarb-Idnani adds first fixed at its value (second launch), everything else goes to the general
// kernel (ik_kernel, warp per instance) through slow_list.
//
// What changes is the mapping to the machine.  With a warp per instance (ik_fast_kernel) every per-instance scalar step (frame
// transforms, task errors, the 3x3 Woodbury core, the 9x9 Cholesky with broadcast pivots, every reduction over the joints) costs
// full warp instructions: 4 400 warp instructions per instance, 39 % issue utilisation, 16 instances resident per SM
// (profiles/r1s_summary.md).  Here ONE THREAD owns one instance: nothing is replicated, reductions over the joints are plain
// loops with the accumulators in registers (C 6, G 27, S 45, right-hand sides 12), the Cholesky factor lives in registers,
// there is not a single shuffle, and all lanes are busy: ~13 000 instructions per instance = ~400 warp instructions per instance.
// Control flow depends only on the model (uniform across the warp); `left_is_fixed` is applied by selects.
//
// Storage.  The serial transform composition needs root_H_parent: links whose transform is read by a later link that is not
// the next one (branch points) or after the loop (the three task frames) park it in a shared-memory slot ([slot][12][32 lanes],
// conflict-free); the host assigns the slots by a linear scan over the tree (build_model, wc_ik_api.cu): 3 KB per slot and
// warp.  Per-joint data that is produced in the link loop but needed after it (joint axis / origin, the running first-moment
// prefix that gives every sub-tree moment as a difference of two prefixes, then the QP columns) goes to a per-instance scratch
// in global memory, field-major [row][B]: written and re-read by the same thread within microseconds, i.e. L2 traffic.  The
// second launch (violators, compacted) restarts from the scratch: no kinematics is repeated.
#include "ik_kernels.cuh"
#include "ik_math.cuh"

#include <cfloat>
#include <cmath>

namespace wcb {

namespace {

constexpr double kQpInfT = 1e20;     // qpOASES INFTY
constexpr double kViolTolT = 1e-10;  // same feasibility test as ik_kernel / ik_fast_kernel
#ifndef WC_TPI_THREADS
#define WC_TPI_THREADS 128
#endif
#ifndef WC_TPI_MINB
#define WC_TPI_MINB 2
#endif
constexpr int kTpiThreads = WC_TPI_THREADS;

// scratch rows of one instance
struct TpiRows {
    int oJ, oP, oINST, total;
    __host__ __device__ TpiRows(int L, int nJ)
    {
        oJ = 0;                 // [nJ][10] link loop: joint axis (3) and origin (3) in the root frame; then, in place, the QP column of
                                //          the joint: swing-foot linear part s a x (p_sw - o) (3), world axis a (3), CoM part (3), g_r
        oP = oJ + 10 * nJ;      // [L+1][3] prefix of the mass-weighted link CoMs in the root frame (P_l = sum over k < l)
        oINST = oP + 3 * (L + 1); // [13]   b_r (9) and the stance-foot error (3), for the second solve (and for the solve part of the
                                  //        split tick, whose kinematics part leaves b_r[6..8] without the CoM term); fallback flag
        total = oINST + 13;
    }
};

// accumulators of the reduced problem over the joints: C = U L^-1 U', G = A L^-1 U', S = A L^-1 A' (upper triangle by rows),
// ra = A L^-1 g, ru = U L^-1 g
struct TpiAcc {
    double C[6], G[27], S[45], ra[9], ru[3];
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int k = 0; k < 6; k++) C[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 27; k++) G[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 45; k++) S[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 9; k++) ra[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) ru[k] = 0.0;
    }
    __device__ __forceinline__ void add(const double (&a)[9], const double (&u)[3], double g, double linv)
    {
        double ut[3];
#pragma unroll
        for (int r = 0; r < 3; r++) ut[r] = u[r] * linv;
        C[0] += ut[0] * u[0]; C[1] += ut[0] * u[1]; C[2] += ut[0] * u[2];
        C[3] += ut[1] * u[1]; C[4] += ut[1] * u[2]; C[5] += ut[2] * u[2];
#pragma unroll
        for (int c = 0; c < 3; c++) ru[c] += ut[c] * g;
        int e = 0;
#pragma unroll
        for (int r = 0; r < 9; r++) {
            const double atr = a[r] * linv;
            ra[r] += atr * g;
#pragma unroll
            for (int c = 0; c < 3; c++) G[3 * r + c] += a[r] * ut[c];
#pragma unroll
            for (int c = r; c < 9; c++) S[e++] += atr * a[c];
        }
    }
};

// path signs of joint j for this stance choice: swing foot, neck
__device__ __forceinline__ void tpi_signs(const IkModelDev& M, bool left, int j, double& ssw, double& sn)
{
    ssw = left ? (double)M.s_frame[1][1][j] : (double)M.s_frame[0][0][j];
    sn = left ? (double)M.s_frame[1][2][j] : (double)M.s_frame[0][2][j];
}

// ---- the solve, from the accumulators (first solve) or from the columns in the scratch (SECOND: variable `pvar` is fixed at
// `beta`, GI's first working-set change).  colj: rows of joint 0 of THIS instance, row stride B.
// Returns 0 = optimal (outputs written), 1 = a bound is violated (id of the most violated bound in *viol_id), 2 = fallback.
template <bool SECOND>
__device__ __forceinline__ int tpi_solve(const IkDev& d, const IkTickIO& io, const IkModelDev& M, TpiAcc& A, const double* colj, size_t B,
                                         int i, int nJ, const double (&br)[9], const double (&es)[3], bool left, int pvar, double beta,
                                         int* viol_id)
{
    const wc_ik_config& cfg = d.cfg;
    const double* Wn = cfg.neck_weight;
    double br_e[9];
#pragma unroll
    for (int r = 0; r < 9; r++) br_e[r] = br[r];
    double wup[3] = {0, 0, 0};
    if (SECOND) {
        double ssw, sn;
        tpi_signs(M, left, pvar, ssw, sn);
        double cp[9];
#pragma unroll
        for (int r = 0; r < 9; r++) cp[r] = colj[(size_t)(10 * pvar + r) * B];
        const double up[3] = {sn * cp[3], sn * cp[4], sn * cp[5]};
#pragma unroll
        for (int r = 0; r < 3; r++) wup[r] = Wn[3 * r] * up[0] + Wn[3 * r + 1] * up[1] + Wn[3 * r + 2] * up[2];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            br_e[r] = br[r] - cp[r] * beta;
            br_e[3 + r] = br[3 + r] - (ssw * cp[3 + r]) * beta;
            br_e[6 + r] = br[6 + r] - cp[6 + r] * beta;
        }
    }
    // ---- accumulate over the joints; the ten rows of the next column are requested before the current one is used
    A.clear();
    {
        // two column buffers in ping-pong (loop unrolled by two): the loads that fill one are in flight while the other is used,
        // and the two groups of loads sit on different scoreboards
        auto load = [&](int jn, double (&buf)[10]) {
#pragma unroll
            for (int r = 0; r < 10; r++) buf[r] = colj[(size_t)(10 * jn + r) * B];
        };
        auto use = [&](int j, const double (&c10)[10]) {
            if (SECOND && j == pvar) return;
            double ssw, sn;
            tpi_signs(M, left, j, ssw, sn);
            const double a[9] = {c10[0], c10[1], c10[2], ssw * c10[3], ssw * c10[4], ssw * c10[5], c10[6], c10[7], c10[8]};
            const double u[3] = {sn * c10[3], sn * c10[4], sn * c10[5]};
            double g = c10[9];
            if (SECOND) g += (u[0] * wup[0] + u[1] * wup[1] + u[2] * wup[2]) * beta;
            A.add(a, u, g, 1.0 / cfg.joint_reg_weight[j]);
        };
        double bA[10], bB[10];
        load(0, bA);
#pragma unroll 1
        for (int j = 0; j < nJ; j += 2) {
            if (j + 1 < nJ) load(j + 1, bB);
            use(j, bA);
            if (j + 1 < nJ) {
                if (j + 2 < nJ) load(j + 2, bA);
                use(j + 1, bB);
            }
        }
    }
    // ---- Woodbury core K = (I + W C)^-1 W
    double K3[9];
    {
        const double Cm[9] = {A.C[0], A.C[1], A.C[2], A.C[1], A.C[3], A.C[4], A.C[2], A.C[4], A.C[5]};
        double M3[9];
        mat3_mul(Wn, Cm, M3);
        M3[0] += 1.0; M3[4] += 1.0; M3[8] += 1.0;
        const double c00 = M3[4] * M3[8] - M3[5] * M3[7], c01 = M3[5] * M3[6] - M3[3] * M3[8], c02 = M3[3] * M3[7] - M3[4] * M3[6];
        const double det = M3[0] * c00 + M3[1] * c01 + M3[2] * c02;
        if (!(fabs(det) > 1e-12)) return 2;
        const double id = 1.0 / det;
        const double Mi[9] = {c00 * id, (M3[2] * M3[7] - M3[1] * M3[8]) * id, (M3[1] * M3[5] - M3[2] * M3[4]) * id,
                              c01 * id, (M3[0] * M3[8] - M3[2] * M3[6]) * id, (M3[2] * M3[3] - M3[0] * M3[5]) * id,
                              c02 * id, (M3[1] * M3[6] - M3[0] * M3[7]) * id, (M3[0] * M3[4] - M3[1] * M3[3]) * id};
        mat3_mul(Mi, Wn, K3);
    }
    // ---- S = S0 - (G K) G', rhs = -(b_r + ra - (G K) ru)
    double* S = A.S;
    double rhs[9];
    {
        int e = 0;
#pragma unroll
        for (int r = 0; r < 9; r++) {
            double gk[3];
#pragma unroll
            for (int c = 0; c < 3; c++) gk[c] = A.G[3 * r] * K3[c] + A.G[3 * r + 1] * K3[3 + c] + A.G[3 * r + 2] * K3[6 + c];
            rhs[r] = -(br_e[r] + (A.ra[r] - (gk[0] * A.ru[0] + gk[1] * A.ru[1] + gk[2] * A.ru[2])));
#pragma unroll
            for (int c = r; c < 9; c++, e++) S[e] -= gk[0] * A.G[3 * c] + gk[1] * A.G[3 * c + 1] + gk[2] * A.G[3 * c + 2];
        }
    }
    // ---- Cholesky S = L L' in place (S holds the upper triangle by rows = L' by rows), S mu = rhs
    // entry (r, c >= r) at r * 9 - r (r - 1) / 2 + (c - r)
#define SI(r, c) ((r) * 9 - (r) * ((r) - 1) / 2 + ((c) - (r)))
    double mmax = 0.0;
#pragma unroll
    for (int r = 0; r < 9; r++) mmax = fmax(mmax, fabs(S[SI(r, r)]));
    const double ptol = 1e-10 * (mmax > 1.0 ? mmax : 1.0);
    double pinv[9];
#pragma unroll
    for (int jc = 0; jc < 9; jc++) {
        const double djj = S[SI(jc, jc)];
        if (!(djj > ptol)) return 2;
        pinv[jc] = rsqrt(djj);
        S[SI(jc, jc)] = djj * pinv[jc];
#pragma unroll
        for (int c = jc + 1; c < 9; c++) S[SI(jc, c)] *= pinv[jc];
#pragma unroll
        for (int r = jc + 1; r < 9; r++)
#pragma unroll
            for (int c = r; c < 9; c++) S[SI(r, c)] -= S[SI(jc, r)] * S[SI(jc, c)];
    }
    double mu[9];
#pragma unroll
    for (int r = 0; r < 9; r++) { // L y = rhs  (L(r, k) = S(k, r))
        double acc = rhs[r];
#pragma unroll
        for (int k = 0; k < r; k++) acc -= S[SI(k, r)] * mu[k];
        mu[r] = acc * pinv[r];
    }
#pragma unroll
    for (int r = 8; r >= 0; r--) { // L' mu = y
        double acc = mu[r];
#pragma unroll
        for (int k = r + 1; k < 9; k++) acc -= S[SI(r, k)] * mu[k];
        mu[r] = acc * pinv[r];
    }
#undef SI
    // ---- dq = -H_r^-1 (g_r + A_r' mu) = -(tt - L^-1 U' K s), tt_j = (g_j + a_j' mu) / lambda_j, s = U tt = ru + G' mu
    double ks[3];
    {
        double s3[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double acc = A.ru[c];
#pragma unroll
            for (int r = 0; r < 9; r++) acc += A.G[3 * r + c] * mu[r];
            s3[c] = acc;
        }
#pragma unroll
        for (int r = 0; r < 3; r++) ks[r] = K3[3 * r] * s3[0] + K3[3 * r + 1] * s3[1] + K3[3 * r + 2] * s3[2];
    }
    // ---- nu_j, the bounds (setVelocityBounds, _qpOASES.cpp:23-51), foot error b - J nu.  dq is written provisionally: an instance
    // that leaves this function unsolved is written again by the solve that finishes it
    double ew[6] = {0, 0, 0, 0, 0, 0};
    bool viol = false;
    double best = INFINITY; int id = 0x7fffffff;
    {
        auto load = [&](int jn, double (&buf)[10]) {
#pragma unroll
            for (int r = 0; r < 10; r++) buf[r] = colj[(size_t)(10 * jn + r) * B];
        };
        auto use = [&](int j, const double (&c10)[10]) {
        double ssw, sn;
        tpi_signs(M, left, j, ssw, sn);
        const double a[9] = {c10[0], c10[1], c10[2], ssw * c10[3], ssw * c10[4], ssw * c10[5], c10[6], c10[7], c10[8]};
        const double u[3] = {sn * c10[3], sn * c10[4], sn * c10[5]};
        double nu;
        if (SECOND && j == pvar) {
            nu = beta;
        } else {
            double t = c10[9];
            if (SECOND) t += (u[0] * wup[0] + u[1] * wup[1] + u[2] * wup[2]) * beta;
#pragma unroll
            for (int r = 0; r < 9; r++) t += a[r] * mu[r];
            const double linv = 1.0 / cfg.joint_reg_weight[j];
            const double tt = t * linv;
            nu = -(tt - (u[0] * linv * ks[0] + u[1] * linv * ks[1] + u[2] * linv * ks[2]));
            const double lb = cfg.min_vel[j], ub = cfg.max_vel[j];
            if ((lb > -kQpInfT && nu < lb - kViolTolT) || (ub < kQpInfT && nu > ub + kViolTolT)) viol = true;
            // most violated bound, GI's rule: min over (slack, id = 2 (j + 6) + upper), ties to the smaller id
            if (lb > -kQpInfT) { const double sl = nu - lb; if (sl < best) { best = sl; id = 2 * (j + 6); } }
            if (ub < kQpInfT) { const double sl = ub - nu; if (sl < best) { best = sl; id = 2 * (j + 6) + 1; } }
        }
        if (io.dq) io.dq[(size_t)j * B + i] = nu;
#pragma unroll
        for (int r = 0; r < 6; r++) ew[r] += a[r] * nu;
        };
        double bA[10], bB[10];
        load(0, bA);
#pragma unroll 1
        for (int j = 0; j < nJ; j += 2) {
            if (j + 1 < nJ) load(j + 1, bB);
            use(j, bA);
            if (j + 1 < nJ) {
                if (j + 2 < nJ) load(j + 2, bA);
                use(j + 1, bB);
            }
        }
    }
    if (viol) { *viol_id = id; return 1; }
    // ---- outputs: get{Left,Right}FootError (:364-401) = b - J nu
    if (io.lf_err || io.rf_err) {
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const double w = br[r] - ew[r], s = r < 3 ? es[r] : 0.0;
            if (io.lf_err) io.lf_err[(size_t)r * B + i] = left ? s : w;
            if (io.rf_err) io.rf_err[(size_t)r * B + i] = left ? w : s;
        }
    }
    if (io.status) io.status[i] = WC_QP_SUCCESSFUL_RETURN;
    if (io.nwsr) io.nwsr[i] = SECOND ? 1 : 0;
    return 0;
}

#define SLOT(s, k) slots[((s) * 12 + (k)) * 32]
// root_H_link of every link, serially (parent[l] < l), with the running first-moment prefix; SCRATCH: the prefix and every
// joint's axis / origin go to the per-instance scratch.  Returns the total first moment in the root frame in P.
template <bool SCRATCH>
__device__ __forceinline__ void tpi_link_loop(const IkModelDev& M, const IkTickIO& io, int i, size_t B, double* slots, double* sc,
                                              const TpiRows& rows, double (&P)[3])
{
    const int L = M.n_links;
    // ---- root_H_link, serially over the links (parent[l] < l); first-moment prefix; joint axes and origins
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0}; // root_H_(l-1) of the previous iteration
    P[0] = P[1] = P[2] = 0.0;
    double qn = 0.0, qn2 = 0.0; // joint angles of the next two links, requested ahead of their use
    {
        const int d0 = M.dof[0];
        if (d0 >= 0) qn = io.q_lin[(size_t)d0 * B + i];
        if (L > 1 && M.dof[1] >= 0) qn2 = io.q_lin[(size_t)M.dof[1] * B + i];
    }
#pragma unroll 1
    for (int l = 0; l < L; l++) {
        const double q = qn;
        qn = qn2;
        if (l + 2 < L) {
            const int dn = M.dof[l + 2];
            if (dn >= 0) qn2 = io.q_lin[(size_t)dn * B + i];
        }
        double Rl[9], pl[3];
#pragma unroll
        for (int k = 0; k < 9; k++) Rl[k] = M.joint_R[l][k];
#pragma unroll
        for (int k = 0; k < 3; k++) pl[k] = M.joint_p[l][k];
        const int dj = M.dof[l];
        const double ax[3] = {M.axis[l][0], M.axis[l][1], M.axis[l][2]};
        if (dj >= 0) {
            double s, c;
            sincos(q, &s, &c);
            const double v = 1.0 - c, x = ax[0], y = ax[1], z = ax[2];
            const double Rq[9] = {c + x * x * v, x * y * v - z * s, x * z * v + y * s,
                                  y * x * v + z * s, c + y * y * v, y * z * v - x * s,
                                  z * x * v - y * s, z * y * v + x * s, c + z * z * v};
            double Rn[9];
            mat3_mul(Rl, Rq, Rn);
#pragma unroll
            for (int k = 0; k < 9; k++) Rl[k] = Rn[k];
        }
        const int ps = M.tpi_pslot[l];
        if (ps >= 0) { // the parent is not the previous link: its transform waits in a slot
#pragma unroll
            for (int k = 0; k < 9; k++) R[k] = SLOT(ps, k);
#pragma unroll
            for (int k = 0; k < 3; k++) p[k] = SLOT(ps, 9 + k);
        }
        if (ps != -1) xf_mul(R, p, Rl, pl, R, p);
        else {
#pragma unroll
            for (int k = 0; k < 9; k++) R[k] = Rl[k];
#pragma unroll
            for (int k = 0; k < 3; k++) p[k] = pl[k];
        }
        const int so = M.tpi_slot[l];
        if (so >= 0) {
#pragma unroll
            for (int k = 0; k < 9; k++) SLOT(so, k) = R[k];
#pragma unroll
            for (int k = 0; k < 3; k++) SLOT(so, 9 + k) = p[k];
        }
        if (SCRATCH) {
#pragma unroll
            for (int r = 0; r < 3; r++) sc[(size_t)(rows.oP + 3 * l + r) * B] = P[r];
        }
        const double m = M.mass[l];
#pragma unroll
        for (int r = 0; r < 3; r++) P[r] += m * (R[3 * r] * M.com[l][0] + R[3 * r + 1] * M.com[l][1] + R[3 * r + 2] * M.com[l][2] + p[r]);
        if (SCRATCH && dj >= 0) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                sc[(size_t)(rows.oJ + 10 * dj + r) * B] = R[3 * r] * ax[0] + R[3 * r + 1] * ax[1] + R[3 * r + 2] * ax[2];
                sc[(size_t)(rows.oJ + 10 * dj + 3 + r) * B] = p[r];
            }
        }
    }
    if (SCRATCH) {
#pragma unroll
        for (int r = 0; r < 3; r++) sc[(size_t)(rows.oP + 3 * L + r) * B] = P[r];
    }
}

// world_H_root from the stance sole (evaluateWorldToBaseTransformation) and the task frames in the world
__device__ __forceinline__ void tpi_frames(const IkModelDev& M, const IkTickIO& io, int i, size_t B, const double* slots, bool left,
                                           double (&wR)[9], double (&wp)[3], double (&lfH)[12], double (&rfH)[12], double (&nR)[9],
                                           double (&pBr)[3])
{
    // ---- world_H_root from the stance sole (evaluateWorldToBaseTransformation), task frames in the world
    {
        double Rf[2][9], pf[2][3];
#pragma unroll
        for (int f = 0; f < 2; f++) {
            const int s = M.tpi_slot[M.frame_link[f]];
            double Rk[9], pk[3];
#pragma unroll
            for (int k = 0; k < 9; k++) Rk[k] = SLOT(s, k);
#pragma unroll
            for (int k = 0; k < 3; k++) pk[k] = SLOT(s, 9 + k);
            xf_mul(Rk, pk, M.frame_R[f], M.frame_p[f], Rf[f], pf[f]); // root_H_sole
            if ((f == 0) == left) {
#pragma unroll
                for (int k = 0; k < 3; k++) pBr[k] = pk[k]; // origin of the base link (= the stance sole's link) in the root frame
            }
        }
        double Rs[9], ps3[3];
#pragma unroll
        for (int k = 0; k < 9; k++) Rs[k] = left ? Rf[0][k] : Rf[1][k];
#pragma unroll
        for (int k = 0; k < 3; k++) ps3[k] = left ? pf[0][k] : pf[1][k];
        double Ri[9], pi[3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) Ri[3 * r + c] = Rs[3 * c + r];
#pragma unroll
        for (int r = 0; r < 3; r++) pi[r] = -(Ri[3 * r] * ps3[0] + Ri[3 * r + 1] * ps3[1] + Ri[3 * r + 2] * ps3[2]);
        const double* bH = (io.base_H ? io.base_H : (left ? io.lf_H_des : io.rf_H_des)) + i;
        double bR[9], bp[3];
        bR[0] = bH[0 * B]; bR[1] = bH[1 * B]; bR[2] = bH[2 * B]; bp[0] = bH[3 * B];
        bR[3] = bH[4 * B]; bR[4] = bH[5 * B]; bR[5] = bH[6 * B]; bp[1] = bH[7 * B];
        bR[6] = bH[8 * B]; bR[7] = bH[9 * B]; bR[8] = bH[10 * B]; bp[2] = bH[11 * B];
        xf_mul(bR, bp, Ri, pi, wR, wp);
        double Rw[9], pw[3];
        xf_mul(wR, wp, Rf[0], pf[0], Rw, pw);
        lfH[0] = Rw[0]; lfH[1] = Rw[1]; lfH[2] = Rw[2]; lfH[3] = pw[0]; lfH[4] = Rw[3]; lfH[5] = Rw[4]; lfH[6] = Rw[5]; lfH[7] = pw[1];
        lfH[8] = Rw[6]; lfH[9] = Rw[7]; lfH[10] = Rw[8]; lfH[11] = pw[2];
        xf_mul(wR, wp, Rf[1], pf[1], Rw, pw);
        rfH[0] = Rw[0]; rfH[1] = Rw[1]; rfH[2] = Rw[2]; rfH[3] = pw[0]; rfH[4] = Rw[3]; rfH[5] = Rw[4]; rfH[6] = Rw[5]; rfH[7] = pw[1];
        rfH[8] = Rw[6]; rfH[9] = Rw[7]; rfH[10] = Rw[8]; rfH[11] = pw[2];
        const int sn = M.tpi_slot[M.frame_link[2]];
        double Rk[9], Rn[9];
#pragma unroll
        for (int k = 0; k < 9; k++) Rk[k] = SLOT(sn, k);
        mat3_mul(Rk, M.frame_R[2], Rn);
        mat3_mul(wR, Rn, nR);
    }
}
#undef SLOT

// ---- kinematics + QP columns + the equality-constrained optimum of every instance; then the violators of the CTA, compacted
// onto its first threads, are solved once more with the bound Goldfarb-Idnani adds first fixed at its value
// PART 0: the whole tick.  PART 1 / PART 2: the same tick as two launches (wc_ik_tick_part_dev) -- 1 = everything that depends on the
// joint state, the base and the desired feet / neck only (kinematics, QP columns, the task right-hand sides except the CoM term):
// it can run before the desired CoM of the tick exists, i.e. beside the MPC block; 2 = the CoM term, the solves, the outputs.
// The arithmetic is that of PART 0, operation for operation.
template <int PART>
__global__ void __launch_bounds__(kTpiThreads, WC_TPI_MINB) ik_tpi_kernel(IkDev d, IkTickIO io)
{
    extern __shared__ double sm_slots[]; // [warp][slot][12][32]  (PART 2: none)
    __shared__ int s_n, s_i[kTpiThreads], s_id[kTpiThreads];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool live = blockIdx.x * kTpiThreads + threadIdx.x < d.B;
    const int i = live ? blockIdx.x * kTpiThreads + threadIdx.x : d.B - 1; // idle threads of the last CTA shadow its last instance
    // model table: one copy per CTA (the link loop reads ~25 words of it per link, the same for every lane), fetched by ONE bulk
    // asynchronous copy (cp.async.bulk, the copy engine's 1-D form: SASS UBLKCP) that completes on an mbarrier -- the CTA's
    // threads issue no loads for it and meet the data at the barrier wait below
    constexpr unsigned kModelBytes = (unsigned)((sizeof(IkModelDev) + 15) / 16 * 16);
    __shared__ __align__(16) unsigned long long s_model[kModelBytes / 8];
    __shared__ __align__(8) unsigned long long s_bar;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
    if (threadIdx.x == 0) {
        s_n = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(kModelBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"((unsigned)__cvta_generic_to_shared(s_model)), "l"(d.model), "r"(kModelBytes), "r"(bar) : "memory");
    }
    __syncthreads(); // the barrier is initialised (and s_n is zero) before anybody waits on it
    {
        unsigned ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(bar) : "memory");
        } while (!ok);
    }
    const size_t B = (size_t)d.B;
    const IkModelDev& M = *reinterpret_cast<const IkModelDev*>(s_model);
    const wc_ik_config& cfg = d.cfg;
    const int L = M.n_links, nJ = d.nJ;
    const TpiRows rows(L, nJ);
    int rc = 0, id = 0;
    if (live) {
    double* sc = d.tpi_scratch + i;
    const bool left = io.left_is_fixed[i] != 0;
    double br[9], es[3];
    bool fallback = false;
    if constexpr (PART != 2) {
    double* slots = sm_slots + (size_t)warp * d.tpi_slots * 12 * 32 + lane;
    // the measured joint positions are needed after the link loop: ask for their lines now (lane k: row k)
    if (lane < nJ && (lane & 1) * 0 == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(io.q_meas + (size_t)lane * B + (i - lane)));
    if (lane < nJ && (size_t)(i - lane) + 16 < B) asm volatile("prefetch.global.L2 [%0];" ::"l"(io.q_meas + (size_t)lane * B + (i - lane) + 16));
    double P[3];
    tpi_link_loop<true>(M, io, i, B, slots, sc, rows, P);
    double wR[9], wp[3];
    double lfH[12], rfH[12], nR[9], pBr[3];
    tpi_frames(M, io, i, B, slots, left, wR, wp, lfH, rfH, nR, pBr);
    const double Mtot = M.total_mass;
    double Mc[3], com[3], pB[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        Mc[r] = wR[3 * r] * P[0] + wR[3 * r + 1] * P[1] + wR[3 * r + 2] * P[2] + Mtot * wp[r];
        com[r] = Mc[r] / Mtot;
        pB[r] = wR[3 * r] * pBr[0] + wR[3 * r + 1] * pBr[1] + wR[3 * r + 2] * pBr[2] + wp[r];
    }
    // ---- vectors of the QP (same expressions as qp_vectors<true>): right-hand side of the task rows, neck term
    const double kpos = d.kpos ? d.kpos[i] : cfg.k_pos_foot;
    const double katt = d.katt ? d.katt[i] : cfg.k_att_foot;
    double bvec[15];
#pragma unroll
    for (int f = 0; f < 2; f++) {
        const double* src = f == 0 ? io.lf_H_des : io.rf_H_des;
        const double* tws = f == 0 ? io.lf_twist : io.rf_twist;
        const double* Hc = f == 0 ? lfH : rfH;
        double Hd[12], tw[6], Rc[9], Rd[9], e[3];
#pragma unroll
        for (int k = 0; k < 12; k++) Hd[k] = src[(size_t)k * B + i];
#pragma unroll
        for (int k = 0; k < 6; k++) tw[k] = tws[(size_t)k * B + i];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) { Rc[3 * r + c] = Hc[4 * r + c]; Rd[3 * r + c] = Hd[4 * r + c]; }
        rot_error(Rc, Rd, e);
#pragma unroll
        for (int r = 0; r < 3; r++) {
            bvec[6 * f + r] = tw[r] - kpos * (Hc[4 * r + 3] - Hd[4 * r + 3]);
            bvec[6 * f + 3 + r] = tw[3 + r] - katt * e[r];
        }
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
        bvec[12 + r] = PART == 1 ? 0.0 // (the solve part adds the CoM term)
                                 : io.com_vel_des[(size_t)r * B + i] - cfg.k_pos_com * (io.com_meas[(size_t)r * B + i] - io.com_pos_des[(size_t)r * B + i]);
    double We[3];
    {
        double Rdes[9], nd[9], en[3];
#pragma unroll
        for (int k = 0; k < 9; k++) nd[k] = io.neck_R_des[(size_t)k * B + i];
        mat3_mul(nd, cfg.additional_rotation, Rdes); // setDesiredNeckOrientation (QPIK.cpp:143-146)
        rot_error(nR, Rdes, en);
#pragma unroll
        for (int r = 0; r < 3; r++)
            We[r] = cfg.neck_weight[3 * r] * (1.0 * (-cfg.k_neck * en[0])) + cfg.neck_weight[3 * r + 1] * (1.0 * (-cfg.k_neck * en[1])) +
                    cfg.neck_weight[3 * r + 2] * (1.0 * (-cfg.k_neck * en[2]));
    }
    // ---- base twist from the stance-foot rows: v_lin + w x r_s = b_lin, w = b_ang; right-hand side of the remaining rows
    double bs[6], bw[6];
#pragma unroll
    for (int r = 0; r < 6; r++) { bs[r] = left ? bvec[r] : bvec[6 + r]; bw[r] = left ? bvec[6 + r] : bvec[r]; }
    const double pS[3] = {left ? lfH[3] : rfH[3], left ? lfH[7] : rfH[7], left ? lfH[11] : rfH[11]};
    const double pW[3] = {left ? rfH[3] : lfH[3], left ? rfH[7] : lfH[7], left ? rfH[11] : lfH[11]};
    const double om[3] = {bs[3], bs[4], bs[5]};
    double rS[3], rW[3], rC[3], t3[3], vlin[3];
#pragma unroll
    for (int r = 0; r < 3; r++) { rS[r] = pS[r] - pB[r]; rW[r] = pW[r] - pB[r]; rC[r] = com[r] - pB[r]; }
    cross3(om, rS, t3);
#pragma unroll
    for (int r = 0; r < 3; r++) { vlin[r] = bs[r] - t3[r]; es[r] = bs[r] - (vlin[r] + t3[r]); }
    cross3(om, rW, t3);
#pragma unroll
    for (int r = 0; r < 3; r++) { br[r] = bw[r] - (vlin[r] + t3[r]); br[3 + r] = bw[3 + r] - om[r]; }
    cross3(om, rC, t3);
#pragma unroll
    for (int r = 0; r < 3; r++) br[6 + r] = PART == 1 ? (vlin[r] + t3[r]) : bvec[12 + r] - (vlin[r] + t3[r]);
    const double* Wn = cfg.neck_weight;
    double Wom[3];
#pragma unroll
    for (int r = 0; r < 3; r++) Wom[r] = Wn[3 * r] * om[0] + Wn[3 * r + 1] * om[1] + Wn[3 * r + 2] * om[2];
#pragma unroll
    for (int r = 0; r < 9; r++) sc[(size_t)(rows.oINST + r) * B] = br[r];
#pragma unroll
    for (int r = 0; r < 3; r++) sc[(size_t)(rows.oINST + 9 + r) * B] = es[r];
    // ---- QP column of every joint (MIXED representation, WalkingForwardKinematics.cpp:31,436-454): into the scratch (in place of
    // the joint's root-frame axis and origin) and straight into the accumulators of the reduced problem
    const double Minv = 1.0 / Mtot;
    // rows of joint j the column needs: axis / origin (6), two prefix entries (3 + 3), the measured joint position
    auto request = [&](int jn, double (&nx)[13]) {
        const int cn = M.link_of_dof[jn], ln = M.tpi_sub_last[jn];
#pragma unroll
        for (int r = 0; r < 6; r++) nx[r] = sc[(size_t)(rows.oJ + 10 * jn + r) * B];
#pragma unroll
        for (int r = 0; r < 3; r++) { nx[6 + r] = sc[(size_t)(rows.oP + 3 * (ln + 1) + r) * B]; nx[9 + r] = sc[(size_t)(rows.oP + 3 * cn + r) * B]; }
        nx[12] = io.q_meas[(size_t)jn * B + i];
    };
    auto column = [&](int j, const double (&nx)[13]) {
        const int c = M.link_of_dof[j];
        double* cj = sc + (size_t)(rows.oJ + 10 * j) * B;
        double ar[3], orr[3], Sr[3];
#pragma unroll
        for (int r = 0; r < 3; r++) { ar[r] = nx[r]; orr[r] = nx[3 + r]; Sr[r] = nx[6 + r] - nx[9 + r]; }
        const double qm = nx[12];
        const double ms = M.sub_mass[c];
        double a3[3], o[3], S[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            a3[r] = wR[3 * r] * ar[0] + wR[3 * r + 1] * ar[1] + wR[3 * r + 2] * ar[2];
            o[r] = wR[3 * r] * orr[0] + wR[3 * r + 1] * orr[1] + wR[3 * r + 2] * orr[2] + wp[r];
            S[r] = wR[3 * r] * Sr[0] + wR[3 * r + 1] * Sr[1] + wR[3 * r + 2] * Sr[2] + ms * wp[r];
        }
        double ssw, sn;
        tpi_signs(M, left, j, ssw, sn);
        const bool bis = (left ? M.base_in_sub[1][j] : M.base_in_sub[0][j]) != 0;
        double r3[3], lin[3], mom[3], cc[3];
#pragma unroll
        for (int r = 0; r < 3; r++) r3[r] = pW[r] - o[r];
        cross3(a3, r3, lin);
#pragma unroll
        for (int r = 0; r < 3; r++)
            mom[r] = (bis ? -((Mc[r] - S[r]) - (Mtot - ms) * o[r]) : (S[r] - ms * o[r])) * Minv;
        cross3(a3, mom, cc);
        const double u[3] = {sn * a3[0], sn * a3[1], sn * a3[2]};
        const double lam = cfg.joint_reg_weight[j];
        if (!(lam > 0.0)) fallback = true;
        double gj = -(u[0] * We[0] + u[1] * We[1] + u[2] * We[2]);
        gj -= lam * (cfg.joint_reg_gain[j] * (cfg.joint_reg_rad[j] - qm));
        const double gr = gj + (u[0] * Wom[0] + u[1] * Wom[1] + u[2] * Wom[2]);
        const double a[9] = {ssw * lin[0], ssw * lin[1], ssw * lin[2], ssw * a3[0], ssw * a3[1], ssw * a3[2], cc[0], cc[1], cc[2]};
#pragma unroll
        for (int r = 0; r < 3; r++) {
            cj[(size_t)r * B] = a[r];
            cj[(size_t)(3 + r) * B] = a3[r];
            cj[(size_t)(6 + r) * B] = cc[r];
        }
        cj[(size_t)9 * B] = gr;
    };
    {
        // three rotating buffers (loop unrolled by three): the rows of joint j + 2 are requested before joint j is used -- half of
        // these rows come back from HBM, not from L2 (the scratch of the instances in flight is ~100 MB)
        double bA[13], bB[13], bC[13];
        request(0, bA);
        if (nJ > 1) request(1, bB);
#pragma unroll 1
        for (int j = 0; j < nJ; j += 3) {
            if (j + 2 < nJ) request(j + 2, bC);
            column(j, bA);
            if (j + 1 < nJ) {
                if (j + 3 < nJ) request(j + 3, bA);
                column(j + 1, bB);
            }
            if (j + 2 < nJ) {
                if (j + 4 < nJ) request(j + 4, bB);
                column(j + 2, bC);
            }
        }
    }
    if constexpr (PART == 1) {
        sc[(size_t)(rows.oINST + 12) * B] = fallback ? 1.0 : 0.0;
        if (io.o_com) { // the CoM the kinematics saw (what wc_fk_* would return for the same posture)
#pragma unroll
            for (int r = 0; r < 3; r++) io.o_com[(size_t)r * B + i] = com[r];
        }
    }
    } else { // PART 2: what the kinematics part left in the scratch, plus the CoM term of this tick
#pragma unroll
        for (int r = 0; r < 9; r++) br[r] = sc[(size_t)(rows.oINST + r) * B];
#pragma unroll
        for (int r = 0; r < 3; r++) es[r] = sc[(size_t)(rows.oINST + 9 + r) * B];
        fallback = sc[(size_t)(rows.oINST + 12) * B] != 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double b12 = io.com_vel_des[(size_t)r * B + i] - cfg.k_pos_com * (io.com_meas[(size_t)r * B + i] - io.com_pos_des[(size_t)r * B + i]);
            br[6 + r] = b12 - br[6 + r];
            sc[(size_t)(rows.oINST + 6 + r) * B] = br[6 + r]; // the second solve reads b_r from the scratch
        }
    }
    if constexpr (PART != 1) {
    TpiAcc A;
    rc = fallback ? 2 : tpi_solve<false>(d, io, M, A, sc + (size_t)rows.oJ * B, B, i, nJ, br, es, left, -1, 0.0, &id);
    if (rc == 1) { // to the second solve, with the selected bound
        const int k = atomicAdd(&s_n, 1);
        s_i[k] = i; s_id[k] = id;
    } else if (rc == 2) {
        d.slow_list[atomicAdd(d.slow_count, 1)] = i;
    }
    }
    } // live
    if constexpr (PART == 1) return;
    __syncthreads(); // also orders the scratch writes of the CTA's threads before the reads below
    const int n2 = s_n;
    if (threadIdx.x == 0 && n2 > 0) atomicAdd(d.slow_count + 1, n2); // statistics: instances that took the second solve
    if ((int)threadIdx.x < n2) {
        const int i2 = s_i[threadIdx.x], id2 = s_id[threadIdx.x];
        const double* sc2 = d.tpi_scratch + i2;
        double br[9], es[3];
#pragma unroll
        for (int r = 0; r < 9; r++) br[r] = sc2[(size_t)(rows.oINST + r) * B];
#pragma unroll
        for (int r = 0; r < 3; r++) es[r] = sc2[(size_t)(rows.oINST + 9 + r) * B];
        const bool left2 = io.left_is_fixed[i2] != 0;
        const int pvar = (id2 >> 1) - 6;
        const double beta = (id2 & 1) ? cfg.max_vel[pvar] : cfg.min_vel[pvar];
        TpiAcc A2;
        int id3 = 0;
        const int rc2 = tpi_solve<true>(d, io, M, A2, sc2 + (size_t)rows.oJ * B, B, i2, nJ, br, es, left2, pvar, beta, &id3);
        if (rc2 != 0) d.slow_list[atomicAdd(d.slow_count, 1)] = i2; // a second working-set change (or a fallback): general kernel
    }
#ifdef WC_TPI_DISCARD // measured: 0.203 -> 0.206 ms (the write-backs it saves are not what the kernel waits for); left as a build switch
    // The scratch of the CTA's instances is dead from here on.  Left alone its lines stay in L2 as dirty data, are written back to
    // HBM one by one (half of the kernel's DRAM traffic) and crowd out the scratch of the instances still in flight; discarding them
    // (`discard.global.L2`: the line is dropped, not written back) leaves L2 to the live rows.  Only whole lines that belong to one
    // warp's instances are discarded: rows are 128-byte aligned when B is a multiple of 16, a full warp's 32 instances are two lines.
    __syncthreads(); // the second solves of the CTA have read what they need
    if ((B & 15) == 0) {
        const size_t i0 = (size_t)blockIdx.x * kTpiThreads + (size_t)warp * 32; // first instance of this warp
        if (i0 + 32 <= B) {
            const double* base = d.tpi_scratch + i0;
            for (int r = lane; r < 2 * rows.total; r += 32) {
                const double* ln = base + (size_t)(r >> 1) * B + (size_t)(r & 1) * 16;
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(ln) : "memory");
            }
        }
    }
#endif
}


// ---- forward kinematics only (wc_fk_* without Jacobians: the measured CoM / feet of the closed loop), one thread per instance
__global__ void __launch_bounds__(kTpiThreads, WC_TPI_MINB) ik_tpi_fk_kernel(IkDev d, IkTickIO io)
{
    extern __shared__ double sm_slots[]; // [warp][slot][12][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr unsigned kModelBytes = (unsigned)((sizeof(IkModelDev) + 15) / 16 * 16);
    __shared__ __align__(16) unsigned long long s_model[kModelBytes / 8];
    {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(d.model);
        for (int k = threadIdx.x; k < (int)(kModelBytes / 8); k += kTpiThreads) s_model[k] = src[k];
    }
    __syncthreads();
    const int i = blockIdx.x * kTpiThreads + threadIdx.x;
    if (i >= d.B) return;
    const size_t B = (size_t)d.B;
    const IkModelDev& M = *reinterpret_cast<const IkModelDev*>(s_model);
    double* slots = sm_slots + (size_t)warp * d.tpi_slots * 12 * 32 + lane;
    const TpiRows rows(M.n_links, d.nJ);
    double P[3];
    tpi_link_loop<false>(M, io, i, B, slots, nullptr, rows, P);
    const bool left = io.left_is_fixed[i] != 0;
    double wR[9], wp[3], lfH[12], rfH[12], nR[9], pBr[3];
    tpi_frames(M, io, i, B, slots, left, wR, wp, lfH, rfH, nR, pBr);
    if (io.o_lf_H) {
#pragma unroll
        for (int k = 0; k < 12; k++) io.o_lf_H[(size_t)k * B + i] = lfH[k];
    }
    if (io.o_rf_H) {
#pragma unroll
        for (int k = 0; k < 12; k++) io.o_rf_H[(size_t)k * B + i] = rfH[k];
    }
    if (io.o_neck_R) {
#pragma unroll
        for (int k = 0; k < 9; k++) io.o_neck_R[(size_t)k * B + i] = nR[k];
    }
    if (io.o_com) {
        const double Mtot = M.total_mass;
#pragma unroll
        for (int r = 0; r < 3; r++) io.o_com[(size_t)r * B + i] = (wR[3 * r] * P[0] + wR[3 * r + 1] * P[1] + wR[3 * r + 2] * P[2] + Mtot * wp[r]) / Mtot;
    }
}

} // namespace

int ik_tpi_rows(int n_links, int n_dofs) { return TpiRows(n_links, n_dofs).total; }

// part 0: the whole tick in one launch; 1: its kinematics part; 2: its solve part (ik_tpi_kernel<PART>)
cudaError_t launch_ik_tpi(const IkDev& d, const IkTickIO& io, cudaStream_t s, int part)
{
    const size_t smem = part == 2 ? 0 : (size_t)(kTpiThreads / 32) * d.tpi_slots * 12 * 32 * sizeof(double);
    auto k = part == 0 ? ik_tpi_kernel<0> : (part == 1 ? ik_tpi_kernel<1> : ik_tpi_kernel<2>);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int blocks = (d.B + kTpiThreads - 1) / kTpiThreads;
    k<<<blocks, kTpiThreads, smem, s>>>(d, io);
    return cudaGetLastError();
}

cudaError_t launch_ik_tpi_fk(const IkDev& d, const IkTickIO& io, cudaStream_t s)
{
    const size_t smem = (size_t)(kTpiThreads / 32) * d.tpi_slots * 12 * 32 * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(ik_tpi_fk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ik_tpi_fk_kernel<<<(d.B + kTpiThreads - 1) / kTpiThreads, kTpiThreads, smem, s>>>(d, io);
    return cudaGetLastError();
}

} // namespace wcb
