/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
 *
 * kin_ref.c: CPU restatement of the kinematics the reference obtains from
 * iDynTree::KinDynComputations through WalkingFK
 * (modules/Walking_module/src/WalkingForwardKinematics.cpp:217-276 base selection,
 * :258-276 setInternalRobotState, :411-454 transform / Jacobian getters):
 *  - floating base = the stance-foot LINK, world_H_base = world_H_sole(desired) * sole_H_link
 *    (:217-256), base twist zero;
 *  - frame velocity representation MIXED (:31): linear velocity of the frame origin and
 *    angular velocity, both in world axes; free-floating Jacobians 6 x (6+nJ), columns
 *    0-5 the base (mixed) twist, columns 6.. the joints in model order;
 *  - centre-of-mass position and 3 x (6+nJ) Jacobian.
 * iDynTree is third-party and absent from /root/reference, and so is the iCub URDF; the
 * tree is given as a wc_model table.  Written link-by-link with explicit ancestor tests
 * (O(L^2)) so that it shares no structure with the GPU implementation.
 * PARITY UNPINNED; validated by finite differences in tests/test_oracle_kin.py.
 */
#include "oracle_internal.h"
#include "wc_oracle.h"

#include <math.h>
#include <string.h>

typedef struct { double R[9]; double p[3]; } xf;

static void xf_mul(const xf* a, const xf* b, xf* o)
{
    xf t;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++)
            t.R[3 * r + c] = a->R[3 * r] * b->R[c] + a->R[3 * r + 1] * b->R[3 + c] + a->R[3 * r + 2] * b->R[6 + c];
        t.p[r] = a->R[3 * r] * b->p[0] + a->R[3 * r + 1] * b->p[1] + a->R[3 * r + 2] * b->p[2] + a->p[r];
    }
    *o = t;
}
static void xf_inv(const xf* a, xf* o)
{
    xf t;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) t.R[3 * r + c] = a->R[3 * c + r];
    for (int r = 0; r < 3; r++) t.p[r] = -(t.R[3 * r] * a->p[0] + t.R[3 * r + 1] * a->p[1] + t.R[3 * r + 2] * a->p[2]);
    *o = t;
}
static void axis_angle(const double* ax, double q, double* R)
{
    const double c = cos(q), s = sin(q), v = 1.0 - c;
    const double x = ax[0], y = ax[1], z = ax[2];
    R[0] = c + x * x * v;     R[1] = x * y * v - z * s; R[2] = x * z * v + y * s;
    R[3] = y * x * v + z * s; R[4] = c + y * y * v;     R[5] = y * z * v - x * s;
    R[6] = z * x * v - y * s; R[7] = z * y * v + x * s; R[8] = c + z * z * v;
}
static void from_H12(const double* H, xf* o)
{
    o->R[0] = H[0]; o->R[1] = H[1]; o->R[2] = H[2]; o->p[0] = H[3];
    o->R[3] = H[4]; o->R[4] = H[5]; o->R[5] = H[6]; o->p[1] = H[7];
    o->R[6] = H[8]; o->R[7] = H[9]; o->R[8] = H[10]; o->p[2] = H[11];
}
static void to_H12(const xf* a, double* H)
{
    H[0] = a->R[0]; H[1] = a->R[1]; H[2] = a->R[2]; H[3] = a->p[0];
    H[4] = a->R[3]; H[5] = a->R[4]; H[6] = a->R[5]; H[7] = a->p[1];
    H[8] = a->R[6]; H[9] = a->R[7]; H[10] = a->R[8]; H[11] = a->p[2];
}
static int is_ancestor_or_self(const wc_model* m, int anc, int link)
{
    while (link >= 0) { if (link == anc) return 1; link = m->parent[link]; }
    return 0;
}
static void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* Jacobian (6 x nV row-major) of a point/frame rigidly attached to `link` at world position pF */
static void frame_jacobian(const wc_model* m, const xf* W, int base, int link, const double* pF, double* J)
{
    const int nV = m->n_dofs + 6;
    memset(J, 0, sizeof(double) * 6 * (size_t)nV);
    const double* pB = W[base].p;
    const double r[3] = {pF[0] - pB[0], pF[1] - pB[1], pF[2] - pB[2]};
    /* base columns: [I, -S(r); 0, I] */
    J[0 * nV + 0] = 1; J[1 * nV + 1] = 1; J[2 * nV + 2] = 1;
    J[0 * nV + 4] = r[2];  J[0 * nV + 5] = -r[1];
    J[1 * nV + 3] = -r[2]; J[1 * nV + 5] = r[0];
    J[2 * nV + 3] = r[1];  J[2 * nV + 4] = -r[0];
    J[3 * nV + 3] = 1; J[4 * nV + 4] = 1; J[5 * nV + 5] = 1;
    for (int c = 1; c < m->n_links; c++) {
        if (m->dof[c] < 0) continue;
        const int on_f = is_ancestor_or_self(m, c, link);
        const int on_b = is_ancestor_or_self(m, c, base);
        if (on_f == on_b) continue;
        const double sg = on_f ? 1.0 : -1.0;
        double aw[3], d[3], lin[3];
        for (int k = 0; k < 3; k++)
            aw[k] = W[c].R[3 * k] * m->axis[c][0] + W[c].R[3 * k + 1] * m->axis[c][1] + W[c].R[3 * k + 2] * m->axis[c][2];
        for (int k = 0; k < 3; k++) d[k] = pF[k] - W[c].p[k];
        cross3(aw, d, lin);
        const int col = 6 + m->dof[c];
        for (int k = 0; k < 3; k++) { J[k * nV + col] = sg * lin[k]; J[(3 + k) * nV + col] = sg * aw[k]; }
    }
}

/* All outputs row-major, single instance. Any output pointer may be NULL. */
void wco_fk(const wc_model* m, const double* q, const double* base_H12, int left_is_fixed,
            double* lf_H12, double* rf_H12, double* neck_R9, double* com3,
            double* J_lf, double* J_rf, double* J_neck, double* J_com)
{
    const int L = m->n_links, nV = m->n_dofs + 6;
    xf rootH[WC_MAX_LINKS], W[WC_MAX_LINKS];
    memset(&rootH[0], 0, sizeof(xf));
    rootH[0].R[0] = rootH[0].R[4] = rootH[0].R[8] = 1.0;
    for (int i = 1; i < L; i++) {
        xf j;
        memcpy(j.R, m->joint_R[i], sizeof(j.R));
        memcpy(j.p, m->joint_p[i], sizeof(j.p));
        if (m->dof[i] >= 0) {
            xf rot; memset(&rot, 0, sizeof(rot));
            axis_angle(m->axis[i], q[m->dof[i]], rot.R);
            xf_mul(&j, &rot, &j);
        }
        xf_mul(&rootH[m->parent[i]], &j, &rootH[i]);
    }
    /* world_H_root = world_H_sole * (root_H_footlink * link_H_sole)^-1 */
    const int base = left_is_fixed ? m->l_sole_link : m->r_sole_link;
    xf sole, rootHsole, inv, wHsole, wHroot;
    memcpy(sole.R, left_is_fixed ? m->l_sole_R : m->r_sole_R, sizeof(sole.R));
    memcpy(sole.p, left_is_fixed ? m->l_sole_p : m->r_sole_p, sizeof(sole.p));
    xf_mul(&rootH[base], &sole, &rootHsole);
    xf_inv(&rootHsole, &inv);
    from_H12(base_H12, &wHsole);
    xf_mul(&wHsole, &inv, &wHroot);
    for (int i = 0; i < L; i++) xf_mul(&wHroot, &rootH[i], &W[i]);

    xf fl, fr, fn, t;
    memcpy(t.R, m->l_sole_R, sizeof(t.R)); memcpy(t.p, m->l_sole_p, sizeof(t.p)); xf_mul(&W[m->l_sole_link], &t, &fl);
    memcpy(t.R, m->r_sole_R, sizeof(t.R)); memcpy(t.p, m->r_sole_p, sizeof(t.p)); xf_mul(&W[m->r_sole_link], &t, &fr);
    memcpy(t.R, m->neck_R, sizeof(t.R)); memcpy(t.p, m->neck_p, sizeof(t.p)); xf_mul(&W[m->neck_link], &t, &fn);
    if (lf_H12) to_H12(&fl, lf_H12);
    if (rf_H12) to_H12(&fr, rf_H12);
    if (neck_R9) memcpy(neck_R9, fn.R, sizeof(fn.R));

    double M = 0.0, c[3] = {0, 0, 0};
    double cw[WC_MAX_LINKS][3];
    for (int i = 0; i < L; i++) {
        for (int k = 0; k < 3; k++)
            cw[i][k] = W[i].R[3 * k] * m->com[i][0] + W[i].R[3 * k + 1] * m->com[i][1] + W[i].R[3 * k + 2] * m->com[i][2] + W[i].p[k];
        M += m->mass[i];
        for (int k = 0; k < 3; k++) c[k] += m->mass[i] * cw[i][k];
    }
    for (int k = 0; k < 3; k++) c[k] /= M;
    if (com3) memcpy(com3, c, sizeof(c));

    if (J_lf) frame_jacobian(m, W, base, m->l_sole_link, fl.p, J_lf);
    if (J_rf) frame_jacobian(m, W, base, m->r_sole_link, fr.p, J_rf);
    if (J_neck) frame_jacobian(m, W, base, m->neck_link, fn.p, J_neck);
    if (J_com) {
        /* mass-weighted mean of the linear Jacobians of the link centres of mass */
        double Jl[6 * (WC_MAX_DOFS + 6)];
        memset(J_com, 0, sizeof(double) * 3 * (size_t)nV);
        for (int i = 0; i < L; i++) {
            if (m->mass[i] == 0.0) continue;
            frame_jacobian(m, W, base, i, cw[i], Jl);
            for (int k = 0; k < 3 * nV; k++) J_com[k] += (m->mass[i] / M) * Jl[k];
        }
    }
}
