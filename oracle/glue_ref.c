/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 * CPU restatement of the closed-loop glue of one control tick:
 *   StableDCMModel::integrateModel          modules/Walking_module/src/StableDCMModel.cpp:63-90
 *   WalkingZMPController::evaluateControl   modules/Walking_module/src/WalkingZMPController.cpp:144-172
 *   joint integration                       modules/Walking_module/src/WalkingModule.cpp:741-744
 *   WalkingZMPController::setPhase          modules/Walking_module/src/WalkingZMPController.cpp:107-126 (gain scheduling;
 *     iCub::ctrl::minJerkTrajGen is external (iCub ctrlLib, not vendored): restated as the Tustin discretisation of its
 *     third-order filter F(s) = -a / (s^3 - c s^2 - b s - a), a = -150.765868956161 / T^3, b = -84.9812819469538 / T^2,
 *     c = -15.9669610709384 / T, run as ctrlLib's Filter does (direct form, init(y0) = steady state at y0))
 * iCub::ctrl::Integrator (external, iCub ctrlLib, not vendored) is restated from its published rule: Tustin
 * integration y_k = y_{k-1} + Ts/2 (x_k + x_{k-1}), reset(y0) sets y = y0 and x_old = 0.
 * and (rank 4):
 *   WalkingDCMReactiveController::evaluateControl   modules/Walking_module/src/WalkingDCMReactiveController.cpp:63-81
 *   WalkingModule::evaluateZMP                       modules/Walking_module/src/WalkingModule.cpp:824-876
 *   WalkingFK::evaluateDCM                           modules/Walking_module/src/WalkingForwardKinematics.cpp:312-340
 *   the logger row of updateModule                   modules/Walking_module/src/WalkingModule.cpp:793-803, 1231-1250
 *     (iDynTree::Rotation::asRPY is external: restated from its documented convention R = Rz(yaw) Ry(pitch) Rx(roll))
 * PARITY UNPINNED (no reference tests / golden vectors); checked against closed forms in tests/test_glue.py.
 */
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#include <stdlib.h>
#include <string.h>

#include "wc_oracle.h"

struct wco_glue {
    wc_glue_config cfg;
    int B;
    double omega;
    double *lipm_c, *lipm_v, *zc_y, *zc_v, *q_y, *q_v;
    double* cf; /* [12][B] CoM filters of WalkingFK: position y, u_old, velocity y, u_old */
    double* gs; /* [14][B] smoothers of kZMP and kCoM (y1..y3, u1..u3 each), current kZMP, kCoM */
    double fb[4], fa[4];
};

wco_glue* wco_glue_create(const wc_glue_config* cfg, int B)
{
    wco_glue* g = (wco_glue*)calloc(1, sizeof(*g));
    g->cfg = *cfg; g->B = B; g->omega = sqrt(cfg->gravity / cfg->com_height);
    const size_t n2 = 2 * (size_t)B, nq = (size_t)cfg->n_dofs * (size_t)B;
    g->lipm_c = (double*)calloc(n2, sizeof(double)); g->lipm_v = (double*)calloc(n2, sizeof(double));
    g->zc_y = (double*)calloc(n2, sizeof(double)); g->zc_v = (double*)calloc(n2, sizeof(double));
    g->q_y = (double*)calloc(nq, sizeof(double)); g->q_v = (double*)calloc(nq, sizeof(double));
    g->cf = (double*)calloc(12 * (size_t)B, sizeof(double));
    for (int i = 0; i < B; i++) { g->cf[2 * (size_t)B + i] = cfg->com_height; g->cf[5 * (size_t)B + i] = cfg->com_height; }
    g->gs = (double*)calloc(14 * (size_t)B, sizeof(double));
    if (cfg->use_gain_scheduling) {
        /* minJerkTrajGen(1, Ts, T)::computeCoeffs + init(k*_stance) (WalkingZMPController.cpp:63-83) */
        const double T = cfg->smoothing_time, k = 2.0 / cfg->sampling_time;
        const double a = -150.765868956161 / (T * T * T), b = -84.9812819469538 / (T * T), c = -15.9669610709384 / T;
        const double k2 = k * k, k3 = k2 * k;
        g->fa[0] = k3 - c * k2 - b * k - a;  g->fa[1] = -3.0 * k3 + c * k2 - b * k - 3.0 * a;
        g->fa[2] = 3.0 * k3 + c * k2 + b * k - 3.0 * a;  g->fa[3] = -k3 - c * k2 + b * k - a;
        g->fb[0] = -a; g->fb[1] = -3.0 * a; g->fb[2] = -3.0 * a; g->fb[3] = -a;
        for (int i = 0; i < B; i++) {
            for (int r = 0; r < 6; r++) { g->gs[(size_t)r * B + i] = cfg->k_zmp_stance; g->gs[(size_t)(6 + r) * B + i] = cfg->k_com_stance; }
            g->gs[(size_t)12 * B + i] = cfg->k_zmp_stance; g->gs[(size_t)13 * B + i] = cfg->k_com_stance;
        }
    }
    return g;
}

void wco_glue_free(wco_glue* g)
{
    if (!g) return;
    free(g->lipm_c); free(g->lipm_v); free(g->zc_y); free(g->zc_v); free(g->q_y); free(g->q_v); free(g->cf); free(g->gs); free(g);
}

/* Integrator::reset for the three integrators (WalkingModule.cpp:470-471) */
void wco_glue_reset(wco_glue* g, const double* com_xy, const double* q)
{
    const size_t B = (size_t)g->B;
    if (com_xy) {
        memcpy(g->lipm_c, com_xy, 2 * B * sizeof(double)); memcpy(g->zc_y, com_xy, 2 * B * sizeof(double));
        memset(g->lipm_v, 0, 2 * B * sizeof(double)); memset(g->zc_v, 0, 2 * B * sizeof(double));
    }
    if (q) {
        memcpy(g->q_y, q, (size_t)g->cfg.n_dofs * B * sizeof(double));
        memset(g->q_v, 0, (size_t)g->cfg.n_dofs * B * sizeof(double));
    }
}

static double tustin(double* y, double* x_old, double x, double ts)
{
    *y = *y + (x + *x_old) * (ts / 2);
    *x_old = x;
    return *y;
}

/* WalkingZMPController::setPhase (:107-126): stance[i], or |dcm_vel_des| < 0.001 (WalkingModule.cpp:659-660) */
void wco_glue_set_phase(wco_glue* g, const double* dcm_vel_des, const unsigned char* stance)
{
    const size_t B = (size_t)g->B;
    if (!g->cfg.use_gain_scheduling) return;
    for (size_t i = 0; i < B; i++) {
        int st;
        if (stance) st = stance[i] != 0;
        else st = sqrt(dcm_vel_des[i] * dcm_vel_des[i] + dcm_vel_des[B + i] * dcm_vel_des[B + i]) < 0.001;
        for (int q = 0; q < 2; q++) {
            const double u = q == 0 ? (st ? g->cfg.k_zmp_stance : g->cfg.k_zmp) : (st ? g->cfg.k_com_stance : g->cfg.k_com);
            double* s = g->gs + (size_t)(6 * q) * B + i;
            /* ctrlLib Filter::filt: y = b0 u; y += b_k u_old[k-1]; y -= a_k y_old[k-1]; y = (1 / a0) y */
            double y = g->fb[0] * u;
            y += g->fb[1] * s[3 * B]; y += g->fb[2] * s[4 * B]; y += g->fb[3] * s[5 * B];
            y -= g->fa[1] * s[0]; y -= g->fa[2] * s[B]; y -= g->fa[3] * s[2 * B];
            y = (1.0 / g->fa[0]) * y;
            s[2 * B] = s[B]; s[B] = s[0]; s[0] = y;
            s[5 * B] = s[4 * B]; s[4 * B] = s[3 * B]; s[3 * B] = u;
            g->gs[(size_t)(12 + q) * B + i] = y;
        }
    }
}

void wco_glue_com_reference(wco_glue* g, const double* dcm_des, const double* zmp_des, const double* zmp_meas,
                            const double* com_meas, const double* com_z_des, const double* com_z_vel_des,
                            const double* k_zmp, const double* k_com, double* com_pos_des, double* com_vel_des)
{
    const size_t B = (size_t)g->B;
    const double ts = g->cfg.sampling_time;
    if (!k_zmp && g->cfg.use_gain_scheduling) k_zmp = g->gs + 12 * B;
    if (!k_com && g->cfg.use_gain_scheduling) k_com = g->gs + 13 * B;
    for (size_t i = 0; i < B; i++) {
        const double kz = k_zmp ? k_zmp[i] : g->cfg.k_zmp, kc = k_com ? k_com[i] : g->cfg.k_com;
        for (int a = 0; a < 2; a++) {
            const size_t k = (size_t)a * B + i;
            /* StableDCMModel.cpp:74-84 */
            const double v = -g->omega * (g->lipm_c[k] - dcm_des[k]);
            const double c = tustin(&g->lipm_c[k], &g->lipm_v[k], v, ts);
            /* WalkingZMPController.cpp:155-167 */
            const double vd = kc * (c - com_meas[k]) - kz * (zmp_des[k] - zmp_meas[k]) + v;
            com_pos_des[k] = tustin(&g->zc_y[k], &g->zc_v[k], vd, ts);
            com_vel_des[k] = vd;
        }
        com_pos_des[2 * B + i] = com_z_des ? com_z_des[i] : 0.0;
        com_vel_des[2 * B + i] = com_z_vel_des ? com_z_vel_des[i] : 0.0;
    }
}

void wco_glue_integrate_joints(wco_glue* g, const double* dq, double* q_des)
{
    const size_t n = (size_t)g->cfg.n_dofs * (size_t)g->B;
    for (size_t k = 0; k < n; k++) {
        const double y = tustin(&g->q_y[k], &g->q_v[k], dq[k], g->cfg.sampling_time);
        if (q_des) q_des[k] = y;
    }
}

/* WalkingDCMReactiveController.cpp:74-77 */
void wco_reactive_dcm(int B, double omega, double k_const, const double* k_dcm, const double* dcm_des,
                      const double* dcm_vel_des, const double* dcm_meas, double* zmp_des)
{
    for (int i = 0; i < B; i++) {
        const double k = k_dcm ? k_dcm[i] : k_const;
        for (int a = 0; a < 2; a++) {
            const size_t j = (size_t)a * B + i;
            zmp_des[j] = dcm_des[j] - 1 / omega * dcm_vel_des[j] - k * (dcm_des[j] - dcm_meas[j]);
        }
    }
}

/* WalkingForwardKinematics.cpp:329-335 (m_useFilters == false) */
void wco_evaluate_dcm(int B, double omega, const double* com, const double* com_vel, double* dcm)
{
    for (int i = 0; i < B; i++)
        for (int a = 0; a < 2; a++) dcm[(size_t)a * B + i] = com[(size_t)a * B + i] + com_vel[(size_t)a * B + i] / omega;
}

/* WalkingModule.cpp:824-876; wrenches [6][B] force then torque, transforms [12][B] row-major 3x4 */
void wco_evaluate_zmp(int B, const double* lw, const double* rw, const double* lf_H, const double* rf_H, double* zmp, int* ok)
{
    for (int i = 0; i < B; i++) {
        const size_t Bz = (size_t)B;
        double zl[3] = {0, 0, 0}, zr[3] = {0, 0, 0}, ld = 0.0, rd = 0.0;
        const double fzr = rw[2 * Bz + i], fzl = lw[2 * Bz + i];
        if (!(fzr < 0.001)) { zr[0] = -rw[4 * Bz + i] / fzr; zr[1] = rw[3 * Bz + i] / fzr; rd = 1.0; }
        if (!(fzl < 0.001)) { zl[0] = -lw[4 * Bz + i] / fzl; zl[1] = lw[3 * Bz + i] / fzl; ld = 1.0; }
        const double totalZ = fzr + fzl;
        zmp[i] = 0.0; zmp[Bz + i] = 0.0;
        if (totalZ < 0.1) { if (ok) ok[i] = 0; continue; }
        double wl[2], wr[2];
        for (int r = 0; r < 2; r++) {
            wl[r] = lf_H[(4 * r) * Bz + i] * zl[0] + lf_H[(4 * r + 1) * Bz + i] * zl[1] + lf_H[(4 * r + 2) * Bz + i] * zl[2] + lf_H[(4 * r + 3) * Bz + i];
            wr[r] = rf_H[(4 * r) * Bz + i] * zr[0] + rf_H[(4 * r + 1) * Bz + i] * zr[1] + rf_H[(4 * r + 2) * Bz + i] * zr[2] + rf_H[(4 * r + 3) * Bz + i];
        }
        for (int r = 0; r < 2; r++) zmp[(size_t)r * Bz + i] = ((fzl * ld) / totalZ) * wl[r] + ((fzr * rd) / totalZ) * wr[r];
        if (ok) ok[i] = 1;
    }
}

static void as_rpy(const double* H, size_t B, size_t i, double* rpy)
{
#define R_(r, c) H[(size_t)(4 * (r) + (c)) * B + i]
    if (R_(2, 0) < 1.0) {
        if (R_(2, 0) > -1.0) { rpy[0] = atan2(R_(2, 1), R_(2, 2)); rpy[1] = asin(-R_(2, 0)); rpy[2] = atan2(R_(1, 0), R_(0, 0)); }
        else { rpy[0] = 0.0; rpy[1] = M_PI / 2.0; rpy[2] = -atan2(-R_(1, 2), R_(1, 1)); }
    } else { rpy[0] = 0.0; rpy[1] = -M_PI / 2.0; rpy[2] = atan2(-R_(1, 2), R_(1, 1)); }
#undef R_
}

/* the 53 columns of WalkingModule.cpp:1231-1250 from the arguments of sendData (:793-803); out [53][B] */
void wco_log_pack(int B, const wc_log_inputs* in, double* out)
{
    const size_t Bz = (size_t)B;
    for (size_t i = 0; i < Bz; i++) {
        int c = 0;
#define PUT(v) do { out[(size_t)c * Bz + i] = (v); c++; } while (0)
#define PUTN(p, n) do { for (int k_ = 0; k_ < (n); k_++) PUT((p) ? (p)[(size_t)k_ * Bz + i] : 0.0); } while (0)
#define PUTPOSE(Hp) do { if (!(Hp)) { for (int k_ = 0; k_ < 6; k_++) PUT(0.0); } else { double rpy_[3]; \
        PUT((Hp)[3 * Bz + i]); PUT((Hp)[7 * Bz + i]); PUT((Hp)[11 * Bz + i]); as_rpy((Hp), Bz, i, rpy_); \
        PUT(rpy_[0]); PUT(rpy_[1]); PUT(rpy_[2]); } } while (0)
        PUTN(in->dcm, 2); PUTN(in->dcm_des, 2); PUTN(in->dcm_vel_des, 2); PUTN(in->zmp, 2); PUTN(in->zmp_des, 2);
        PUTN(in->com, 3); PUTN(in->com_des_xy, 2); PUTN(in->com_vel_des_xy, 2);
        PUTPOSE(in->lf_H); PUTPOSE(in->rf_H); PUTPOSE(in->lf_H_des); PUTPOSE(in->rf_H_des);
        PUTN(in->lf_err, 6); PUTN(in->rf_err, 6);
#undef PUT
#undef PUTN
#undef PUTPOSE
    }
}

/* WalkingFK::evaluateCoM / evaluateDCM (WalkingForwardKinematics.cpp:278-340) with iCub::ctrl::FirstOrderLowPassFilter (external
 * ctrlLib, filters.cpp: num = [Ts, Ts], den = [2 tau + Ts, Ts - 2 tau], tau = 1 / (2 pi fc); init(y0): y = y0, u_old = y0) */
void wco_glue_reset_com_filter(wco_glue* g)
{
    const size_t B = (size_t)g->B;
    memset(g->cf, 0, 12 * B * sizeof(double));
    for (size_t i = 0; i < B; i++) { g->cf[2 * B + i] = g->cfg.com_height; g->cf[5 * B + i] = g->cfg.com_height; }
}

void wco_glue_evaluate_com(wco_glue* g, const double* com, const double* J_com, const double* dq, double cut_frequency,
                           int use_filters, double* com_out, double* vel_out, double* dcm_out)
{
    const size_t B = (size_t)g->B;
    const int nJ = g->cfg.n_dofs, nV = nJ + 6;
    const double Ts = g->cfg.sampling_time;
    const int filt = cut_frequency > 0.0;
    double fa = 0.0, fb = 0.0;
    if (filt) {
        const double tau = 1.0 / (2.0 * M_PI * cut_frequency);
        fa = (Ts - 2.0 * tau) / (2.0 * tau + Ts); fb = Ts / (2.0 * tau + Ts);
    }
    for (size_t i = 0; i < B; i++) {
        double c[3], v[3] = {0, 0, 0}, cfil[3], vfil[3];
        for (int r = 0; r < 3; r++) c[r] = com[r * B + i];
        for (int j = 0; j < nJ; j++)
            for (int r = 0; r < 3; r++) v[r] += J_com[(size_t)(r * nV + 6 + j) * B + i] * dq[(size_t)j * B + i];
        for (int r = 0; r < 3; r++) {
            cfil[r] = c[r]; vfil[r] = v[r];
            if (filt) {
                double *py = g->cf + r * B + i, *pu = g->cf + (3 + r) * B + i, *vy = g->cf + (6 + r) * B + i, *vu = g->cf + (9 + r) * B + i;
                cfil[r] = (fb * (c[r] + *pu) - fa * *py); *py = cfil[r]; *pu = c[r];
                vfil[r] = (fb * (v[r] + *vu) - fa * *vy); *vy = vfil[r]; *vu = v[r];
            }
        }
        const double* co = (filt && use_filters) ? cfil : c;
        const double* vo = (filt && use_filters) ? vfil : v;
        for (int r = 0; r < 3; r++) { if (com_out) com_out[r * B + i] = co[r]; if (vel_out) vel_out[r * B + i] = vo[r]; }
        if (dcm_out) { dcm_out[i] = co[0] + vo[0] / g->omega; dcm_out[B + i] = co[1] + vo[1] / g->omega; }
    }
}
