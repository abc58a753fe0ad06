/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 * CPU restatement of the closed-loop glue of one control tick:
 *   StableDCMModel::integrateModel          modules/Walking_module/src/StableDCMModel.cpp:63-90
 *   WalkingZMPController::evaluateControl   modules/Walking_module/src/WalkingZMPController.cpp:144-172
 *   joint integration                       modules/Walking_module/src/WalkingModule.cpp:741-744
 * iCub::ctrl::Integrator (external, iCub ctrlLib, not vendored) is restated from its published rule: Tustin
 * integration y_k = y_{k-1} + Ts/2 (x_k + x_{k-1}), reset(y0) sets y = y0 and x_old = 0.
 * PARITY UNPINNED (no reference tests / golden vectors); checked against closed forms in tests/test_glue.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "wc_oracle.h"

struct wco_glue {
    wc_glue_config cfg;
    int B;
    double omega;
    double *lipm_c, *lipm_v, *zc_y, *zc_v, *q_y, *q_v;
};

wco_glue* wco_glue_create(const wc_glue_config* cfg, int B)
{
    wco_glue* g = (wco_glue*)calloc(1, sizeof(*g));
    g->cfg = *cfg; g->B = B; g->omega = sqrt(cfg->gravity / cfg->com_height);
    const size_t n2 = 2 * (size_t)B, nq = (size_t)cfg->n_dofs * (size_t)B;
    g->lipm_c = (double*)calloc(n2, sizeof(double)); g->lipm_v = (double*)calloc(n2, sizeof(double));
    g->zc_y = (double*)calloc(n2, sizeof(double)); g->zc_v = (double*)calloc(n2, sizeof(double));
    g->q_y = (double*)calloc(nq, sizeof(double)); g->q_v = (double*)calloc(nq, sizeof(double));
    return g;
}

void wco_glue_free(wco_glue* g)
{
    if (!g) return;
    free(g->lipm_c); free(g->lipm_v); free(g->zc_y); free(g->zc_v); free(g->q_y); free(g->q_v); free(g);
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

void wco_glue_com_reference(wco_glue* g, const double* dcm_des, const double* zmp_des, const double* zmp_meas,
                            const double* com_meas, const double* com_z_des, const double* com_z_vel_des,
                            const double* k_zmp, const double* k_com, double* com_pos_des, double* com_vel_des)
{
    const size_t B = (size_t)g->B;
    const double ts = g->cfg.sampling_time;
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
