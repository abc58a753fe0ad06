/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 * CPU twin of the synthetic input producer (SURVEY.md §8(d) workload, §8(f) rank 3): the signals WalkingModule reads from
 * TrajectoryGenerator (modules/Walking_module/src/TrajectoryGenerator.cpp:501-598) for periodic gaits with the shipped
 * planner parameters (app/robots/iCubGazeboV2_5/plannerParams.ini).  The reference's producer (UnicyclePlanner, external,
 * not vendored) is NOT restated; this file defines the synthetic workload independently of the CUDA implementation
 * (plain loops, libm), and tests/test_traj.py also checks it against the torch generator walking-controllers_b200/synthetic.py.
 * PARITY UNPINNED with respect to the reference (there is nothing in the reference to pin a synthetic workload to).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "wc_oracle.h"

struct wco_traj {
    wc_traj_config cfg;
    int B, S2;
    double omega;
    double *fx, *fy, *fyaw, *zx, *zy, *ex, *ey; /* [S2][B] */
    int32_t* phase; uint8_t* first_left;
};

wco_traj* wco_traj_create(const wc_traj_config* cfg, int B)
{
    wco_traj* g = (wco_traj*)calloc(1, sizeof(*g));
    g->cfg = *cfg; g->B = B; g->S2 = cfg->n_steps + 2; g->omega = sqrt(cfg->gravity / cfg->com_height);
    const size_t n = (size_t)g->S2 * (size_t)B;
    double** t[7] = {&g->fx, &g->fy, &g->fyaw, &g->zx, &g->zy, &g->ex, &g->ey};
    for (int k = 0; k < 7; k++) *t[k] = (double*)calloc(n, sizeof(double));
    g->phase = (int32_t*)calloc((size_t)B, sizeof(int32_t)); g->first_left = (uint8_t*)calloc((size_t)B, 1);
    return g;
}

void wco_traj_free(wco_traj* g)
{
    if (!g) return;
    free(g->fx); free(g->fy); free(g->fyaw); free(g->zx); free(g->zy); free(g->ex); free(g->ey); free(g->phase); free(g->first_left);
    free(g);
}

static double foot_sign(const wco_traj* g, int k, int i) { return (((k % 2) == 0) == (g->first_left[i] != 0)) ? 1.0 : -1.0; }
static int clampj(const wco_traj* g, int k) { int j = k + 2; if (j < 0) j = 0; if (j > g->S2 - 1) j = g->S2 - 1; return j; }
#define TB(arr, k) (g->arr[(size_t)clampj(g, (k)) * (size_t)g->B + i])

void wco_traj_set_gait(wco_traj* g, const int32_t* phase, const double* step_len, const double* dtheta, const double* theta0,
                       const uint8_t* first_left)
{
    const int B = g->B, S2 = g->S2;
    const double decay = exp(-g->omega * g->cfg.step_ticks * g->cfg.sampling_time);
    for (int i = 0; i < B; i++) {
        g->phase[i] = phase ? phase[i] : 0; g->first_left[i] = first_left ? first_left[i] : 1;
        const double L = step_len[i], dth = dtheta ? dtheta[i] : 0.0, th0 = theta0 ? theta0[i] : 0.0;
        double cx = 0.0, cy = 0.0;
        for (int j = 0; j < S2; j++) {
            const int k = j - 2;
            const double th = th0 + (double)k * dth, c = cos(th), s = sin(th);
            cx += L * c; cy += L * s;
            const double sg = foot_sign(g, k, i);
            const size_t o = (size_t)j * B + i;
            g->fx[o] = cx - 0.5 * g->cfg.nominal_width * sg * s;
            g->fy[o] = cy + 0.5 * g->cfg.nominal_width * sg * c;
            g->fyaw[o] = th;
            const double zl = -g->cfg.zmp_delta[1] * sg;
            g->zx[o] = g->fx[o] + c * g->cfg.zmp_delta[0] - s * zl;
            g->zy[o] = g->fy[o] + s * g->cfg.zmp_delta[0] + c * zl;
        }
        size_t o = (size_t)(S2 - 1) * B + i;
        g->ex[o] = g->zx[o]; g->ey[o] = g->zy[o];
        for (int j = S2 - 2; j >= 0; j--) {
            const size_t o1 = (size_t)(j + 1) * B + i; o = (size_t)j * B + i;
            g->ex[o] = g->zx[o1] + decay * (g->ex[o1] - g->zx[o1]);
            g->ey[o] = g->zy[o1] + decay * (g->ey[o1] - g->zy[o1]);
        }
    }
}

static int floordiv_i(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

static void dcm_at(const wco_traj* g, int s, int i, double* pos, double* vel)
{
    const int a = s + g->phase[i], T = g->cfg.step_ticks;
    const int k = floordiv_i(a, T);
    const double tau = (double)(a - k * T) * g->cfg.sampling_time;
    const double e = exp(g->omega * (tau - T * g->cfg.sampling_time));
    const double zx = TB(zx, k), zy = TB(zy, k), gx = TB(ex, k) - zx, gy = TB(ey, k) - zy;
    pos[0] = zx + e * gx; pos[1] = zy + e * gy;
    if (vel) { vel[0] = g->omega * (e * gx); vel[1] = g->omega * (e * gy); }
}

void wco_traj_dcm(const wco_traj* g, int first_sample, int n_samples, double* pos_out, double* vel_out)
{
    const size_t B = (size_t)g->B;
    for (int r = 0; r < n_samples; r++)
        for (int i = 0; i < g->B; i++) {
            double p[2], v[2];
            dcm_at(g, first_sample + r, i, p, v);
            if (pos_out) { pos_out[(size_t)(2 * r) * B + i] = p[0]; pos_out[(size_t)(2 * r + 1) * B + i] = p[1]; }
            if (vel_out) { vel_out[(size_t)(2 * r) * B + i] = v[0]; vel_out[(size_t)(2 * r + 1) * B + i] = v[1]; }
        }
}

/* Philox-4x32-10, Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3" (SC'11) */
void wco_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void wco_traj_measured_dcm(const wco_traj* g, int tick, double sigma, uint64_t seed, double* out)
{
    for (int i = 0; i < g->B; i++) {
        double p[2];
        dcm_at(g, tick, i, p, NULL);
        uint32_t c[4] = {(uint32_t)i, (uint32_t)tick, 0u, 0u};
        wco_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u1 = ((double)((((uint64_t)c[0] << 32) | c[1]) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
        const double u2 = (double)((((uint64_t)c[2] << 32) | c[3]) >> 11) * (1.0 / 9007199254740992.0);
        const double rad = sqrt(-2.0 * log(u1)), ang = 6.283185307179586476925 * u2;
        out[i] = p[0] + sigma * (rad * cos(ang));
        out[(size_t)g->B + i] = p[1] + sigma * (rad * sin(ang));
    }
}

static void put_H(double* H, size_t B, int i, double x, double y, double z, double yaw)
{
    const double c = cos(yaw), s = sin(yaw);
    const double v[12] = {c, -s, 0.0, x, s, c, 0.0, y, 0.0, 0.0, 1.0, z};
    for (int e = 0; e < 12; e++) H[(size_t)e * B + i] = v[e];
}

void wco_traj_feet(const wco_traj* g, int tick, const wc_traj_feet* o)
{
    const size_t B = (size_t)g->B;
    const int T = g->cfg.step_ticks, DS = g->cfg.ds_ticks;
    const double dt = g->cfg.sampling_time, h = g->cfg.step_height;
    for (int i = 0; i < g->B; i++) {
        const int a = tick + g->phase[i], k = floordiv_i(a, T), ii = a - k * T;
        const int in_ds = ii < DS;
        const double sx = TB(fx, k), sy = TB(fy, k), syaw = TB(fyaw, k);
        const double px = TB(fx, k - 1), py = TB(fy, k - 1), pyaw = TB(fyaw, k - 1);
        const double nx = TB(fx, k + 1), ny = TB(fy, k + 1), nyaw = TB(fyaw, k + 1);
        const double ss_len = (double)(T - DS);
        double u = (double)(ii - DS) / ss_len;
        if (u < 0.0) u = 0.0;
        if (u > 1.0) u = 1.0;
        const double blend = 3.0 * u * u - 2.0 * u * u * u;
        const double dblend = in_ds ? 0.0 : (6.0 * u - 6.0 * u * u) / (ss_len * dt);
        const double ox = px + blend * (nx - px), oy = py + blend * (ny - py), oyaw = pyaw + blend * (nyaw - pyaw);
        const double oz = in_ds ? 0.0 : 4.0 * h * u * (1.0 - u);
        const double tw[6] = {dblend * (nx - px), dblend * (ny - py), in_ds ? 0.0 : 4.0 * h * (1.0 - 2.0 * u) / (ss_len * dt), 0.0, 0.0,
                              dblend * (nyaw - pyaw)};
        const int stance_left = foot_sign(g, k, i) > 0.0;
        double* Hs = stance_left ? o->lf_H : o->rf_H; double* Ho = stance_left ? o->rf_H : o->lf_H;
        if (Hs) put_H(Hs, B, i, sx, sy, 0.0, syaw);
        if (Ho) put_H(Ho, B, i, ox, oy, oz, oyaw);
        double* ts = stance_left ? o->lf_twist : o->rf_twist; double* to = stance_left ? o->rf_twist : o->lf_twist;
        for (int e = 0; e < 6; e++) { if (ts) ts[(size_t)e * B + i] = 0.0; if (to) to[(size_t)e * B + i] = tw[e]; }
        if (o->l_contact) o->l_contact[i] = (stance_left || in_ds) ? 1 : 0;
        if (o->r_contact) o->r_contact[i] = (!stance_left || in_ds) ? 1 : 0;
        if (o->left_is_fixed) o->left_is_fixed[i] = stance_left ? 1 : 0;
        if (o->com_height) o->com_height[i] = g->cfg.com_height;
        if (o->com_height_vel) o->com_height_vel[i] = 0.0;
    }
}
