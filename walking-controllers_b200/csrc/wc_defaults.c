/*
 * wc_defaults.c — host-only default parameter sets (no CUDA).
 *
 * Values are the ones shipped with the reference for iCubGazeboV2_5
 * (app/robots/iCubGazeboV2_5/{controllerParams,qpInverseKinematics,dcmWalkingCoordinator,
 * robotControl}.ini).  The kinematic model is NOT shipped with the reference (the URDF is
 * located at run time through YARP's ResourceFinder, WalkingModule.cpp:107-113), so
 * wc_model_icub_like() is a synthetic iCub-sized 23-DoF tree with the reference's joint
 * order; any other model can be supplied through wc_model.
 */
#include "../../include/wc_capi.h"

#include <math.h>
#include <string.h>

void wc_osqp_default_settings(wc_osqp_settings* s)
{
    memset(s, 0, sizeof(*s));
    s->rho = 0.1;
    s->sigma = 1e-6;
    s->alpha = 1.6;
    s->eps_abs = 1e-3;
    s->eps_rel = 1e-3;
    s->eps_prim_inf = 1e-4;
    s->eps_dual_inf = 1e-4;
    s->adaptive_rho_tolerance = 5.0;
    s->infty = 1e30;
    s->max_iter = 4000;
    s->scaling = 10;
    s->check_termination = 25;
    s->adaptive_rho = 1;
    s->adaptive_rho_interval = 0;
    s->warm_start = 1;
    s->osqp_eigen_checks_status = 0; /* osqp-eigen of the reference's time: solve() does not look at the status */
    s->reserved1 = 0;
}

void wc_mpc_default_config(wc_mpc_config* cfg, int32_t horizon)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->horizon = horizon;
    cfg->sampling_time = 0.01;  /* dcmWalkingCoordinator.ini:21 */
    cfg->com_height = 0.53;     /* dcmWalkingCoordinator.ini:19 */
    cfg->gravity = 9.81;
    cfg->Q[0] = 7500.0; cfg->Q[3] = 7500.0;         /* controllerParams.ini:3 */
    cfg->R[0] = 9000000.0; cfg->R[3] = 9000000.0;   /* controllerParams.ini:4 */
    cfg->foot_x[0] = -0.02; cfg->foot_x[1] = 0.05;  /* controllerParams.ini:7 */
    cfg->foot_y[0] = -0.025; cfg->foot_y[1] = 0.025;
    cfg->initial_zmp[0] = 0.0; cfg->initial_zmp[1] = 0.0;
    cfg->hull_tolerance = 0.05; /* controllerParams.ini:10 */
    wc_osqp_default_settings(&cfg->osqp);
}

void wc_ik_default_config(wc_ik_config* cfg)
{
    static const double reg_deg[23] = {15, 0, 0, -7, 22, 11, 30, -7, 22, 11, 30,
                                       5.082, 0.406, -0.131, -45.249, -26.454, -0.351,
                                       5.082, 0.406, -0.131, -45.249, -26.454, -0.351};
    static const double reg_w[23] = {1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
    int i;
    memset(cfg, 0, sizeof(*cfg));
    cfg->n_dofs = 23;
    cfg->use_com_as_constraint = 1;
    cfg->com_weight[0] = cfg->com_weight[4] = cfg->com_weight[8] = 100.0; /* commented-out default */
    cfg->neck_weight[0] = cfg->neck_weight[4] = cfg->neck_weight[8] = 5.0;
    /* additional_rotation ((0 0 1),(1 0 0),(0 1 0)) */
    cfg->additional_rotation[2] = 1.0;
    cfg->additional_rotation[3] = 1.0;
    cfg->additional_rotation[7] = 1.0;
    for (i = 0; i < 23; i++) {
        cfg->joint_reg_rad[i] = reg_deg[i] * (3.14159265358979323846 / 180.0);
        cfg->joint_reg_weight[i] = reg_w[i];
        cfg->joint_reg_gain[i] = 5.0;
        cfg->min_vel[i] = -1.0;
        cfg->max_vel[i] = 1.0;
    }
    cfg->k_pos_com = 1.0;
    cfg->k_pos_foot = 4.0;
    cfg->k_att_foot = 2.0;
    cfg->k_neck = 1.0;
    cfg->nwsr = 100;
}

/* ---------------------------------------------------------------- model table helpers */
static void rot_rpy(double r, double p, double y, double* R)
{
    const double cr = cos(r), sr = sin(r), cp = cos(p), sp = sin(p), cy = cos(y), sy = sin(y);
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

static int add_link(wc_model* m, int parent, int dof, double px, double py, double pz,
                    double rr, double rp, double ry, double ax, double ay, double az,
                    double mass, double cx, double cy, double cz)
{
    const int i = m->n_links++;
    const double nrm = sqrt(ax * ax + ay * ay + az * az);
    m->parent[i] = parent;
    m->dof[i] = dof;
    rot_rpy(rr, rp, ry, m->joint_R[i]);
    m->joint_p[i][0] = px; m->joint_p[i][1] = py; m->joint_p[i][2] = pz;
    m->axis[i][0] = nrm > 0 ? ax / nrm : 0; m->axis[i][1] = nrm > 0 ? ay / nrm : 0; m->axis[i][2] = nrm > 0 ? az / nrm : 0;
    m->mass[i] = mass;
    m->com[i][0] = cx; m->com[i][1] = cy; m->com[i][2] = cz;
    return i;
}

void wc_model_icub_like(wc_model* m)
{
    int s, k;
    memset(m, 0, sizeof(*m));
    /* 0: root_link (pelvis) */
    add_link(m, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4.72, 0.0, 0.0, -0.03);
    /* torso: pitch (y), roll (x), yaw (z) */
    const int t1 = add_link(m, 0, 0, 0.0, 0.0, 0.02, 0, 0, 0, 0, 1, 0, 1.1, 0.0, 0.0, 0.01);
    const int t2 = add_link(m, t1, 1, 0.0, 0.0, 0.032, 0, 0, 0, 1, 0, 0, 1.2, 0.0, 0.0, 0.02);
    const int chest = add_link(m, t2, 2, 0.0, 0.0, 0.05, 0, 0, 0, 0, 0, 1, 7.6, -0.01, 0.0, 0.09);
    /* head lump on locked neck joints; neck_2 frame lives here */
    const int head = add_link(m, chest, -1, -0.005, 0.0, 0.21, 0, 0.05, 0, 0, 0, 0, 2.6, 0.01, 0.0, 0.08);
    /* arms: shoulder pitch/roll/yaw + elbow, shoulder frames tilted by 15 deg */
    int dofa = 3;
    for (s = 0; s < 2; s++) {
        const double sg = s == 0 ? 1.0 : -1.0; /* left: +y */
        const int s1 = add_link(m, chest, dofa + 0, 0.0, sg * 0.11, 0.175, sg * 0.2618, 0, 0, 0, 1, 0, 0.9, 0.0, sg * 0.02, 0.0);
        const int s2 = add_link(m, s1, dofa + 1, 0.0, sg * 0.03, 0.0, 0, 0, 0, 1, 0, 0, 0.6, 0.0, sg * 0.01, -0.01);
        const int ua = add_link(m, s2, dofa + 2, 0.0, 0.0, -0.02, 0, 0, 0, 0, 0, 1, 1.3, 0.0, 0.0, -0.07);
        add_link(m, ua, dofa + 3, 0.015, 0.0, -0.152, 0, 0, 0, 0, 1, 0, 1.0, 0.0, 0.0, -0.09);
        dofa += 4;
    }
    /* legs: hip pitch/roll/yaw, knee, ankle pitch/roll */
    int foot[2] = {0, 0};
    int dofl = 11;
    for (s = 0; s < 2; s++) {
        const double sg = s == 0 ? 1.0 : -1.0;
        const int h1 = add_link(m, 0, dofl + 0, 0.0, sg * 0.068, -0.12, 0, 0, 0, 0, 1, 0, 0.75, 0.0, sg * 0.01, 0.0);
        const int h2 = add_link(m, h1, dofl + 1, 0.0, 0.0, 0.0, 0, 0, 0, 1, 0, 0, 0.53, 0.0, 0.0, -0.02);
        const int ul = add_link(m, h2, dofl + 2, 0.0, 0.0, -0.04, 0, 0, 0, 0, 0, 1, 2.2, 0.0, 0.0, -0.09);
        const int ll = add_link(m, ul, dofl + 3, 0.0, 0.0, -0.2, 0, 0, 0, 0, 1, 0, 1.3, 0.0, 0.0, -0.1);
        const int a1 = add_link(m, ll, dofl + 4, 0.0, 0.0, -0.21, 0, 0, 0, 0, 1, 0, 0.75, 0.0, 0.0, 0.0);
        foot[s] = add_link(m, a1, dofl + 5, 0.0, 0.0, -0.02, 0, 0, 0, 1, 0, 0, 0.65, 0.025, 0.0, -0.02);
        dofl += 6;
    }
    m->n_dofs = 23;
    m->l_sole_link = foot[0];
    m->r_sole_link = foot[1];
    m->neck_link = head;
    for (k = 0; k < 9; k++) { m->l_sole_R[k] = 0; m->r_sole_R[k] = 0; m->neck_R[k] = 0; }
    m->l_sole_R[0] = m->l_sole_R[4] = m->l_sole_R[8] = 1.0;
    m->r_sole_R[0] = m->r_sole_R[4] = m->r_sole_R[8] = 1.0;
    m->l_sole_p[0] = 0.0; m->l_sole_p[1] = 0.0; m->l_sole_p[2] = -0.045;
    m->r_sole_p[0] = 0.0; m->r_sole_p[1] = 0.0; m->r_sole_p[2] = -0.045;
    /* neck_2 frame of the iCub has its axes permuted with respect to the world-aligned chest:
     * world_R_neck * additional_rotation^-1 ~ identity at the nominal posture, with
     * additional_rotation = ((0 0 1),(1 0 0),(0 1 0)) (qpInverseKinematics.ini:7) */
    m->neck_R[2] = 1.0; m->neck_R[3] = 1.0; m->neck_R[7] = 1.0;
    m->neck_p[0] = 0.0; m->neck_p[1] = 0.0; m->neck_p[2] = 0.03;
}

/* iCubGazeboV2_5: dcmWalkingCoordinator.ini:20 (sampling_time), GENERAL com_height, zmpControllerParams.ini
 * (kZMP_walking, kCoM_walking: the gains used when gain scheduling is off) */
void wc_glue_default_config(wc_glue_config* c)
{
    if (!c) return;
    c->n_dofs = 23;
    c->use_gain_scheduling = 1; /* zmpControllerParams.ini:2 */
    c->sampling_time = 0.01;
    c->com_height = 0.53;
    c->gravity = 9.81;
    c->k_zmp = 3.0;
    c->k_com = 9.0;
    c->smoothing_time = 0.05;
    c->k_zmp_stance = 0.9;
    c->k_com_stance = 6.0;
}

/* iCubGazeboV2_5 plannerParams.ini: nominalDuration 0.9, switchOverSwingRatio 0.7 (double support = 0.7 / 1.7 of a
 * step = 37 ticks), nominalWidth 0.16, stepHeight 0.025, left/rightZMPDelta (0.03, -/+0.005); dcmWalkingCoordinator.ini
 * sampling_time 0.01, com_height 0.53.  Eight look-ahead steps truncate the backward DCM recursion at e^{-omega 7.2}. */
void wc_traj_default_config(wc_traj_config* c, int32_t n_ticks, int32_t horizon)
{
    if (!c) return;
    c->step_ticks = 90; c->ds_ticks = 37; c->reserved = 0;
    c->n_steps = (n_ticks + horizon + 90) / 90 + 8 + 3;
    c->sampling_time = 0.01;
    c->com_height = 0.53;
    c->gravity = 9.81;
    c->nominal_width = 0.16;
    c->step_height = 0.025;
    c->zmp_delta[0] = 0.03; c->zmp_delta[1] = 0.005;
}
