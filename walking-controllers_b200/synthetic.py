"""Synthetic iCub walking inputs for the batched hot path (tests, smoke, bench).

The reference gets these signals from TrajectoryGenerator/UnicyclePlanner
(modules/Walking_module/src/TrajectoryGenerator.cpp:501-613), which is out of scope
(it produces the hot path's *inputs*).  This module is the synthetic stand-in
described in SURVEY.md §8(d): an analytic LIPM divergent-component-of-motion reference for
periodic walking with the shipped planner parameters
(app/robots/iCubGazeboV2_5/plannerParams.ini: nominalWidth 0.16, nominalDuration 0.9 s,
switchOverSwingRatio 0.7, stepHeight 0.025, ZMP deltas (0.03, -/+0.005)), per-instance phase
offset, step length and heading rate.

Everything is torch (CPU or CUDA, float64) and field-major: element e of instance i of a
per-instance vector is at [e, i], matching the SoA layout of include/wc_capi.h.
"""
import math

import torch

DT = 0.01
STEP_TICKS = 90          # nominalDuration 0.9 s
DS_TICKS = 37            # 0.9 * 0.7 / 1.7 s of double support (switchOverSwingRatio 0.7)
WIDTH = 0.16             # nominalWidth
STEP_HEIGHT = 0.025
ZMP_DX, ZMP_DY = 0.03, 0.005
LOOKAHEAD_STEPS = 8      # e^{-omega * 0.9 * 8} ~ 1e-14: truncation of the backward DCM recursion


class GaitBatch:
    """Per-instance periodic gait; all queries are vectorised over the B instances."""

    def __init__(self, B, n_ticks, horizon, seed=0x5EED, device="cpu", com_height=0.53, gravity=9.81,
                 noise_dcm=0.002, random_phase=True, turning=True):
        g = torch.Generator(device="cpu").manual_seed(seed)
        f64 = torch.float64
        self.B, self.T, self.N, self.device = B, n_ticks, horizon, device
        self.omega = math.sqrt(gravity / com_height)
        self.noise_dcm = noise_dcm
        self.seed = seed
        if random_phase:
            self.phase = torch.randint(0, STEP_TICKS, (B,), generator=g).to(device)
        else:
            self.phase = torch.zeros(B, dtype=torch.int64, device=device)
        self.step_len = (0.05 + 0.15 * torch.rand(B, generator=g, dtype=f64)).to(device)
        dth = math.radians(10.0) if turning else 0.0
        self.dtheta = ((2 * torch.rand(B, generator=g, dtype=f64) - 1) * dth).to(device)
        self.theta0 = ((2 * torch.rand(B, generator=g, dtype=f64) - 1) * math.pi).to(device)
        self.first_left = (torch.rand(B, generator=g) < 0.5).to(device)
        # footsteps k = -2 .. S-1  (index k+2)
        S = (n_ticks + horizon + STEP_TICKS) // STEP_TICKS + LOOKAHEAD_STEPS + 3
        self.S = S
        k = torch.arange(-2, S, device=device, dtype=f64)[:, None]            # [S+2, 1]
        theta = self.theta0[None, :] + k * self.dtheta[None, :]                # yaw of footstep k
        # midpoint path: c_k = c_{k-1} + L (cos th_k, sin th_k)
        dx = self.step_len[None, :] * torch.cos(theta)
        dy = self.step_len[None, :] * torch.sin(theta)
        cx = torch.cumsum(dx, 0); cy = torch.cumsum(dy, 0)
        sign = torch.where((torch.arange(-2, S, device=device) % 2 == 0)[:, None] == self.first_left[None, :],
                           torch.ones((), dtype=f64, device=device), -torch.ones((), dtype=f64, device=device))
        self.foot_sign = sign                                                   # +1 = left foot
        self.foot_x = cx - 0.5 * WIDTH * sign * torch.sin(theta)
        self.foot_y = cy + 0.5 * WIDTH * sign * torch.cos(theta)
        self.foot_yaw = theta
        # ZMP of step k: foot position + R(yaw) (0.03, -/+0.005)
        zx_l, zy_l = ZMP_DX, -ZMP_DY * sign
        self.zmp_x = self.foot_x + torch.cos(theta) * zx_l - torch.sin(theta) * zy_l
        self.zmp_y = self.foot_y + torch.sin(theta) * zx_l + torch.cos(theta) * zy_l
        # backward DCM recursion: xi_end(k) = xi_start(k+1), xi_start(k) = z_k + e^{-w T}(xi_end(k) - z_k)
        decay = math.exp(-self.omega * STEP_TICKS * DT)
        ex = torch.empty_like(self.zmp_x); ey = torch.empty_like(self.zmp_y)
        ex[-1] = self.zmp_x[-1]; ey[-1] = self.zmp_y[-1]
        for j in range(S + 2 - 2, -1, -1):
            sx = self.zmp_x[j + 1] + decay * (ex[j + 1] - self.zmp_x[j + 1])
            sy = self.zmp_y[j + 1] + decay * (ey[j + 1] - self.zmp_y[j + 1])
            ex[j] = sx; ey[j] = sy
        self.xi_end_x, self.xi_end_y = ex, ey

    # -- helpers -----------------------------------------------------------------------------
    def _gather(self, table, k):
        """table [S+2, B], k [.., B] step indices (>= -2) -> values [.., B]"""
        idx = (k + 2).clamp(0, self.S + 1)
        return torch.gather(table, 0, idx.reshape(-1, self.B)).reshape(k.shape)

    def dcm_reference(self, s):
        """DCM reference at absolute sample indices s [n, B] (long) -> (x, y) each [n, B]."""
        a = s + self.phase[None, :]
        k = torch.div(a, STEP_TICKS, rounding_mode="floor")
        tau = (a - k * STEP_TICKS).to(torch.float64) * DT
        zx = self._gather(self.zmp_x, k); zy = self._gather(self.zmp_y, k)
        ex = self._gather(self.xi_end_x, k); ey = self._gather(self.xi_end_y, k)
        e = torch.exp(self.omega * (tau - STEP_TICKS * DT))
        return zx + e * (ex - zx), zy + e * (ey - zy)

    def trajectory(self):
        """Whole DCM reference [(T+N+1)*2, B]: sample s axis a at row 2*s+a.  The window of tick t
        (the reference's deque, advanced one sample per tick, WalkingModule.cpp:35-96) is rows
        [2t, 2t+2(N+1))."""
        n = self.T + self.N + 1
        s = torch.arange(n, device=self.device)[:, None].expand(n, self.B)
        x, y = self.dcm_reference(s)
        return torch.stack([x, y], 1).reshape(2 * n, self.B).contiguous()

    def measured_dcm(self, t):
        """Reference at tick t plus N(0, noise^2), deterministic in (seed, t). [2, B]"""
        s = torch.full((1, self.B), t, device=self.device, dtype=torch.int64)
        x, y = self.dcm_reference(s)
        g = torch.Generator(device="cpu").manual_seed(self.seed * 1000003 + t)
        nz = (torch.randn(2, self.B, generator=g, dtype=torch.float64) * self.noise_dcm).to(self.device)
        return torch.cat([x, y], 0) + nz

    @staticmethod
    def _H(x, y, z, yaw):
        """row-major 3x4 [R|p] of a planar pose, field-major [12, B]"""
        c, s = torch.cos(yaw), torch.sin(yaw)
        o, l = torch.zeros_like(x), torch.ones_like(x)
        return torch.stack([c, -s, o, x, s, c, o, y, o, o, l, z], 0)

    def feet(self, t):
        """Feet state at tick t: dict with lf_H, rf_H [12,B], l_contact, r_contact [B] uint8,
        lf_twist, rf_twist [6,B] (mixed: linear, angular), left_is_fixed [B] uint8."""
        f64 = torch.float64
        a = t + self.phase
        k = torch.div(a, STEP_TICKS, rounding_mode="floor")
        i = a - k * STEP_TICKS
        in_ds = i < DS_TICKS
        kk = k[None, :]
        # stance foot: footstep k.  other foot: k-1 (DS) or swinging k-1 -> k+1 (SS)
        sx, sy, syaw = (self._gather(tb, kk)[0] for tb in (self.foot_x, self.foot_y, self.foot_yaw))
        px, py, pyaw = (self._gather(tb, kk - 1)[0] for tb in (self.foot_x, self.foot_y, self.foot_yaw))
        nx, ny, nyaw = (self._gather(tb, kk + 1)[0] for tb in (self.foot_x, self.foot_y, self.foot_yaw))
        ss_len = float(STEP_TICKS - DS_TICKS)
        u = ((i - DS_TICKS).to(f64) / ss_len).clamp(0.0, 1.0)
        blend = 3 * u ** 2 - 2 * u ** 3
        dblend = torch.where(in_ds, torch.zeros_like(u), (6 * u - 6 * u ** 2) / (ss_len * DT))
        ox = px + blend * (nx - px); oy = py + blend * (ny - py); oyaw = pyaw + blend * (nyaw - pyaw)
        oz = torch.where(in_ds, torch.zeros_like(u), 4 * STEP_HEIGHT * u * (1 - u))
        ovx = dblend * (nx - px); ovy = dblend * (ny - py); owz = dblend * (nyaw - pyaw)
        ovz = torch.where(in_ds, torch.zeros_like(u), 4 * STEP_HEIGHT * (1 - 2 * u) / (ss_len * DT))
        stance_H = self._H(sx, sy, torch.zeros_like(sx), syaw)
        other_H = self._H(ox, oy, oz, oyaw)
        zero = torch.zeros_like(sx)
        stance_tw = torch.stack([zero] * 6, 0)
        other_tw = torch.stack([ovx, ovy, ovz, zero, zero, owz], 0)
        stance_left = self._gather(self.foot_sign, kk)[0] > 0
        lf_H = torch.where(stance_left[None, :], stance_H, other_H)
        rf_H = torch.where(stance_left[None, :], other_H, stance_H)
        lf_tw = torch.where(stance_left[None, :], stance_tw, other_tw)
        rf_tw = torch.where(stance_left[None, :], other_tw, stance_tw)
        l_contact = (stance_left | in_ds).to(torch.uint8)
        r_contact = (~stance_left | in_ds).to(torch.uint8)
        mean_yaw = torch.atan2(torch.sin(syaw) + torch.sin(oyaw), torch.cos(syaw) + torch.cos(oyaw))
        return dict(lf_H=lf_H.contiguous(), rf_H=rf_H.contiguous(), l_contact=l_contact, r_contact=r_contact,
                    lf_twist=lf_tw.contiguous(), rf_twist=rf_tw.contiguous(),
                    left_is_fixed=stance_left.to(torch.uint8), mean_yaw=mean_yaw,
                    stance_H=stance_H.contiguous())


def rotz(yaw):
    """[9, B] row-major rotation about z"""
    c, s = torch.cos(yaw), torch.sin(yaw)
    o, l = torch.zeros_like(yaw), torch.ones_like(yaw)
    return torch.stack([c, -s, o, s, c, o, o, o, l], 0)


def small_rotation(rx, ry, rz):
    """[9, B] row-major Rz(rz) Ry(ry) Rx(rx)"""
    cx, sx, cy, sy, cz, sz = torch.cos(rx), torch.sin(rx), torch.cos(ry), torch.sin(ry), torch.cos(rz), torch.sin(rz)
    return torch.stack([cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx,
                        sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx,
                        -sy, cy * sx, cy * cx], 0)


def matmul3(a, b):
    """[9,B] x [9,B] row-major 3x3 products"""
    rows = []
    for r in range(3):
        for c in range(3):
            rows.append(a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c])
    return torch.stack(rows, 0)


def ik_inputs(B, seed, q_reg, fk, device="cpu", perturb=1.0):
    """Synthetic QP-IK inputs around a random posture.

    fk(q_lin [nJ,B], base_H [12,B], left_is_fixed [B]) -> dict(lf_H, rf_H [12,B], neck_R [9,B], com [3,B])
    is the forward kinematics used to make desired quantities consistent with the posture (the GPU
    kernel in bench/smoke, the oracle in CPU tests).  Desired feet = actual feet composed with a
    small perturbation, so that task errors have the size seen in closed loop."""
    gdev = "cuda" if str(device).startswith("cuda") else "cpu"   # draw on the target device (no PCIe for the bench)
    g = torch.Generator(device=device if gdev == "cuda" else "cpu").manual_seed(seed)
    f64 = torch.float64

    def U(*shape, lo=-1.0, hi=1.0):
        return lo + (hi - lo) * torch.rand(*shape, generator=g, dtype=f64, device=device)

    nJ = q_reg.numel()
    q_reg = q_reg.to(device=device, dtype=f64)
    q_lin = q_reg[:, None] + 0.1 * U(nJ, B)
    q_meas = q_lin + 0.005 * U(nJ, B)
    yaw = math.pi * U(B)
    base_H = GaitBatch._H(2.0 * U(B), 2.0 * U(B), torch.zeros(B, dtype=f64, device=device), yaw)
    left_is_fixed = (torch.rand(B, generator=g, device=device) < 0.5).to(torch.uint8)
    kin = fk(q_lin.contiguous(), base_H.contiguous(), left_is_fixed)

    def perturb_H(H):
        dR = small_rotation(0.02 * perturb * U(B), 0.02 * perturb * U(B), 0.02 * perturb * U(B))
        R = torch.stack([H[0], H[1], H[2], H[4], H[5], H[6], H[8], H[9], H[10]], 0)
        Rn = matmul3(dR, R)
        p = torch.stack([H[3], H[7], H[11]], 0) + 0.005 * perturb * U(3, B)
        return torch.stack([Rn[0], Rn[1], Rn[2], p[0], Rn[3], Rn[4], Rn[5], p[1], Rn[6], Rn[7], Rn[8], p[2]], 0)

    lf_des = perturb_H(kin["lf_H"]); rf_des = perturb_H(kin["rf_H"])
    lf_tw = 0.05 * perturb * U(6, B); rf_tw = 0.05 * perturb * U(6, B)
    com_meas = kin["com"] + 0.001 * U(3, B)
    com_des = kin["com"] + 0.005 * perturb * U(3, B)
    com_vel = 0.05 * perturb * U(3, B)
    neck_des = rotz(yaw + 0.05 * perturb * U(B))
    c = lambda x: x.contiguous()
    return dict(q_meas=c(q_meas), q_lin=c(q_lin), base_H=c(base_H), left_is_fixed=left_is_fixed,
                lf_H_des=c(lf_des), rf_H_des=c(rf_des), lf_twist=c(lf_tw), rf_twist=c(rf_tw),
                com_pos_des=c(com_des), com_vel_des=c(com_vel), com_meas=c(com_meas), neck_R_des=c(neck_des))


# The closed-loop workload (wc_loop_*, tests/test_closed_loop.py, bench.py `closed_loop`).  The synthetic 23-DoF model
# (wc_model_icub_like) stands 0.34 m tall at its regularisation posture, not the iCub's 0.53 m: the LIPM height of the
# closed loop is the model's own.  The shipped +-1 rad/s velocity limit is a placeholder (the reference reads the real
# limits from the robot, RobotHelper.cpp:288-298) under which this model cannot follow the gait at all; +-2 rad/s keeps
# every instance feasible while ~9 % of the instance-ticks hit a bound (the share of the open-loop benchmark workload).
CL_COM_HEIGHT = 0.34
CL_VMAX = 2.0
CL_PREP_ITERS = 400


def prepare_posture(fk, ik_tick, q_reg, feet, com_target, iters=CL_PREP_ITERS, dt=DT):
    """The analogue of WalkingModule::prepareRobot (WalkingModule.cpp:452-540: an initial joint configuration that puts the
    CoM where the trajectory starts, above the desired feet) for the synthetic model: iterate the QP-IK itself, with joint
    velocity limits the caller has opened, from the regularisation posture -- q <- q + dt dq with the feet held at `feet`
    (the producer's signals of tick 0) and the CoM drawn to com_target [3, B] -- until it has converged.
    fk(q, base_H, left_is_fixed) -> dict(com=...), ik_tick(inp) -> dict(dq=...): the product's IkBatch or the oracle's
    restatement behind torch adapters.  Returns q0 [nJ, B]."""
    B = com_target.shape[1]
    dev = com_target.device
    q = q_reg.to(dev)[:, None].repeat(1, B).contiguous()
    lif = feet["left_is_fixed"]
    base = torch.where(lif[None, :] != 0, feet["lf_H"], feet["rf_H"]).contiguous()
    yl = torch.atan2(feet["lf_H"][4], feet["lf_H"][0]); yr = torch.atan2(feet["rf_H"][4], feet["rf_H"][0])
    my = torch.atan2(torch.sin(yl) + torch.sin(yr), torch.cos(yl) + torch.cos(yr))
    neck = rotz(my).contiguous()
    zero6 = torch.zeros(6, B, dtype=torch.float64, device=dev)
    zero3 = torch.zeros(3, B, dtype=torch.float64, device=dev)
    for _ in range(iters):
        com = fk(q, base, lif)["com"]
        inp = dict(q_meas=q, q_lin=q, base_H=base, left_is_fixed=lif, lf_H_des=feet["lf_H"], rf_H_des=feet["rf_H"],
                   lf_twist=zero6, rf_twist=zero6, com_pos_des=com_target.contiguous(), com_vel_des=zero3, com_meas=com,
                   neck_R_des=neck)
        q = (q + dt * ik_tick(inp)["dq"]).contiguous()
    return q
