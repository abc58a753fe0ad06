"""Closed-loop glue: StableDCMModel + WalkingZMPController::evaluateControl + joint integrator.

CPU: the oracle restatement against closed forms (parity unpinned: the reference has no tests and
iCub::ctrl::Integrator is an external library; the Tustin rule is restated).  GPU: the CUDA kernels against the
oracle through the C ABI (wc_glue_*), and a device-resident closed-loop rollout MPC -> glue -> FK + QP-IK ->
integrator whose every tick is checked against the same chain on the CPU oracle."""
import math

import numpy as np
import pytest
import torch


def test_tustin_integrator_closed_form(oracle):
    cfg = oracle.glue_default_config(); cfg.n_dofs = 3
    B = 4
    g = oracle.GlueBatch(cfg, B)
    q0 = np.arange(3 * B, dtype=float).reshape(3, B)
    g.reset(q=q0)
    # constant input v from rest: y_k = y_0 + Ts v (k - 1/2)  (x_old = 0 after reset)
    v = np.full((3, B), 0.7)
    for k in range(1, 6):
        q = g.integrate_joints(v)
        assert np.allclose(q, q0 + cfg.sampling_time * 0.7 * (k - 0.5), rtol=0, atol=1e-13)


def test_lipm_tracks_constant_dcm(oracle):
    """dc = -omega (c - xi): for a constant xi the CoM converges to xi with rate omega (Tustin-discretised pole)"""
    cfg = oracle.glue_default_config()
    B = 3
    g = oracle.GlueBatch(cfg, B)
    c0 = np.array([[0.0, 0.1, -0.2], [0.0, 0.0, 0.05]])
    g.reset(com_xy=c0)
    xi = np.array([[0.05, 0.0, 0.1], [0.02, -0.03, 0.0]])
    z = np.zeros((2, B)); cm = np.zeros((3, B))
    omega = math.sqrt(cfg.gravity / cfg.com_height); ts = cfg.sampling_time
    c = c0.copy(); v_old = np.zeros_like(c)
    for k in range(300):
        out = g.com_reference(xi, z, z, cm)
        v = -omega * (c - xi); c = c + (v + v_old) * ts / 2; v_old = v
        # ZMP-CoM law with zero ZMP error and com_meas = 0: vd = kCoM c + v; setPhase is never called, so the scheduled
        # gain stays where WalkingZMPController::initialize leaves it: the stance value (WalkingZMPController.cpp:82-83)
        kc = cfg.k_com_stance if cfg.use_gain_scheduling else cfg.k_com
        assert np.allclose(out["com_vel_des"][:2], kc * c + v, rtol=1e-13, atol=1e-15)
    assert np.abs(c - xi).max() < 1e-4


def test_gain_schedule_smoother(oracle):
    """WalkingZMPController::setPhase with the restated minJerkTrajGen: the gains start at the stance values, move to the
    walking values when the phase flips, reach 90 % of the step after the smoothing time (the filter's design point), settle
    with unit DC gain, and the two smoothers are the same filter (their normalised responses coincide)."""
    cfg = oracle.glue_default_config()
    assert cfg.use_gain_scheduling == 1 and cfg.smoothing_time == 0.05            # zmpControllerParams.ini:2,4
    assert (cfg.k_zmp, cfg.k_com, cfg.k_zmp_stance, cfg.k_com_stance) == (3.0, 9.0, 0.9, 6.0)
    B = 2
    g = oracle.GlueBatch(cfg, B)
    z = np.zeros((2, B)); c = np.ones((2, B)); cm = np.zeros((3, B))
    g.reset(com_xy=c)

    # drive: stance flag for instance 0 stays 1 (gains stay at stance), instance 1 walks from tick 0
    stance = np.array([1, 0], np.uint8)
    ks = []
    for k in range(60):
        g.set_phase(stance=stance)
        out = g.com_reference(c, z, z, cm)          # dcm_des = c_lipm = 1 -> v = 0, vd = kCoM * 1
        ks.append(out["com_vel_des"][0].copy())
    ks = np.array(ks)
    assert np.allclose(ks[:, 0], cfg.k_com_stance, rtol=1e-12)                      # unit DC gain at rest
    step = (ks[:, 1] - cfg.k_com_stance) / (cfg.k_com - cfg.k_com_stance)
    n_T = int(round(cfg.smoothing_time / cfg.sampling_time))
    assert 0.85 < step[n_T - 1] < 0.95, step[:8]                                    # ~90 % of the step at t = T
    assert np.all(np.diff(step[:n_T]) > 0) and abs(step[-1] - 1.0) < 1e-6          # monotone rise, settles at the target
    # the velocity-threshold rule of updateModule: |dcm_vel_des| < 0.001 <=> stance
    g2 = oracle.GlueBatch(cfg, B); g2.reset(com_xy=c)
    vel = np.array([[0.0005, 0.002], [0.0, 0.0]])
    for k in range(60):
        g2.set_phase(dcm_vel_des=vel)
        out2 = g2.com_reference(c, z, z, cm)
    assert np.allclose(out2["com_vel_des"][0], [cfg.k_com_stance, cfg.k_com], rtol=1e-6)


@pytest.mark.gpu
def test_glue_gpu_parity(wc, oracle, cuda_device):
    B, nJ, T = 1000, 23, 40
    rng = np.random.default_rng(5)
    go = oracle.GlueBatch(oracle.glue_default_config(), B)
    gg = wc.GlueBatch(wc.glue_default_config(), B)
    c0 = rng.normal(size=(2, B)) * 0.05; q0 = rng.normal(size=(nJ, B)) * 0.3
    go.reset(c0, q0); gg.reset(torch.from_numpy(c0).to(cuda_device), torch.from_numpy(q0).to(cuda_device))
    kz = 3.0 * rng.choice([0.3, 1.0], size=B); kc = 9.0 * rng.choice([0.66, 1.0], size=B)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda_device)
    for t in range(T):
        dcm = rng.normal(size=(2, B)) * 0.05; zd = rng.normal(size=(2, B)) * 0.05; zm = zd + rng.normal(size=(2, B)) * 0.01
        cm = rng.normal(size=(3, B)) * 0.05; hz = 0.53 + rng.normal(size=B) * 0.001; hv = rng.normal(size=B) * 0.01
        use_k = t % 2 == 0
        vel = rng.normal(size=(2, B)) * 0.0012   # about half of the instances below the 0.001 stance threshold
        go.set_phase(dcm_vel_des=vel); gg.set_phase(dcm_vel_des=d(vel))
        oo = go.com_reference(dcm, zd, zm, cm, hz, hv, kz if use_k else None, kc if use_k else None)
        og = gg.com_reference(d(dcm), d(zd), d(zm), d(cm), d(hz), d(hv), d(kz) if use_k else None, d(kc) if use_k else None)
        dq = rng.normal(size=(nJ, B))
        qo = go.integrate_joints(dq); qg = gg.integrate_joints(d(dq))
        torch.cuda.synchronize()
        # same operation order; the only difference is FMA contraction on the device
        assert np.abs(og["com_pos_des"].cpu().numpy() - oo["com_pos_des"]).max() <= 1e-13
        assert np.abs(og["com_vel_des"].cpu().numpy() - oo["com_vel_des"]).max() <= 1e-13
        assert np.abs(qg.cpu().numpy() - qo).max() <= 1e-13
    assert gg.launches == 1 + 3 * T
    gg.close()


@pytest.mark.gpu
def test_closed_loop_rollout_on_device(wc, syn, oracle, cuda_device):
    """MPC -> LIPM/ZMP-CoM glue -> FK + QP-IK -> joint integration, T ticks, no host round trip on the GPU side;
    the CPU oracle runs the same chain.  The IK linearises about the integrated desired posture (quirk 9)."""
    B, N, T = 128, 25, 12
    g = syn.GaitBatch(B, T, N, seed=21)
    traj = g.trajectory()
    mo = oracle.MpcBatch(oracle.mpc_default_config(N), B); mg = wc.MpcBatch(wc.mpc_default_config(N), B)
    go = oracle.GlueBatch(oracle.glue_default_config(), B); gg = wc.GlueBatch(wc.glue_default_config(), B)
    icfg_o, model_o = oracle.ik_default_config(), oracle.model_icub_like()
    for j in range(23):
        icfg_o.min_vel[j], icfg_o.max_vel[j] = -20.0, 20.0
    icfg_g = wc.ik_default_config()
    for j in range(23):
        icfg_g.min_vel[j], icfg_g.max_vel[j] = -20.0, 20.0
    ik = wc.IkBatch(icfg_g, wc.model_icub_like(), B)
    qreg = torch.tensor(list(icfg_o.joint_reg_rad[:23]), dtype=torch.float64)

    def fkb(ql, bH, lif):
        o = oracle.fk_batch(model_o, ql.numpy(), bH.numpy(), lif.numpy(), jacobians=False)
        return {k: torch.from_numpy(v) for k, v in o.items()}

    inp = syn.ik_inputs(B, 9, qreg, fkb, perturb=0.2)
    q0 = inp["q_lin"].numpy().copy()
    com0 = fkb(inp["q_lin"], inp["base_H"], inp["left_is_fixed"])["com"].numpy()
    go.reset(com0[:2], q0)
    dev = cuda_device
    gg.reset(torch.from_numpy(np.ascontiguousarray(com0[:2])).to(dev), torch.from_numpy(q0).to(dev))
    inp_d = {k: v.to(dev) for k, v in inp.items()}
    traj_d = traj.to(dev)
    q_o = q0.copy(); q_g = inp_d["q_lin"].clone()
    for t in range(T):
        f = g.feet(t); dcm = g.measured_dcm(t); ref = traj[2 * t: 2 * t + 2 * (N + 1)].contiguous()
        # ---- oracle chain
        zo, so, _, _ = mo.tick(f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(), dcm.numpy(), ref.numpy())
        ko = oracle.fk_batch(model_o, q_o, inp["base_H"].numpy(), inp["left_is_fixed"].numpy(), jacobians=False)
        co = go.com_reference(ref[0:2].numpy(), zo, zo + 0.001, ko["com"], ko["com"][2], None)
        ino = {k: v.numpy() for k, v in inp.items()}
        ino.update(q_lin=q_o, q_meas=q_o, com_pos_des=co["com_pos_des"], com_vel_des=co["com_vel_des"], com_meas=ko["com"])
        oo = oracle.ik_batch_tick(icfg_o, model_o, ino)
        q_o = go.integrate_joints(oo["dq"])
        # ---- device chain (all tensors stay on the GPU)
        mgo = mg.tick(f["lf_H"].to(dev), f["rf_H"].to(dev), f["l_contact"].to(dev), f["r_contact"].to(dev), dcm.to(dev),
                      traj_d[2 * t: 2 * t + 2 * (N + 1)])
        kg = ik.fk(q_g, inp_d["base_H"], inp_d["left_is_fixed"], jacobians=False)
        cg = gg.com_reference(traj_d[2 * t: 2 * t + 2], mgo["zmp"], mgo["zmp"] + 0.001, kg["com"], kg["com"][2].contiguous(), None)
        ing = dict(inp_d)
        ing.update(q_lin=q_g, q_meas=q_g, com_pos_des=cg["com_pos_des"], com_vel_des=cg["com_vel_des"], com_meas=kg["com"])
        og = ik.tick(ing)
        q_g = gg.integrate_joints(og["dq"])
        torch.cuda.synchronize()
        assert np.array_equal(mgo["status"].cpu().numpy(), so)
        assert np.array_equal(og["status"].cpu().numpy(), oo["status"])
        assert np.abs(mgo["zmp"].cpu().numpy() - zo).max() <= 1e-6
        assert np.abs(og["dq"].cpu().numpy() - oo["dq"]).max() <= 1e-7, f"tick {t}"
        assert np.abs(q_g.cpu().numpy() - q_o).max() <= 1e-8
    mg.close(); ik.close(); gg.close()
