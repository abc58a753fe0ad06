"""on-device producer of the path's per-tick inputs (the signals of TrajectoryGenerator.cpp:501-598 for
periodic gaits with the shipped plannerParams.ini values).

CPU: the C twin (oracle/traj_ref.c) against the torch generator (walking-controllers_b200/synthetic.py, the workload every
other test and the bench use), against LIPM properties (xi' = omega (xi - z), continuity at step boundaries), and Philox
known-answer vectors.  GPU: the CUDA kernels against the C twin through the C ABI (tolerance 1e-12: libm vs CUDA sin/cos/exp
differ in the last ulp), then an MPC run fed by the device producer against one fed by the host generator."""
import math

import numpy as np
import pytest
import torch


def _gait(syn, B, T, N, seed=11):
    g = syn.GaitBatch(B, T, N, seed=seed)
    return g, dict(step_len=g.step_len.numpy(), phase=g.phase.numpy().astype(np.int32), dtheta=g.dtheta.numpy(),
                   theta0=g.theta0.numpy(), first_left=g.first_left.numpy().astype(np.uint8))


def test_philox_known_answers(oracle):
    # Random123 kat_vectors: philox4x32 10 rounds
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_traj_twin_matches_torch_generator(oracle, syn):
    B, T, N = 257, 200, 25
    g, par = _gait(syn, B, T, N)
    cfg = oracle.traj_default_config(T, N)
    assert cfg.n_steps + 2 == g.S + 2 and cfg.step_ticks == syn.STEP_TICKS and cfg.ds_ticks == syn.DS_TICKS
    tw = oracle.TrajBatch(cfg, B)
    tw.set_gait(**par)
    pos, vel = tw.dcm(0, T + N + 1)
    ref = g.trajectory().numpy()
    assert np.abs(pos - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    for t in (0, 1, 36, 37, 38, 89, 90, 91, 133, 199):
        f = g.feet(t); o = tw.feet(t)
        for k in ("lf_H", "rf_H", "lf_twist", "rf_twist"):
            assert np.abs(o[k] - f[k].numpy()).max() <= 1e-12 * max(1.0, f[k].abs().max().item()), (t, k)
        for k in ("l_contact", "r_contact", "left_is_fixed"):
            assert np.array_equal(o[k], f[k].numpy()), (t, k)
        assert np.all(o["com_height"] == cfg.com_height) and not o["com_height_vel"].any()


def test_traj_lipm_properties(oracle, syn):
    B, T, N = 64, 300, 25
    g, par = _gait(syn, B, T, N, seed=3)
    cfg = oracle.traj_default_config(T, N)
    tw = oracle.TrajBatch(cfg, B); tw.set_gait(**par)
    omega = math.sqrt(cfg.gravity / cfg.com_height)
    pos, vel = tw.dcm(0, T)
    pos = pos.reshape(T, 2, B); vel = vel.reshape(T, 2, B)
    # continuity: no jump larger than a tick of DCM motion at step boundaries
    assert np.abs(np.diff(pos, axis=0)).max() < 0.02
    # xi' = omega (xi - z) with z piecewise constant: within a step, (xi - vel / omega) is the constant ZMP of that step
    z = pos - vel / omega
    a = np.arange(T)[:, None] + par["phase"][None, :]
    same_step = (a[1:] // cfg.step_ticks) == (a[:-1] // cfg.step_ticks)
    dz = np.abs(np.diff(z, axis=0)).max(axis=1)
    assert dz[same_step].max() < 1e-12
    # the velocity is the derivative of the position (central differences inside a step)
    inner = same_step[1:] & same_step[:-1]
    fd = (pos[2:] - pos[:-2]) / (2 * cfg.sampling_time)
    err = np.abs(fd - vel[1:-1]).max(axis=1)
    assert err[inner].max() < 5e-4
    # measured DCM: zero-mean noise of the requested size, deterministic in (seed, tick), different across ticks
    m0 = tw.measured_dcm(7, 0.002, 0x5EED); m1 = tw.measured_dcm(7, 0.002, 0x5EED); m2 = tw.measured_dcm(8, 0.002, 0x5EED)
    assert np.array_equal(m0, m1) and not np.array_equal(m0 - pos[7], m2 - pos[8])
    big = oracle.TrajBatch(cfg, 20000); big.set_gait(step_len=np.full(20000, 0.1))
    nz = big.measured_dcm(3, 0.002, 1) - big.dcm(3, 1)[0]
    assert abs(nz.mean()) < 1e-4 and abs(nz.std() - 0.002) < 5e-5
    assert np.array_equal(tw.measured_dcm(7, 0.0, 5), pos[7])


@pytest.mark.gpu
def test_traj_gpu_parity(wc, oracle, syn, cuda_device):
    B, T, N = 3001, 260, 25
    g, par = _gait(syn, B, T, N, seed=5)
    cfg_o = oracle.traj_default_config(T, N); cfg_g = wc.traj_default_config(T, N)
    tw = oracle.TrajBatch(cfg_o, B); tw.set_gait(**par)
    tg = wc.TrajBatch(cfg_g, B)
    tg.set_gait(**{k: torch.from_numpy(v).to(cuda_device) for k, v in par.items()})
    po, vo = tw.dcm(-5, T + N + 6)
    pg, vg = tg.dcm(-5, T + N + 6)
    assert np.abs(pg.cpu().numpy() - po).max() <= 1e-12 * max(1.0, np.abs(po).max())
    assert np.abs(vg.cpu().numpy() - vo).max() <= 1e-12 * max(1.0, np.abs(vo).max())
    for t in (0, 36, 37, 89, 90, 180, 259):
        fo = tw.feet(t); fg = tg.feet(t)
        for k in ("lf_H", "rf_H", "lf_twist", "rf_twist", "com_height", "com_height_vel"):
            assert np.abs(fg[k].cpu().numpy() - fo[k]).max() <= 1e-12 * max(1.0, np.abs(fo[k]).max()), (t, k)
        for k in ("l_contact", "r_contact", "left_is_fixed"):
            assert np.array_equal(fg[k].cpu().numpy(), fo[k]), (t, k)
        mo = tw.measured_dcm(t, 0.002, 0x5EED); mg = tg.measured_dcm(t, 0.002, 0x5EED)
        assert np.abs(mg.cpu().numpy() - mo).max() <= 1e-12
    assert tg.launches == 2 + 14
    tg.close()


@pytest.mark.gpu
def test_mpc_fed_by_device_producer(wc, oracle, syn, cuda_device):
    """Device-resident loop: producer -> MPC, no host generator; the oracle MPC fed by the CPU twin must agree."""
    B, T, N = 512, 100, 25
    g, par = _gait(syn, B, T, N, seed=9)
    tw = oracle.TrajBatch(oracle.traj_default_config(T, N), B); tw.set_gait(**par)
    tg = wc.TrajBatch(wc.traj_default_config(T, N), B)
    tg.set_gait(**{k: torch.from_numpy(v).to(cuda_device) for k, v in par.items()})
    mo = oracle.MpcBatch(oracle.mpc_default_config(N), B); mg = wc.MpcBatch(wc.mpc_default_config(N), B)
    for t in range(T):
        fo = tw.feet(t); ro, _ = tw.dcm(t, N + 1); do = tw.measured_dcm(t, 0.002, 77)
        fg = tg.feet(t); rg, _ = tg.dcm(t, N + 1, want_vel=False); dg = tg.measured_dcm(t, 0.002, 77)
        out = mg.tick(fg["lf_H"], fg["rf_H"], fg["l_contact"], fg["r_contact"], dg, rg)
        torch.cuda.synchronize()
        # the producers agree to the last ulps of sin / cos / exp; the MPC comparison itself runs on identical inputs
        assert np.abs(rg.cpu().numpy() - ro).max() <= 1e-12 and np.abs(dg.cpu().numpy() - do).max() <= 1e-12
        assert np.abs(fg["lf_H"].cpu().numpy() - fo["lf_H"]).max() <= 1e-12 and np.array_equal(fg["l_contact"].cpu().numpy(), fo["l_contact"])
        zo, so, io_, flo = mo.tick(fg["lf_H"].cpu().numpy(), fg["rf_H"].cpu().numpy(), fg["l_contact"].cpu().numpy(),
                                   fg["r_contact"].cpu().numpy(), dg.cpu().numpy(), rg.cpu().numpy())
        assert np.array_equal(out["status"].cpu().numpy(), so) and np.array_equal(out["iters"].cpu().numpy(), io_), t
        zg = out["zmp"].cpu().numpy()
        assert (np.abs(zg - zo) / np.maximum(1.0, np.abs(zo))).max() <= 1e-6
    mg.close(); tg.close()


@pytest.mark.gpu
def test_device_resident_rollout_all_blocks(wc, oracle, syn, cuda_device):
    """The whole tick on the device with no host generator and no host round trip: input producer -> DCM-MPC (and the reactive
    controller beside it) -> LIPM / ZMP-CoM law -> FK + QP-IK about the integrated posture -> joint integration -> logger row;
    the CPU chain (producer twin + oracle) runs the same sequence independently."""
    B, N, T = 96, 25, 10
    g, par = _gait(syn, B, T, N, seed=17)
    dev = cuda_device
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    tw = oracle.TrajBatch(oracle.traj_default_config(T, N), B); tw.set_gait(**par)
    tg = wc.TrajBatch(wc.traj_default_config(T, N), B); tg.set_gait(**{k: d(v) for k, v in par.items()})
    mo = oracle.MpcBatch(oracle.mpc_default_config(N), B); mg = wc.MpcBatch(wc.mpc_default_config(N), B)
    go = oracle.GlueBatch(oracle.glue_default_config(), B); gg = wc.GlueBatch(wc.glue_default_config(), B)
    icfg_o, model_o, icfg_g = oracle.ik_default_config(), oracle.model_icub_like(), wc.ik_default_config()
    for j in range(23):
        icfg_o.min_vel[j], icfg_o.max_vel[j] = -50.0, 50.0
        icfg_g.min_vel[j], icfg_g.max_vel[j] = -50.0, 50.0
    ik = wc.IkBatch(icfg_g, wc.model_icub_like(), B)
    omega = math.sqrt(go.cfg.gravity / go.cfg.com_height)
    q0 = np.tile(np.array(icfg_o.joint_reg_rad[:23])[:, None], (1, B)) + 0.01 * np.random.default_rng(2).normal(size=(23, B))
    f0 = tw.feet(0)
    base0 = np.where(f0["left_is_fixed"][None, :] != 0, f0["lf_H"], f0["rf_H"])
    com0 = oracle.fk_batch(model_o, q0, base0, f0["left_is_fixed"], jacobians=False)["com"]
    go.reset(com0[:2], q0); gg.reset(d(com0[:2]), d(q0))
    q_o, q_g = q0.copy(), d(q0)
    eye = np.tile(np.eye(3).reshape(9, 1), (1, B)); eye_g = d(eye)
    for t in range(T):
        # ---- CPU chain
        fo = tw.feet(t); ro, vo = tw.dcm(t, N + 1); do = tw.measured_dcm(t, 0.002, 99)
        zo, so, io_, _ = mo.tick(fo["lf_H"], fo["rf_H"], fo["l_contact"], fo["r_contact"], do, ro)
        zr_o = oracle.reactive_dcm(omega, 1.5, ro[0:2], vo[0:2], do)
        base_o = np.where(fo["left_is_fixed"][None, :] != 0, fo["lf_H"], fo["rf_H"])
        ko = oracle.fk_batch(model_o, q_o, base_o, fo["left_is_fixed"], jacobians=False)
        co = go.com_reference(ro[0:2], zo, zo + 0.001, ko["com"], fo["com_height"], fo["com_height_vel"])
        ino = dict(q_meas=q_o, q_lin=q_o, base_H=base_o, left_is_fixed=fo["left_is_fixed"], lf_H_des=fo["lf_H"], rf_H_des=fo["rf_H"],
                   lf_twist=fo["lf_twist"], rf_twist=fo["rf_twist"], com_pos_des=co["com_pos_des"], com_vel_des=co["com_vel_des"],
                   com_meas=ko["com"], neck_R_des=eye)
        oo = oracle.ik_batch_tick(icfg_o, model_o, ino)
        q_o = go.integrate_joints(oo["dq"])
        row_o = oracle.log_pack(B, dcm=do, dcm_des=ro[0:2], dcm_vel_des=vo[0:2], zmp=zo + 0.001, zmp_des=zo, com=ko["com"],
                                com_des_xy=co["com_pos_des"], com_vel_des_xy=co["com_vel_des"], lf_H=ko["lf_H"], rf_H=ko["rf_H"],
                                lf_H_des=fo["lf_H"], rf_H_des=fo["rf_H"], lf_err=oo["lf_err"], rf_err=oo["rf_err"])
        # ---- device chain: every tensor stays on the GPU
        fg = tg.feet(t); rg, vg = tg.dcm(t, N + 1); dg = tg.measured_dcm(t, 0.002, 99)
        mgo = mg.tick(fg["lf_H"], fg["rf_H"], fg["l_contact"], fg["r_contact"], dg, rg)
        zr_g = gg.reactive_dcm(1.5, rg[0:2], vg[0:2], dg)
        base_g = torch.where(fg["left_is_fixed"][None, :] != 0, fg["lf_H"], fg["rf_H"]).contiguous()
        kg = ik.fk(q_g, base_g, fg["left_is_fixed"], jacobians=False)
        cg = gg.com_reference(rg[0:2], mgo["zmp"], mgo["zmp"] + 0.001, kg["com"], fg["com_height"], fg["com_height_vel"])
        ing = dict(q_meas=q_g, q_lin=q_g, base_H=base_g, left_is_fixed=fg["left_is_fixed"], lf_H_des=fg["lf_H"], rf_H_des=fg["rf_H"],
                   lf_twist=fg["lf_twist"], rf_twist=fg["rf_twist"], com_pos_des=cg["com_pos_des"], com_vel_des=cg["com_vel_des"],
                   com_meas=kg["com"], neck_R_des=eye_g)
        og = ik.tick(ing)
        q_g = gg.integrate_joints(og["dq"])
        row_g = gg.log_pack(dcm=dg, dcm_des=rg[0:2], dcm_vel_des=vg[0:2], zmp=mgo["zmp"] + 0.001, zmp_des=mgo["zmp"], com=kg["com"],
                            com_des_xy=cg["com_pos_des"], com_vel_des_xy=cg["com_vel_des"], lf_H=kg["lf_H"], rf_H=kg["rf_H"],
                            lf_H_des=fg["lf_H"], rf_H_des=fg["rf_H"], lf_err=og["lf_err"], rf_err=og["rf_err"])
        torch.cuda.synchronize()
        assert np.array_equal(mgo["status"].cpu().numpy(), so) and np.array_equal(mgo["iters"].cpu().numpy(), io_), t
        assert np.array_equal(og["status"].cpu().numpy(), oo["status"]), t
        assert np.abs(mgo["zmp"].cpu().numpy() - zo).max() <= 1e-6 and np.abs(zr_g.cpu().numpy() - zr_o).max() <= 1e-11
        assert np.abs(og["dq"].cpu().numpy() - oo["dq"]).max() <= 1e-5, t
        assert np.abs(q_g.cpu().numpy() - q_o).max() <= 1e-7
        assert np.abs(row_g.cpu().numpy() - row_o).max() <= 1e-5
    mg.close(); ik.close(); gg.close(); tg.close()
