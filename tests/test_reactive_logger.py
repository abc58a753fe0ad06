"""WalkingDCMReactiveController (WalkingDCMReactiveController.cpp:63-81), WalkingFK::evaluateDCM
(WalkingForwardKinematics.cpp:312-340), WalkingModule::evaluateZMP (WalkingModule.cpp:824-876) and the logger row /
file format (WalkingModule.cpp:793-803, 1231-1250; WalkingLoggerModule.cpp:54-90, 142-171).

CPU: the oracle against closed forms (parity unpinned: the reference holds no tests).  GPU: the kernels against the oracle
through the C ABI; elementwise fp64, same operation order, tolerance 1e-13 (FMA contraction on the device only)."""
import math

import numpy as np
import pytest
import torch


def _rot(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]]); Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _poses(rng, B):
    """[12, B] row-major 3x4 transforms + the rpy / position they were built from"""
    rpy = rng.uniform(-1.2, 1.2, size=(3, B)); pos = rng.normal(size=(3, B)) * 0.3
    H = np.zeros((12, B))
    for i in range(B):
        R = _rot(*rpy[:, i])
        for r in range(3):
            H[4 * r: 4 * r + 3, i] = R[r]; H[4 * r + 3, i] = pos[r, i]
    return H, rpy, pos


def test_reactive_dcm_closed_form(oracle):
    rng = np.random.default_rng(0)
    B = 7
    xd, vd, x = rng.normal(size=(2, B)), rng.normal(size=(2, B)), rng.normal(size=(2, B))
    omega = math.sqrt(9.81 / 0.53)
    u = oracle.reactive_dcm(omega, 1.5, xd, vd, x)
    assert np.allclose(u, xd - vd / omega - 1.5 * (xd - x), rtol=1e-15, atol=1e-15)
    # on the reference trajectory (x = xd) the output is the ZMP of the LIPM: xd - dxd / omega
    assert np.allclose(oracle.reactive_dcm(omega, 1.5, xd, vd, xd), xd - vd / omega, rtol=1e-15, atol=1e-15)
    k = rng.uniform(0.5, 2.0, size=B)
    assert np.allclose(oracle.reactive_dcm(omega, k, xd, vd, x), xd - vd / omega - k * (xd - x), rtol=1e-15, atol=1e-15)


def test_evaluate_zmp_properties(oracle):
    rng = np.random.default_rng(1)
    B = 64
    lf, _, _ = _poses(rng, B); rf, _, _ = _poses(rng, B)
    lw = rng.normal(size=(6, B)); rw = rng.normal(size=(6, B))
    lw[2] = rng.uniform(50, 200, size=B); rw[2] = rng.uniform(50, 200, size=B)
    rw[2, :8] = 0.0005      # right foot unloaded: ZMP = left local ZMP in the world, weight fzl / (fzl + fzr)
    lw[2, 8:12] = 0.01; rw[2, 8:12] = 0.02   # total < 0.1: the reference returns false
    z, ok = oracle.evaluate_zmp(lw, rw, lf, rf)
    assert np.array_equal(ok[8:12], np.zeros(4, dtype=np.int32)) and ok[:8].all() and ok[12:].all()
    for i in (0, 3, 20, 40):
        pl = np.array([-lw[4, i] / lw[2, i], lw[3, i] / lw[2, i], 0.0, 1.0])
        pr = np.array([-rw[4, i] / rw[2, i], rw[3, i] / rw[2, i], 0.0, 1.0])
        wl = lf[:, i].reshape(3, 4) @ pl; wr = rf[:, i].reshape(3, 4) @ pr
        tot = lw[2, i] + rw[2, i]
        want = lw[2, i] / tot * wl[:2] + (0.0 if i < 8 else rw[2, i] / tot) * wr[:2]
        assert np.allclose(z[:, i], want, rtol=1e-13, atol=1e-14)


def test_log_row_and_dataset_format(oracle, wc, tmp_path):
    rng = np.random.default_rng(2)
    B = 5
    lf, rpy_l, pos_l = _poses(rng, B); rf, rpy_r, pos_r = _poses(rng, B)
    f = dict(dcm=rng.normal(size=(2, B)), dcm_des=rng.normal(size=(2, B)), dcm_vel_des=rng.normal(size=(2, B)),
             zmp=rng.normal(size=(2, B)), zmp_des=rng.normal(size=(2, B)), com=rng.normal(size=(3, B)),
             com_des_xy=rng.normal(size=(2, B)), com_vel_des_xy=rng.normal(size=(2, B)), lf_H=lf, rf_H=rf, lf_H_des=rf, rf_H_des=lf,
             lf_err=rng.normal(size=(6, B)), rf_err=rng.normal(size=(6, B)))
    rows = oracle.log_pack(B, **f)
    assert rows.shape == (53, B) and len(wc.LOG_NAMES) == 53
    names = list(wc.LOG_NAMES)
    assert np.array_equal(rows[names.index("dcm_x"):names.index("dcm_x") + 2], f["dcm"])
    assert np.array_equal(rows[names.index("com_x"):names.index("com_x") + 3], f["com"])
    assert np.array_equal(rows[names.index("lf_x"):names.index("lf_x") + 3], pos_l)
    assert np.allclose(rows[names.index("lf_roll"):names.index("lf_roll") + 3], rpy_l, rtol=0, atol=1e-14)   # asRPY inverts RPY
    assert np.allclose(rows[names.index("rf_des_roll"):names.index("rf_des_roll") + 3], rpy_l, rtol=0, atol=1e-14)
    assert np.array_equal(rows[names.index("rf_err_x"):], f["rf_err"])
    # missing inputs -> zero columns
    part = oracle.log_pack(B, dcm=f["dcm"])
    assert np.array_equal(part[:2], f["dcm"]) and not part[2:].any()
    # file format of the logger module: "time " + names, then "t v ... " rows in ostream default formatting
    path = tmp_path / "Dataset_test.txt"
    T = 3
    wc.write_dataset(path, [0.01 * k for k in range(T)], [rows[:, 0]] * T)
    lines = path.read_text().splitlines()
    assert lines[0] == "time " + " ".join(names) + " "
    assert len(lines) == 1 + T
    vals = lines[2].split()
    assert len(vals) == 54 and vals[0] == "0.01"
    assert np.allclose([float(v) for v in vals[1:]], rows[:, 0], rtol=1e-5, atol=1e-12)   # 6 significant digits


@pytest.mark.gpu
def test_rank4_gpu_parity(wc, oracle, cuda_device):
    rng = np.random.default_rng(3)
    B = 3000
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda_device)
    cfg = wc.glue_default_config()
    gg = wc.GlueBatch(cfg, B)
    omega = math.sqrt(cfg.gravity / cfg.com_height)
    xd, vd, x = rng.normal(size=(2, B)) * 0.1, rng.normal(size=(2, B)) * 0.1, rng.normal(size=(2, B)) * 0.1
    k = rng.uniform(0.5, 2.0, size=B)
    ug = gg.reactive_dcm(1.5, d(xd), d(vd), d(x)).cpu().numpy()
    assert np.abs(ug - oracle.reactive_dcm(omega, 1.5, xd, vd, x)).max() <= 1e-13
    ug = gg.reactive_dcm(d(k), d(xd), d(vd), d(x)).cpu().numpy()
    assert np.abs(ug - oracle.reactive_dcm(omega, k, xd, vd, x)).max() <= 1e-13
    # host-pointer entry point used by the drop-in class (batch of one)
    g1 = wc.GlueBatch(cfg, 1)
    out = np.zeros((2, 1))
    a, b, c = np.ascontiguousarray(xd[:, :1]), np.ascontiguousarray(vd[:, :1]), np.ascontiguousarray(x[:, :1])
    st = wc.lib().wc_glue_reactive_dcm_host(g1.h, 1.5, None, a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data)
    assert st == 0 and np.abs(out - oracle.reactive_dcm(omega, 1.5, a, b, c)).max() <= 1e-13
    g1.close()
    com, cv = rng.normal(size=(3, B)), rng.normal(size=(3, B))
    assert np.abs(gg.evaluate_dcm(d(com), d(cv)).cpu().numpy() - oracle.evaluate_dcm(omega, com, cv)).max() <= 1e-13
    lf, _, _ = _poses(rng, B)
    rf, _, _ = _poses(rng, B)
    lw = rng.normal(size=(6, B)); rw = rng.normal(size=(6, B))
    lw[2] = rng.uniform(-5, 200, size=B); rw[2] = rng.uniform(-5, 200, size=B)
    zo, oko = oracle.evaluate_zmp(lw, rw, lf, rf)
    zg, okg = gg.evaluate_zmp(d(lw), d(rw), d(lf), d(rf))
    assert np.array_equal(okg.cpu().numpy(), oko) and 0 < oko.sum() < B
    assert np.abs(zg.cpu().numpy() - zo).max() <= 1e-12
    f = dict(dcm=x, dcm_des=xd, dcm_vel_des=vd, zmp=zo, zmp_des=ug, com=com, com_des_xy=com[:2], com_vel_des_xy=cv[:2],
             lf_H=lf, rf_H=rf, lf_H_des=rf, rf_H_des=lf, lf_err=lw, rf_err=rw)
    ro = oracle.log_pack(B, **f)
    rg = gg.log_pack(**{k_: d(v) for k_, v in f.items()}).cpu().numpy()
    assert np.abs(rg - ro).max() <= 1e-13
    rg = gg.log_pack(dcm=d(x)).cpu().numpy()
    assert np.array_equal(rg[:2], x) and not rg[2:].any()
    gg.close()


def test_evaluate_com_filters_closed_form(oracle):
    """CoM velocity = J_com [0; dq]; the low-pass filters have unit DC gain, start at (0, 0, com_height) / 0 and follow the
    bilinear-transform recurrence; use_filters only selects the outputs (the filter state advances either way)"""
    rng = np.random.default_rng(4)
    B, nJ, nV = 6, 23, 29
    cfg = oracle.glue_default_config()
    g = oracle.GlueBatch(cfg, B)
    omega = math.sqrt(cfg.gravity / cfg.com_height)
    J = rng.normal(size=(3 * nV, B)); dq = rng.normal(size=(nJ, B)); com = rng.normal(size=(3, B)) * 0.1 + np.array([[0.0], [0.0], [0.5]])
    o = g.evaluate_com(com, J, dq)
    v = np.einsum("rjb,jb->rb", J.reshape(3, nV, B)[:, 6:], dq)
    assert np.allclose(o["com_vel"], v, rtol=1e-13, atol=1e-14) and np.array_equal(o["com"], com)
    assert np.allclose(o["dcm"], (com + v / omega)[:2], rtol=1e-13, atol=1e-14)
    fc, Ts = 10.0, cfg.sampling_time
    tau = 1 / (2 * math.pi * fc)
    y = np.zeros((3, B)); y[2] = cfg.com_height; u = y.copy(); yv = np.zeros((3, B)); uv = np.zeros((3, B))
    g.reset_com_filter()
    for k in range(60):
        o = g.evaluate_com(com, J, dq, cut_frequency=fc, use_filters=True)
        y = (Ts * (com + u) - (Ts - 2 * tau) * y) / (2 * tau + Ts); u = com
        yv = (Ts * (v + uv) - (Ts - 2 * tau) * yv) / (2 * tau + Ts); uv = v
        assert np.allclose(o["com"], y, rtol=1e-12, atol=1e-13) and np.allclose(o["com_vel"], yv, rtol=1e-12, atol=1e-13)
    assert np.abs(o["com"] - com).max() < 0.05 * np.abs(com - np.array([[0], [0], [cfg.com_height]])).max() + 1e-6   # settles on the input (DC gain 1)
    raw = g.evaluate_com(com, J, dq, cut_frequency=fc, use_filters=False)
    assert np.array_equal(raw["com"], com)


@pytest.mark.gpu
def test_evaluate_com_gpu_parity(wc, oracle, cuda_device):
    rng = np.random.default_rng(5)
    B, nJ, nV = 2000, 23, 29
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda_device)
    go = oracle.GlueBatch(oracle.glue_default_config(), B); gg = wc.GlueBatch(wc.glue_default_config(), B)
    for k in range(12):
        J = rng.normal(size=(3 * nV, B)); dq = rng.normal(size=(nJ, B)); com = rng.normal(size=(3, B)) * 0.1
        if k == 6:
            go.reset_com_filter(); gg.reset_com_filter()
        fc, uf = (10.0, True) if k % 3 else ((0.0, False) if k % 2 else (10.0, False))
        oo = go.evaluate_com(com, J, dq, fc, uf); og = gg.evaluate_com(d(com), d(J), d(dq), fc, uf)
        for key in ("com", "com_vel", "dcm"):
            assert np.abs(og[key].cpu().numpy() - oo[key]).max() <= 1e-12, (k, key)
    g1 = wc.GlueBatch(wc.glue_default_config(), 1)
    co, vo, do = np.zeros((3, 1)), np.zeros((3, 1)), np.zeros((2, 1))
    J1, dq1, c1 = np.ascontiguousarray(J[:, :1]), np.ascontiguousarray(dq[:, :1]), np.ascontiguousarray(com[:, :1])
    st = wc.lib().wc_glue_evaluate_com_host(g1.h, c1.ctypes.data, J1.ctypes.data, dq1.ctypes.data, 0.0, 0, co.ctypes.data, vo.ctypes.data, do.ctypes.data)
    ref = oracle.GlueBatch(oracle.glue_default_config(), 1).evaluate_com(c1, J1, dq1)
    assert st == 0 and np.abs(vo - ref["com_vel"]).max() <= 1e-12 and np.abs(do - ref["dcm"]).max() <= 1e-12
    g1.close(); gg.close()
