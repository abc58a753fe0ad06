"""The double-buffered host-pointer entry points (wc_*_tick_host_async / _sync_one / _sync) give the same results as
the synchronous ones, with two ticks in flight."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_mpc_and_ik_host_pipeline_matches_sync_api(wc, syn, oracle, cuda_device):
    B, N, T = 300, 25, 9
    g = syn.GaitBatch(B, T, N, seed=31)
    traj = g.trajectory()
    ins = []
    for t in range(T):
        f = g.feet(t)
        ins.append(tuple(x.contiguous().pin_memory() for x in (f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t),
                                                               traj[2 * t: 2 * t + 2 * (N + 1)])))
    ms, ma = wc.MpcBatch(wc.mpc_default_config(N), B), wc.MpcBatch(wc.mpc_default_config(N), B)
    ref = [ms.tick_host(*[x.numpy() for x in a]) for a in ins]
    ref = [{k: v.copy() for k, v in r.items()} for r in ref]
    outs = [dict(zmp=torch.empty(2, B, dtype=torch.float64).pin_memory(), status=torch.empty(B, dtype=torch.int32).pin_memory(),
                 iters=torch.empty(B, dtype=torch.int32).pin_memory(), flags=torch.empty(B, dtype=torch.int32).pin_memory())
            for _ in range(2)]
    got = []
    for t in range(T):
        ma.tick_host_async(*ins[t], out=outs[t % 2])
        if t >= 1:
            ma.sync_one()
            got.append({k: v.numpy().copy() for k, v in outs[(t - 1) % 2].items()})
    ma.sync()
    got.append({k: v.numpy().copy() for k, v in outs[(T - 1) % 2].items()})
    for t in range(T):
        for k in ("zmp", "status", "iters", "flags"):
            assert np.array_equal(got[t][k], ref[t][k]), (t, k)
    ms.close(); ma.close()

    # ---- IK (stateless): four in-flight-rotated input sets
    icfg, model = wc.ik_default_config(), wc.model_icub_like()
    qreg = torch.tensor(list(icfg.joint_reg_rad[:23]), dtype=torch.float64)
    mo = oracle.model_icub_like()

    def fkb(ql, bH, lif):
        o = oracle.fk_batch(mo, ql.numpy(), bH.numpy(), lif.numpy(), jacobians=False)
        return {k: torch.from_numpy(v) for k, v in o.items()}

    sets = [{k: v.pin_memory() for k, v in syn.ik_inputs(B, 50 + r, qreg, fkb, perturb=0.5).items()} for r in range(4)]
    ik = wc.IkBatch(icfg, model, B)
    refs = [{k: v.copy() for k, v in ik.tick_host({k: v.numpy() for k, v in s.items()}).items()} for s in sets]
    outs = [dict(dq=torch.empty(23, B, dtype=torch.float64).pin_memory(), lf_err=torch.empty(6, B, dtype=torch.float64).pin_memory(),
                 rf_err=torch.empty(6, B, dtype=torch.float64).pin_memory(), status=torch.empty(B, dtype=torch.int32).pin_memory(),
                 nwsr=torch.empty(B, dtype=torch.int32).pin_memory()) for _ in range(2)]
    got = []
    for t in range(8):
        ik.tick_host_async(sets[t % 4], out=outs[t % 2])
        if t >= 1:
            ik.sync_one()
            got.append({k: v.numpy().copy() for k, v in outs[(t - 1) % 2].items()})
    ik.sync()
    got.append({k: v.numpy().copy() for k, v in outs[1].items()})
    for t in range(8):
        for k in ("dq", "status", "nwsr"):
            assert np.array_equal(got[t][k], refs[t % 4][k]), (t, k)
    ik.close()


def test_reference_ring_matches_full_window(wc, syn, cuda_device):
    """wc_mpc_tick_ring_dev (one new DCM sample per tick, device-side ring; the whole window on the first tick and after a
    re-plan) is bit-identical to wc_mpc_tick_dev fed the whole window every tick, across support-phase changes (new solver:
    the gradient is rebuilt from all N + 1 ring slots) and a resetTrajectory tick."""
    B, N, T = 257, 25, 130
    g = syn.GaitBatch(B, T, N, seed=41, device=cuda_device)
    traj = g.trajectory()
    mw, mr = wc.MpcBatch(wc.mpc_default_config(N), B), wc.MpcBatch(wc.mpc_default_config(N), B)
    new = 0
    for t in range(T):
        f = g.feet(t)
        win = traj[2 * t: 2 * t + 2 * (N + 1)]
        if t == 70:  # a re-plan: the window no longer continues the previous one; both sides pass resetTrajectory
            traj = traj + 0.01
            win = traj[2 * t: 2 * t + 2 * (N + 1)]
        reset = torch.ones(B, dtype=torch.uint8, device=cuda_device) if t == 70 else None
        a = (f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t))
        ow = mw.tick(*a, win.contiguous(), reset=reset)
        if t == 0 or t == 70:
            orr = mr.tick_ring(*a, ref_window=win.contiguous(), reset=reset)
        else:
            orr = mr.tick_ring(*a, ref_newest=win[2 * N: 2 * N + 2].contiguous(), reset=reset)
        torch.cuda.synchronize()
        for k in ("zmp", "status", "iters", "flags"):
            assert torch.equal(ow[k], orr[k]), (t, k)
        new += int(((ow["flags"] & 16) != 0).sum())
    assert new > B  # solvers were rebuilt after the first tick: the full-ring read path ran
    mw.close(); mr.close()


@pytest.mark.parametrize("arena", [False, True])
def test_fused_host_tick_matches_the_two_blocks(wc, syn, oracle, cuda_device, arena):
    """wc_tick_host_async (one upload for both blocks: shared feet transforms, stance base selected on the device, one DCM
    sample per tick) gives the results of wc_mpc_tick_host + wc_ik_tick_host on the same data, two ticks in flight.
    arena: the tick's inputs / results live in ONE pinned arena each, in the packed order (copies merged into two uploads and
    one download); otherwise in separately allocated buffers (one copy per array).  B = 203: the byte flags are padded."""
    B, N, T = 203 if arena else 200, 25, 12
    g = syn.GaitBatch(B, T, N, seed=43)
    traj = g.trajectory()
    icfg, model = wc.ik_default_config(), wc.model_icub_like()
    qreg = torch.tensor(list(icfg.joint_reg_rad[:23]), dtype=torch.float64)
    mo = oracle.model_icub_like()

    def fkb(ql, bH, lif):
        o = oracle.fk_batch(mo, ql.numpy(), bH.numpy(), lif.numpy(), jacobians=False)
        return {k: torch.from_numpy(v) for k, v in o.items()}

    ms, mf = wc.MpcBatch(wc.mpc_default_config(N), B), wc.MpcBatch(wc.mpc_default_config(N), B)
    ik_s, ik_f = wc.IkBatch(icfg, model, B), wc.IkBatch(icfg, model, B)
    pin = lambda x: x.contiguous().pin_memory()
    outs = [dict(zmp=pin(torch.empty(2, B, dtype=torch.float64)), osqp_status=pin(torch.empty(B, dtype=torch.int32)),
                 iters=pin(torch.empty(B, dtype=torch.int32)), flags=pin(torch.empty(B, dtype=torch.int32)),
                 dq=pin(torch.empty(23, B, dtype=torch.float64)), lf_err=pin(torch.empty(6, B, dtype=torch.float64)),
                 rf_err=pin(torch.empty(6, B, dtype=torch.float64)), qp_status=pin(torch.empty(B, dtype=torch.int32)),
                 nwsr=pin(torch.empty(B, dtype=torch.int32))) for _ in range(2)]
    if arena:
        outs = [wc.host_tick_arena_out(B, 23) for _ in range(2)]
    refs, got = [], []
    for t in range(T):
        f = g.feet(t)
        win = traj[2 * t: 2 * t + 2 * (N + 1)].contiguous()
        dcm = g.measured_dcm(t)
        # IK inputs of the tick about the gait's own feet: desired feet = the MPC's feet, base = the stance foot
        base = torch.where(f["left_is_fixed"][None, :] != 0, f["lf_H"], f["rf_H"]).contiguous()
        ii = syn.ik_inputs(B, 60 + t, qreg, fkb, perturb=0.5)
        ii.update(base_H=base, left_is_fixed=f["left_is_fixed"], lf_H_des=f["lf_H"], rf_H_des=f["rf_H"],
                  lf_twist=f["lf_twist"], rf_twist=f["rf_twist"])
        rm = ms.tick_host(f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(), dcm.numpy(), win.numpy())
        ri = ik_s.tick_host({k: np.ascontiguousarray(v.numpy()) for k, v in ii.items()})
        refs.append(({k: v.copy() for k, v in rm.items()}, {k: v.copy() for k, v in ri.items()}))
        inp = dict(lf_H=pin(f["lf_H"]), rf_H=pin(f["rf_H"]), l_contact=pin(f["l_contact"]), r_contact=pin(f["r_contact"]),
                   left_is_fixed=pin(f["left_is_fixed"]), dcm_meas=pin(dcm), q_meas=pin(ii["q_meas"]), q_lin=pin(ii["q_lin"]),
                   lf_twist=pin(f["lf_twist"]), rf_twist=pin(f["rf_twist"]), com_pos_des=pin(ii["com_pos_des"]),
                   com_vel_des=pin(ii["com_vel_des"]), com_meas=pin(ii["com_meas"]), neck_R_des=pin(ii["neck_R_des"]))
        if t == 0:
            inp["ref_window"] = pin(win)
        else:
            inp["ref_newest"] = pin(win[2 * N: 2 * N + 2])
        if arena:
            packed = wc.host_tick_arena_in(B, N, 23, window=(t == 0))
            for k, v in inp.items():
                packed[k].copy_(v.reshape(packed[k].shape))
            inp = packed
        mf.tick_fused_host_async(ik_f, inp, outs[t % 2])
        if t >= 1:
            mf.sync_one()
            got.append({k: v.numpy().copy() for k, v in outs[(t - 1) % 2].items() if k != "_arena"})
    mf.sync()
    got.append({k: v.numpy().copy() for k, v in outs[(T - 1) % 2].items() if k != "_arena"})
    for t in range(T):
        rm, ri = refs[t]
        for a, b in (("zmp", "zmp"), ("osqp_status", "status"), ("iters", "iters"), ("flags", "flags")):
            assert np.array_equal(got[t][a], rm[b]), (t, a)
        for a, b in (("dq", "dq"), ("lf_err", "lf_err"), ("rf_err", "rf_err"), ("qp_status", "status"), ("nwsr", "nwsr")):
            assert np.array_equal(got[t][a], ri[b]), (t, a)
    ms.close(); mf.close(); ik_s.close(); ik_f.close()


def test_wrappers_reject_bad_buffers(wc, syn, cuda_device):
    """the ctypes view validates dtype, contiguity, size and side of the bus before a raw pointer reaches the C ABI"""
    B, N = 8, 10
    g = syn.GaitBatch(B, 2, N, seed=1, device=cuda_device)
    traj = g.trajectory()
    f = g.feet(0)
    m = wc.MpcBatch(wc.mpc_default_config(N), B)
    a = [f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(0), traj[0: 2 * (N + 1)].contiguous()]
    with pytest.raises(TypeError):
        m.tick(a[0], a[1], a[2].to(torch.int64), a[3], a[4], a[5])          # an int64 contact mask
    with pytest.raises(ValueError):
        m.tick(a[0].t().contiguous().t(), a[1], a[2], a[3], a[4], a[5])     # a transposed view
    with pytest.raises(ValueError):
        m.tick(a[0].cpu(), a[1], a[2], a[3], a[4], a[5])                    # a CPU tensor to a *_dev call
    with pytest.raises(ValueError):
        m.tick(a[0], a[1], a[2], a[3], a[4], a[5][:-2])                     # a short reference window
    short = dict(zmp=torch.empty(2, B - 1, dtype=torch.float64, device=cuda_device), status=torch.empty(B, dtype=torch.int32, device=cuda_device),
                 iters=torch.empty(B, dtype=torch.int32, device=cuda_device), flags=torch.empty(B, dtype=torch.int32, device=cuda_device))
    with pytest.raises(ValueError):
        m.tick(*a, out=short)                                                # an undersized output buffer
    m.tick(*a)
    m.close()
