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
