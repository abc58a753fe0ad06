"""Full-size (BASELINE configs[2]: B = 65536, N = 25) checks through size-independent properties.

The oracle cannot run 65536 instances in seconds, so at full size the CUDA path is checked by
  * the oracle on a random subset of the batch (instances are independent: a column subset is a valid batch),
  * agreement of the three independently written ADMM kernel families (split-chain, chain-per-thread, cooperative) and of the two
    IK paths (stance-foot elimination + Schur solve vs null-space / active-set) on every instance,
  * problem-level invariants: ZMP inside the support polygon + tolerance, OSQP solved everywhere, IK task
    residuals ~ 0, joint velocity bounds respected,
  * bitwise determinism of a repeated run.
Calls go through the C ABI (wc_mpc_tick_dev / wc_ik_tick_dev)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, N, T = 65536, 25, 3
SUB = 192


def _mpc_run(wc, g, traj_d, dev, path, monkeypatch):
    if path == "split":   # what wc_mpc_create picks at N = 25
        monkeypatch.delenv("WC_MPC_PATH", raising=False)
    else:
        monkeypatch.setenv("WC_MPC_PATH", path)
    m = wc.MpcBatch(wc.mpc_default_config(N), B)
    assert m.path == path
    outs = []
    for t in range(T):
        f = g.feet(t)
        o = m.tick(f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t), traj_d[2 * t: 2 * t + 2 * (N + 1)])
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in o.items()})
    m.close()
    return outs


def test_mpc_full_size(wc, syn, oracle, cuda_device, monkeypatch):
    g = syn.GaitBatch(B, T, N, seed=0xB200, device=cuda_device)
    traj = g.trajectory()
    chain = _mpc_run(wc, g, traj, cuda_device, "split", monkeypatch)
    chain2 = _mpc_run(wc, g, traj, cuda_device, "split", monkeypatch)
    coop = _mpc_run(wc, g, traj, cuda_device, "coop", monkeypatch)
    third = _mpc_run(wc, g, traj, cuda_device, "chain", monkeypatch)
    rng = np.random.default_rng(1)
    idx = np.sort(rng.choice(B, SUB, replace=False))
    idx_t = torch.from_numpy(idx).to(cuda_device)
    mo = oracle.MpcBatch(oracle.mpc_default_config(N), SUB)
    for t in range(T):
        c, c2, k = chain[t], chain2[t], coop[t]
        # determinism: the same launch sequence gives the same bits (atomics only order the work lists)
        for key in ("zmp", "status", "iters", "flags"):
            assert torch.equal(c[key], c2[key]), f"tick {t}: {key} not deterministic"
        # three independently written kernel families (split-chain, cooperative, chain-per-thread), every instance
        for other in (k, third[t]):
            assert torch.equal(c["status"], other["status"]) and torch.equal(c["iters"], other["iters"]) and torch.equal(c["flags"], other["flags"])
            assert (c["zmp"] - other["zmp"]).abs().max().item() <= 1e-9
        # invariants
        assert (c["status"] == 1).all(), "OSQP did not solve every instance of the nominal workload"
        assert ((c["flags"] & (4 | 8)) == 0).all(), "a solved instance left the support polygon + tolerance"
        # oracle on a column subset
        f = g.feet(t)
        sub = lambda x: x[..., idx_t].cpu().numpy() if x.dim() > 1 else x[idx_t].cpu().numpy()
        zo, so, io_, fo = mo.tick(sub(f["lf_H"]), sub(f["rf_H"]), sub(f["l_contact"]), sub(f["r_contact"]),
                                  sub(g.measured_dcm(t)), np.ascontiguousarray(sub(traj[2 * t: 2 * t + 2 * (N + 1)])))
        zg = c["zmp"][:, idx_t].cpu().numpy()
        assert np.array_equal(c["status"][idx_t].cpu().numpy(), so)
        assert np.array_equal(c["iters"][idx_t].cpu().numpy(), io_)
        assert np.array_equal(c["flags"][idx_t].cpu().numpy(), fo)
        assert (np.abs(zg - zo) / np.maximum(1.0, np.abs(zo))).max() <= 1e-6


def test_ik_full_size(wc, syn, oracle, cuda_device, monkeypatch):
    cfg = wc.ik_default_config()
    model = wc.model_icub_like()
    qreg = torch.tensor(list(cfg.joint_reg_rad[:23]), dtype=torch.float64)
    monkeypatch.delenv("WC_IK_PATH", raising=False)
    ik = wc.IkBatch(cfg, model, B)
    inp = syn.ik_inputs(B, 77, qreg, lambda ql, bH, lif: ik.fk(ql, bH, lif, jacobians=False), device=cuda_device, perturb=0.5)
    fast = {k: v.clone() for k, v in ik.tick(inp).items()}
    torch.cuda.synchronize()
    ik.close()
    monkeypatch.setenv("WC_IK_PATH", "general")
    ikg = wc.IkBatch(cfg, model, B)
    gen = {k: v.clone() for k, v in ikg.tick(inp).items()}
    torch.cuda.synchronize()
    ikg.close()
    assert torch.equal(fast["status"], gen["status"]) and torch.equal(fast["nwsr"], gen["nwsr"])
    assert (fast["status"] == 0).all()
    assert (fast["dq"] - gen["dq"]).abs().max().item() <= 1e-8
    # task rows hold (equality constraints), velocity bounds hold
    assert fast["lf_err"].abs().max().item() <= 1e-9 and fast["rf_err"].abs().max().item() <= 1e-9
    vmax = torch.tensor(list(cfg.max_vel[:23]), dtype=torch.float64, device=cuda_device)[:, None]
    vmin = torch.tensor(list(cfg.min_vel[:23]), dtype=torch.float64, device=cuda_device)[:, None]
    assert (fast["dq"] <= vmax + 1e-9).all() and (fast["dq"] >= vmin - 1e-9).all()
    assert 0.02 < (fast["nwsr"] > 0).double().mean().item() < 0.5  # the workload exercises the active-set path
    # oracle on a column subset
    rng = np.random.default_rng(2)
    idx = np.sort(rng.choice(B, SUB, replace=False))
    idx_t = torch.from_numpy(idx).to(cuda_device)
    sub = {k: (v[:, idx_t] if v.dim() > 1 else v[idx_t]).cpu().numpy() for k, v in inp.items()}
    oo = oracle.ik_batch_tick(oracle.ik_default_config(), oracle.model_icub_like(), sub)
    assert np.array_equal(fast["status"][idx_t].cpu().numpy(), oo["status"])
    assert np.array_equal(fast["nwsr"][idx_t].cpu().numpy(), oo["nwsr"])
    assert np.abs(fast["dq"][:, idx_t].cpu().numpy() - oo["dq"]).max() <= 1e-8
