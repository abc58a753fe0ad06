"""overlapped set-up (WC_MPC_SETUP_OVERLAP) against the serial launch sequence: per tick, which instances differ"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wc = importlib.import_module("walking-controllers_b200")
syn = importlib.import_module("walking-controllers_b200.synthetic")
dev = torch.device("cuda:0")
N, T = 25, 6
for B in (64, 4096, 65536):
    g = syn.GaitBatch(B, T, N, seed=0xB200, device=dev)
    traj = g.trajectory()
    runs = {}
    for name, ov in (("serial", "0"), ("overlap", "1"), ("overlap2", "1")):
        os.environ["WC_MPC_SETUP_OVERLAP"] = ov
        m = wc.MpcBatch(wc.mpc_default_config(N), B)
        outs = []
        for t in range(T):
            f = g.feet(t)
            o = m.tick(f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t), traj[2 * t: 2 * t + 2 * (N + 1)])
            torch.cuda.synchronize()
            outs.append({k: v.clone() for k, v in o.items()})
        m.close()
        runs[name] = outs
    for t in range(T):
        a, b, c = runs["serial"][t], runs["overlap"][t], runs["overlap2"][t]
        dz = (a["zmp"] - b["zmp"]).abs().amax(0)
        bad = (dz > 0) | (a["iters"] != b["iters"]) | (a["status"] != b["status"]) | (a["flags"] != b["flags"])
        dz2 = (b["zmp"] - c["zmp"]).abs().amax(0)
        idx = bad.nonzero().flatten()[:8].tolist()
        print(f"B={B} t={t}: serial vs overlap differ on {int(bad.sum())} (max dz {dz.max().item():.3e}); overlap vs overlap: {int((dz2 > 0).sum())}; first {idx}",
              [(int(a['iters'][i]), int(b['iters'][i]), int(a['flags'][i]), int(b['flags'][i]), float(dz[i])) for i in idx[:4]])

# timing: steady-state ticks at B = 65536, serial against overlapped set-up
B, T = 65536, 60
g = syn.GaitBatch(B, T, N, seed=3, device=dev)
traj = g.trajectory()
ins = [(g.feet(t), g.measured_dcm(t), traj[2 * t: 2 * t + 2 * (N + 1)].contiguous()) for t in range(T)]
for ov in ("0", "1", "0", "1"):
    os.environ["WC_MPC_SETUP_OVERLAP"] = ov
    m = wc.MpcBatch(wc.mpc_default_config(N), B)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for t in range(T):
        if t == 20: e0.record()
        f, dcm, ref = ins[t]
        m.tick(f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], dcm, ref)
    e1.record(); torch.cuda.synchronize()
    print(f"overlap={ov}: {e0.elapsed_time(e1) / (T - 20):.4f} ms per MPC tick")
    m.close()
