"""Does a concurrent pinned-memory H2D / D2H stream slow the tick's kernels down?  (explains e2e vs value)
usage: python tools/dma_interference_probe.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
wc = importlib.import_module("walking-controllers_b200")
syn = importlib.import_module("walking-controllers_b200.synthetic")
B, N, T = 65536, 25, 40
dev = torch.device("cuda:0")
g = syn.GaitBatch(B, T, N, seed=3, device=dev); traj = g.trajectory()
mpc_in = []
for t in range(T):
    f = g.feet(t); mpc_in.append((f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t)))
icfg = wc.ik_default_config(); ik = wc.IkBatch(icfg, wc.model_icub_like(), B)
qreg = torch.tensor(list(icfg.joint_reg_rad[:23]), dtype=torch.float64)
ik_in = [syn.ik_inputs(B, r, qreg, lambda ql, bH, lif: ik.fk(ql, bH, lif, jacobians=False), device=dev, perturb=0.5) for r in range(4)]
mpc = wc.MpcBatch(wc.mpc_default_config(N), B)
h = torch.empty(54 * 1024 * 1024 // 8, dtype=torch.float64).pin_memory(); d = torch.empty_like(h, device=dev)
h2 = torch.empty(20 * 1024 * 1024 // 8, dtype=torch.float64).pin_memory(); d2 = torch.ones(h2.numel(), dtype=torch.float64, device=dev)
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
mo = io_ = None
def run(dma):
    global mo, io_
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(8, T):
        if dma:
            with torch.cuda.stream(s_up): d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s_dn): h2.copy_(d2, non_blocking=True)
        a = mpc_in[t]
        mo = mpc.tick(a[0], a[1], a[2], a[3], a[4], traj[2 * t: 2 * t + 2 * (N + 1)], out=mo)
        io_ = ik.tick(ik_in[t % 4], out=io_)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (T - 8)
for t in range(8):
    a = mpc_in[t]; mo = mpc.tick(a[0], a[1], a[2], a[3], a[4], traj[2 * t: 2 * t + 2 * (N + 1)], out=mo); io_ = ik.tick(ik_in[t % 4], out=io_)
print("tick ms without DMA %.4f" % run(False))
print("tick ms with 54 MB H2D + 20 MB D2H per tick in flight %.4f" % run(True))
print("tick ms without DMA %.4f" % run(False))
