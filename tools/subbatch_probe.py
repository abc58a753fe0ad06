"""Does running the batch as S concurrent sub-batches (one CUDA stream each, every instance still ticks once per step) hide the
latency-bound tail kernels (set-up, phase 2, general QP-IK) and the last partial wave of the big ones?
usage: python tools/subbatch_probe.py [B] [N] [steps]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
wc = importlib.import_module("walking-controllers_b200")
syn = importlib.import_module("walking-controllers_b200.synthetic")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
warm = 6
dev = torch.device("cuda:0")
for S in (1, 2, 4):
    Bs = B // S
    T = warm + steps
    subs = []
    for s in range(S):
        g = syn.GaitBatch(Bs, T, N, seed=0x5EED + 17 * s, device=dev)
        traj = g.trajectory()
        mpc_in = []
        for t in range(T):
            f = g.feet(t)
            mpc_in.append((f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t)))
        icfg = wc.ik_default_config()
        ik = wc.IkBatch(icfg, wc.model_icub_like(), Bs, device=0)
        qreg = torch.tensor(list(icfg.joint_reg_rad[:23]), dtype=torch.float64)
        ik_in = [syn.ik_inputs(Bs, 1000 * s + r, qreg, lambda ql, bH, lif: ik.fk(ql, bH, lif, jacobians=False), device=dev, perturb=0.5) for r in range(4)]
        mpc = wc.MpcBatch(wc.mpc_default_config(N), Bs, device=0)
        subs.append(dict(g=g, traj=traj, mpc_in=mpc_in, ik=ik, ik_in=ik_in, mpc=mpc, st=torch.cuda.Stream(), mo=None, io=None))
    torch.cuda.synchronize()
    def step(t):
        for sb in subs:
            with torch.cuda.stream(sb["st"]):
                a = sb["mpc_in"][t]
                sb["mo"] = sb["mpc"].tick(a[0], a[1], a[2], a[3], a[4], sb["traj"][2 * t: 2 * t + 2 * (N + 1)], out=sb["mo"])
                sb["io"] = sb["ik"].tick(sb["ik_in"][t % 4], out=sb["io"])
    for t in range(warm):
        step(t)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for sb in subs:
        sb["st"].wait_event(e0)
    for k in range(steps):
        step(warm + k)
    for sb in subs:
        ev = torch.cuda.Event(); ev.record(sb["st"]); main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"S={S}: {ms:.4f} ms/step  {B / ms / 1e3:.2f}e6 ticks/s", flush=True)
    for sb in subs:
        sb["mpc"].close(); sb["ik"].close()
