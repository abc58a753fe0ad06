"""The tick with the FK + QP-IK kinematics part (wc_ik_tick_part_dev) on a side stream beside the MPC block, against the serial
tick: how much of the MPC block's latency-bound head (K1, osqp_setup) and tail (phase 2) it fills.
usage: python tools/split_tick_probe.py [B] [steps]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
wc = importlib.import_module("walking-controllers_b200")
syn = importlib.import_module("walking-controllers_b200.synthetic")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
N, warm = 25, 10
dev = torch.device("cuda:0")
T = warm + steps
g = syn.GaitBatch(B, T, N, seed=0x5EED, device=dev)
traj = g.trajectory()
mpc_in = [(lambda f: (f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t)))(g.feet(t)) for t in range(T)]
icfg = wc.ik_default_config()
ik = wc.IkBatch(icfg, wc.model_icub_like(), B)
qreg = torch.tensor(list(icfg.joint_reg_rad[:23]), dtype=torch.float64)
ik_in = [syn.ik_inputs(B, r, qreg, lambda ql, bH, lif: ik.fk(ql, bH, lif, jacobians=False), device=dev, perturb=0.5) for r in range(4)]
side = torch.cuda.Stream()           # default (= least) priority: the kinematics part only takes what the MPC block leaves free
main = torch.cuda.Stream(priority=-1)
torch.cuda.set_stream(main)
for mode in ("serial", "kinematics beside the MPC", "serial", "kinematics beside the MPC"):
    mpc = wc.MpcBatch(wc.mpc_default_config(N), B)
    mo = io = None
    ev_f, ev_j = torch.cuda.Event(), torch.cuda.Event()
    def step(t):
        global mo, io
        a = mpc_in[t]; inp = ik_in[t % 4]
        if mode == "serial":
            mo = mpc.tick(a[0], a[1], a[2], a[3], a[4], traj[2 * t: 2 * t + 2 * (N + 1)], out=mo)
            io = ik.tick(inp, out=io)
        else:
            ev_f.record(main); side.wait_event(ev_f)
            with torch.cuda.stream(side):
                ik.tick_kinematics(inp); ev_j.record(side)
            mo = mpc.tick(a[0], a[1], a[2], a[3], a[4], traj[2 * t: 2 * t + 2 * (N + 1)], out=mo)
            main.wait_event(ev_j)
            io = ik.tick_solve(inp, out=io)
    for t in range(warm):
        step(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for k in range(steps):
        step(warm + k)
    e1.record(main); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{mode}: {ms:.4f} ms per tick, {B / ms * 1e-3:.2f}e6 ticks/s")
    mpc.close()
