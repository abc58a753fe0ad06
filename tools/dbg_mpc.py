"""debug: compare the chain path, the cooperative path and the oracle on small MPC cases"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
wc = importlib.import_module("walking-controllers_b200")
syn = importlib.import_module("walking-controllers_b200.synthetic")
from oracle import oracle as O
O.build()
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T = int(sys.argv[3]) if len(sys.argv) > 3 else 6
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 100 + N
g = syn.GaitBatch(B, T, N, seed=seed)
traj = g.trajectory()
mo = O.MpcBatch(O.mpc_default_config(N), B)
os.environ["WC_MPC_PATH"] = "coop"; mc = wc.MpcBatch(wc.mpc_default_config(N), B)
os.environ.pop("WC_MPC_PATH"); mg = wc.MpcBatch(wc.mpc_default_config(N), B)
for t in range(T):
    f = g.feet(t); dcm = g.measured_dcm(t); ref = traj[2 * t: 2 * t + 2 * (N + 1)].contiguous()
    a = (f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(), dcm.numpy(), ref.numpy())
    zo, so, io_, fo = mo.tick(*a)
    oc = mc.tick_host(*a); og = mg.tick_host(*a)
    pc = mc.debug()[0]; pg = mg.debug()[0]; po = mo.solution()[0]
    bad = np.nonzero((og["status"] != so) | (og["iters"] != io_))[0]
    print(f"tick {t}: oracle iters {np.unique(io_, return_counts=True)} chain-vs-oracle primal {np.abs(pg-po).max():.3e} coop-vs-oracle {np.abs(pc-po).max():.3e} "
          f"chain-vs-coop {np.abs(pg-pc).max():.3e} zmp {np.abs(og['zmp']-zo).max():.3e} bad {bad.tolist()}")
    for i in bad[:4]:
        print("   inst", i, "chain status/iters", og["status"][i], og["iters"][i], "oracle", so[i], io_[i], "coop", oc["status"][i], oc["iters"][i],
              "contacts", a[2][i], a[3][i], "flags", og["flags"][i], fo[i])
        print("   primal chain", pg[:, i][:12]); print("   primal orac ", po[:, i][:12])
