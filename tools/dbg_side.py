import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wc = importlib.import_module("walking-controllers_b200")
syn = importlib.import_module("walking-controllers_b200.synthetic")
dev = torch.device("cuda:0")
N, B, T = 25, 65536, 16
g = syn.GaitBatch(B, T, N, seed=3, device=dev)
traj = g.trajectory()
m = wc.MpcBatch(wc.mpc_default_config(N), B)
for t in range(T):
    f = g.feet(t)
    m.tick(f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t), traj[2 * t: 2 * t + 2 * (N + 1)].contiguous())
torch.cuda.synchronize()
