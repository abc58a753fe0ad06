"""Small workload touching every kernel family once (a quick all-kernels exercise; written for compute-sanitizer, which is
closed on the GPU pool of this project, so it only serves as a fast end-to-end run):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
wc = importlib.import_module("walking-controllers_b200")
syn = importlib.import_module("walking-controllers_b200.synthetic")
dev = torch.device("cuda:0")
B, N, T = 48, 25, 3
g = syn.GaitBatch(B, T, N, seed=1, device=dev)
traj = g.trajectory()
mpc = wc.MpcBatch(wc.mpc_default_config(N), B)
ik = wc.IkBatch(wc.ik_default_config(), wc.model_icub_like(), B)
ik2 = wc.IkBatch(wc.ik_default_config(), wc.model_icub_like(), B); ik2.set_solver(wc.SOLVER_OSQP)
gl = wc.GlueBatch(wc.glue_default_config(), B)
tj = wc.TrajBatch(wc.traj_default_config(T, N), B)
tj.set_gait(g.step_len, g.phase.to(torch.int32), g.dtheta, g.theta0, g.first_left.to(torch.uint8))
qreg = torch.tensor(list(wc.ik_default_config().joint_reg_rad[:23]), dtype=torch.float64)
inp = syn.ik_inputs(B, 5, qreg, lambda ql, bH, lif: ik.fk(ql, bH, lif, jacobians=False), device=dev, perturb=0.5)
for t in range(T):
    f = tj.feet(t); ref, vel = tj.dcm(t, N + 1); dcm = tj.measured_dcm(t, 0.002, 7)
    mo = mpc.tick(f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], dcm, ref)
    o1 = ik.tick(inp); o2 = ik2.tick(inp)
    k = ik.fk(inp["q_lin"], inp["base_H"], inp["left_is_fixed"])
    c = gl.com_reference(ref[0:2], mo["zmp"], mo["zmp"], k["com"])
    gl.integrate_joints(o1["dq"]); gl.reactive_dcm(1.5, ref[0:2], vel[0:2], dcm)
    gl.evaluate_com(k["com"], k["J_com"], o1["dq"], 10.0, True)
    gl.log_pack(dcm=dcm, lf_H=k["lf_H"], rf_H=k["rf_H"], lf_err=o1["lf_err"], rf_err=o1["rf_err"])
torch.cuda.synchronize()
print("ok", int((mo["status"] == 1).sum()), int((o1["status"] == 0).sum()), int((o2["status"] == 1).sum()))
