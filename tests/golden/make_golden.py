"""Generates the golden fixtures of tests/golden/ from the oracle (oracle/liboracle.so) with fixed seeds.

The reference repository holds no golden vectors for this path and cannot be built or
imported offline, so these fixtures freeze the ORACLE's outputs ("parity unpinned"): they
guard the oracle against regressions and give the GPU tests a second, file-based comparison point.
Run from the repo root:  python tests/golden/make_golden.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

syn = importlib.import_module("walking-controllers_b200.synthetic")
HERE = os.path.dirname(os.path.abspath(__file__))


def mpc_case(B=12, T=100, N=25, seed=2024):
    g = syn.GaitBatch(B, T, N, seed=seed)
    traj = g.trajectory().numpy()
    mb = O.MpcBatch(O.mpc_default_config(N), B)
    zmp, st, it, fl = [], [], [], []
    for t in range(T):
        f = g.feet(t)
        o = mb.tick(f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(),
                    g.measured_dcm(t).numpy(), np.ascontiguousarray(traj[2 * t: 2 * t + 2 * (N + 1)]))
        zmp.append(o[0].copy()); st.append(o[1].copy()); it.append(o[2].copy()); fl.append(o[3].copy())
    primal = mb.solution()[0]
    np.savez_compressed(os.path.join(HERE, "mpc_n25.npz"), B=B, T=T, N=N, seed=seed, zmp=np.array(zmp), status=np.array(st),
                        iters=np.array(it), flags=np.array(fl), primal_last=primal)


def ik_case(B=48, seed=77):
    m = O.model_icub_like(); cfg = O.ik_default_config()
    qreg = torch.tensor(list(cfg.joint_reg_rad[:23]), dtype=torch.float64)

    def fkb(ql, bH, lif):
        o = O.fk_batch(m, ql.numpy(), bH.numpy(), lif.numpy(), jacobians=False)
        return {k: torch.from_numpy(v) for k, v in o.items()}

    inp = {k: v.numpy() for k, v in syn.ik_inputs(B, seed, qreg, fkb).items()}
    out = O.ik_batch_tick(cfg, m, inp)
    kin = O.fk_batch(m, inp["q_lin"], inp["base_H"], inp["left_is_fixed"])
    np.savez_compressed(os.path.join(HERE, "ik_icub23.npz"), B=B, seed=seed, **{"in_" + k: v for k, v in inp.items()},
                        **{"out_" + k: v for k, v in out.items()}, **{"kin_" + k: v for k, v in kin.items()})


def ik_osqp_case(B=24, T=5, seed=78):
    """WalkingQPIK_osqp over a warm-started sequence on caller-supplied kinematics (the Jacobians are part of the fixture, so the
    Hessian's exact-zero pattern — which decides the solver life cycle — is the same wherever the fixture is replayed)"""
    m = O.model_icub_like(); cfg = O.ik_default_config()
    qreg = torch.tensor(list(cfg.joint_reg_rad[:23]), dtype=torch.float64)

    def fkb(ql, bH, lif):
        o = O.fk_batch(m, ql.numpy(), bH.numpy(), lif.numpy(), jacobians=False)
        return {k: torch.from_numpy(v) for k, v in o.items()}

    inp = {k: v.numpy().copy() for k, v in syn.ik_inputs(B, seed, qreg, fkb).items()}
    rng = np.random.default_rng(seed)
    L = O.lib()
    import ctypes as C
    dp = C.POINTER(C.c_double)
    L.wco_ikosqp_create.argtypes = [C.c_void_p, C.c_void_p]; L.wco_ikosqp_create.restype = C.c_void_p
    L.wco_ikosqp_solve.argtypes = [C.c_void_p] + [dp] * 16 + [dp] * 3 + [C.POINTER(C.c_int)] * 3
    L.wco_ikosqp_setups.argtypes = [C.c_void_p]
    objs = [L.wco_ikosqp_create(C.byref(cfg), None) for _ in range(B)]
    rec = dict(q_meas=[], left_is_fixed=[], kin=[], dq=[], status=[], iters=[], flags=[], setups=[])
    for t in range(T):
        if t == 3:
            inp["left_is_fixed"] = (1 - inp["left_is_fixed"]).astype(np.uint8)
        kin = O.fk_batch(m, inp["q_lin"], inp["base_H"], inp["left_is_fixed"])
        dq = np.zeros((23, B)); st = np.zeros(B, np.int32); it = np.zeros(B, np.int32); fl = np.zeros(B, np.int32); su = np.zeros(B, np.int32)
        for i in range(B):
            col = lambda a: np.ascontiguousarray(a[:, i])
            args = [col(inp["q_meas"]), col(kin["lf_H"]), col(kin["rf_H"]), col(kin["neck_R"]), col(inp["com_meas"]), col(kin["J_lf"]),
                    col(kin["J_rf"]), col(kin["J_neck"]), col(kin["J_com"]), col(inp["lf_H_des"]), col(inp["rf_H_des"]), col(inp["lf_twist"]),
                    col(inp["rf_twist"]), col(inp["com_pos_des"]), col(inp["com_vel_des"]), col(inp["neck_R_des"])]
            d = np.zeros(23); le = np.zeros(6); re = np.zeros(6); a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
            L.wco_ikosqp_solve(objs[i], *[x.ctypes.data_as(dp) for x in args], d.ctypes.data_as(dp), le.ctypes.data_as(dp),
                               re.ctypes.data_as(dp), C.byref(a), C.byref(b), C.byref(c))
            dq[:, i] = d; st[i], it[i], fl[i] = a.value, b.value, c.value; su[i] = L.wco_ikosqp_setups(objs[i])
        rec["q_meas"].append(inp["q_meas"].copy()); rec["left_is_fixed"].append(inp["left_is_fixed"].copy())
        rec["kin"].append(np.concatenate([kin[k] for k in KIN_KEYS], 0))
        rec["dq"].append(dq); rec["status"].append(st); rec["iters"].append(it); rec["flags"].append(fl); rec["setups"].append(su)
        inp["q_meas"] = inp["q_meas"] + 0.002 * rng.normal(size=inp["q_meas"].shape)
    static = {k: inp[k] for k in inp if k not in ("q_meas", "left_is_fixed")}
    np.savez_compressed(os.path.join(HERE, "ik_osqp_seq.npz"), B=B, T=T, seed=seed, **{"in_" + k: v for k, v in static.items()},
                        **{"seq_" + k: np.array(v) for k, v in rec.items()})


KIN_KEYS = ("lf_H", "rf_H", "neck_R", "com", "J_lf", "J_rf", "J_neck", "J_com")
KIN_ROWS = (12, 12, 9, 3, 6 * 29, 6 * 29, 6 * 29, 3 * 29)


def traj_glue_case(B=16, T=200, N=25, seed=79):
    """input producer (oracle/traj_ref.c) and the stateless rank-4 glue on its signals"""
    g = syn.GaitBatch(B, T, N, seed=seed)
    par = dict(step_len=g.step_len.numpy(), phase=g.phase.numpy().astype(np.int32), dtheta=g.dtheta.numpy(),
               theta0=g.theta0.numpy(), first_left=g.first_left.numpy().astype(np.uint8))
    tw = O.TrajBatch(O.traj_default_config(T, N), B); tw.set_gait(**par)
    pos, vel = tw.dcm(0, T + N + 1)
    ticks = (0, 36, 37, 89, 90, 133, 199)
    feet = [tw.feet(t) for t in ticks]
    meas = np.array([tw.measured_dcm(t, 0.002, 0x5EED) for t in ticks])
    omega = float(np.sqrt(9.81 / 0.53))
    zr = np.array([O.reactive_dcm(omega, 1.5, pos[2 * t: 2 * t + 2], vel[2 * t: 2 * t + 2], meas[k]) for k, t in enumerate(ticks)])
    rows = np.array([O.log_pack(B, dcm=meas[k], dcm_des=pos[2 * t: 2 * t + 2], dcm_vel_des=vel[2 * t: 2 * t + 2], zmp_des=zr[k],
                                lf_H=feet[k]["lf_H"], rf_H=feet[k]["rf_H"], lf_err=feet[k]["lf_twist"], rf_err=feet[k]["rf_twist"])
                     for k, t in enumerate(ticks)])
    np.savez_compressed(os.path.join(HERE, "traj_glue.npz"), B=B, T=T, N=N, seed=seed, ticks=np.array(ticks), dcm_pos=pos, dcm_vel=vel,
                        meas=meas, zmp_reactive=zr, log_rows=rows, **{"par_" + k: v for k, v in par.items()},
                        **{"feet_" + k: np.array([f[k] for f in feet]) for k in feet[0]})


if __name__ == "__main__":
    O.build()
    if "--new-only" not in sys.argv:
        mpc_case()
        ik_case()
    ik_osqp_case()
    traj_glue_case()
    print("golden fixtures written to", HERE)
