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


if __name__ == "__main__":
    O.build()
    mpc_case()
    ik_case()
    print("golden fixtures written to", HERE)
