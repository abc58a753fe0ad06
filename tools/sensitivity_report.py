"""How far can "within 1e-6 of the restatement" be from the real osqp-eigen + iDynTree stack?  (task 3 ii)

The oracle is PARITY UNPINNED: OSQP, osqp-eigen and iDynTree are not available offline and the reference holds no golden
vectors.  Three things the restatement had to FIX are free in the real stack and change OSQP's iterates (not the QP):
  * the order and scaling of the convex-hull rows (iDynTree::ConvexHullProjectionConstraint::buildConvexHull, external:
    WM/src/WalkingDCMModelPredictiveController.cpp:466,486) -- here: monotone-chain order, un-normalised edge normals;
  * OSQP's adaptive_rho_interval = 0, which real OSQP derives from wall-clock timings (non-deterministic) -- here: 100;
  * the KKT back-end (QDLDL on the quasi-definite KKT matrix vs. any other exact solve) -- rounding only.
This script runs the ORACLE ONLY (CPU) on the benchmark workload under those variations and reports how much the
controller output u_0 (desired ZMP), the OSQP status and the iteration count move.  OSQP stops at eps_abs = eps_rel = 1e-3,
so two exact implementations of the same algorithm that differ in any of the above agree only to about that tolerance:
the spread measured here is the honest error bar on every "1e-6" in this repository.

    python tools/sensitivity_report.py [B] [T]   ->  profiles/r2_sensitivity.md
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

syn = importlib.import_module("walking-controllers_b200.synthetic")


def run(B, T, N, noise, push, **kw):
    g = syn.GaitBatch(B, T, N, seed=0x5EED, noise_dcm=noise)
    traj = g.trajectory().numpy()
    cfg = O.mpc_default_config(N)
    if "rho_interval" in kw:
        cfg.osqp.adaptive_rho_interval = kw["rho_interval"]
    if "eps" in kw:
        cfg.osqp.eps_abs = cfg.osqp.eps_rel = kw["eps"]
    mb = O.MpcBatch(cfg, B, linsys=kw.get("linsys", 0), hull_variant=kw.get("hull_variant", 0))
    rng = np.random.default_rng(7)
    pt = rng.integers(5, T - 5, B); mag = rng.uniform(0, push, B); ang = rng.uniform(0, 2 * np.pi, B)
    Z, S, I = [], [], []
    for t in range(T):
        f = g.feet(t)
        dcm = g.measured_dcm(t).numpy().copy()
        hit = pt == t
        dcm[0, hit] += (mag * np.cos(ang))[hit]; dcm[1, hit] += (mag * np.sin(ang))[hit]
        z, st, it, fl = mb.tick(f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(), dcm,
                                np.ascontiguousarray(traj[2 * t: 2 * t + 2 * (N + 1)]), threads=os.cpu_count())
        Z.append(z.copy()); S.append(st.copy()); I.append(it.copy())
    return np.array(Z), np.array(S), np.array(I)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    N = 25
    out = ["# Sensitivity of the oracle's MPC output to what the restatement had to fix (r2)", "",
           f"Oracle only (CPU), benchmark workload: B = {B} instances x T = {T} ticks, N = {N}, DCM noise 2 mm, an impulsive DCM push",
           "U(0, 0.06) m at a random tick (BASELINE configs[3]), OSQP defaults (eps_abs = eps_rel = 1e-3).  Open loop: every variant",
           "sees the same measured DCM; its own u_prev feeds its next gradient (MPCSolver.cpp:244-245), so differences accumulate",
           "as they would on the robot.  `rel` = |u0 - u0_base| / max(1, |u0_base|) over all instance-ticks and both axes.", "",
           "| variant | max rel | p99 rel | median rel | status != base | iterations != base | iteration histogram |",
           "|---|---|---|---|---|---|---|"]
    base = run(B, T, N, 0.002, 0.06)
    variants = [("baseline repeated (determinism)", {}),
                ("hull rows rotated by 1", dict(hull_variant=2)), ("hull rows rotated by 3", dict(hull_variant=6)),
                ("hull rows normalised (unit normals)", dict(hull_variant=1)),
                ("hull rows normalised + rotated by 2", dict(hull_variant=5)),
                ("adaptive_rho_interval = 25", dict(rho_interval=25)), ("adaptive_rho_interval = 50", dict(rho_interval=50)),
                ("adaptive_rho_interval = 200", dict(rho_interval=200)),
                ("KKT back-end: reduced Cholesky instead of sparse LDL'", dict(linsys=1)),
                ("eps_abs = eps_rel = 1e-6 (what a converged solve gives)", dict(eps=1e-6))]
    for name, kw in variants:
        Z, S, I = run(B, T, N, 0.002, 0.06, **kw)
        rel = np.abs(Z - base[0]) / np.maximum(1.0, np.abs(base[0]))
        hist = dict(zip(*[a.tolist() for a in np.unique(I, return_counts=True)]))
        out.append(f"| {name} | {rel.max():.2e} | {np.quantile(rel, 0.99):.2e} | {np.median(rel):.2e} | "
                   f"{int((S != base[1]).sum())} | {int((I != base[2]).sum())} | {hist} |")
    hist0 = dict(zip(*[a.tolist() for a in np.unique(base[2], return_counts=True)]))
    out += ["", f"Baseline: status histogram {dict(zip(*[a.tolist() for a in np.unique(base[1], return_counts=True)]))}, "
            f"iteration histogram {hist0}.", "",
            "Reading.  (1) The ORDER of the hull rows does not matter: ADMM is invariant under a permutation of the constraint rows,",
            "and the restatement reproduces that to rounding (1e-13).  (2) Their SCALING does: iDynTree may hand OSQP unit-normal",
            "rows where the restatement passes un-normalised edge normals; the QP is the same, but Ruiz equilibration and the",
            "rho-weighted residuals see different rows, the iterates move, and OSQP -- which stops at eps = 1e-3 -- returns a u0",
            "that differs by up to ~1e-2 relative (p99 3e-3): exactly the distance of the baseline to the converged solution (last",
            "row), i.e. OSQP's own termination tolerance.  (3) The iteration at which rho is first adapted (wall-clock derived in",
            "real OSQP) changes one instance-tick in 61 440 here; (4) the KKT back-end only changes rounding (1e-11).",
            "So: GPU-vs-oracle parity (1e-6, identical status and iteration counts) certifies that the kernels implement THIS",
            "restatement exactly; against a real osqp-eigen + iDynTree build the expected agreement of u0 is OSQP's tolerance",
            "(1e-3 .. 1e-2 on a few per cent of the ticks, 1e-7 in the median), and iteration counts need not coincide instance by",
            "instance (378 of 61 440 differ under row normalisation).  Status codes did not change under any variation."]
    path = os.path.join(ROOT, "profiles", "r2_sensitivity.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
