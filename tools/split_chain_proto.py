"""Lane-accurate CPU prototype of the FOUR-LANES-PER-INSTANCE reduced-KKT solve planned for the MPC phase-1
kernel ("plan for the next round").  Test/design infrastructure only: nothing in the
product imports it.  Run:  python tools/split_chain_proto.py [horizon]

Today (mpc_chain.cu) one thread owns one axis chain of the reduced KKT matrix
    K = P + sigma I + A' diag(rho) A                (OSQP's KKT system after eliminating the nu block),
ordered  x_N, u_{N-1}, x_{N-1}, ..., u_1, x_1 | u_0, x_0 : pentadiagonal per axis plus a dense 4x4 corner
on (u_0, x_0) of both axes (the hull rows couple u_0x and u_0y).  One warp = 16 instances; its dependent
sweep over 2N+1 unknowns is what bounds the kernel (one lone warp per SM sub-partition, 104.7 KB each).

Here every axis chain is cut at a two-unknown separator S = (c[s], c[s+1]) in its middle and given to TWO lanes:
  * upper lane: unknowns c[0..s-1], eliminated top-down exactly as now; its Schur complement lands on S;
  * lower lane: unknowns c[s+2..] without u_0, eliminated BOTTOM-UP starting at x_0, with u_0 carried as a
    border column (one extra multiplier per unknown: eliminating x_0 couples u_0 to x_1, eliminating x_1
    couples it to u_1 and x_2, ...); its Schur complement lands on S and on u_0;
  * what is left is a dense 6x6 system on (S_x, S_y, u_0x, u_0y) per instance, solved redundantly by the
    four lanes after one exchange (shuffles), then both lanes substitute back away from the separator.
Each lane's dependent chain is half as long, four lanes x 8 instances fill a warp with ~52 KB of shared
memory, so four single-warp CTAs fit an SM and every sub-partition has a warp.

The prototype factorises and solves with exactly the recurrences a lane would run (scalars and the lane's
own arrays only; the 6x6 is the one shared step) and checks the result against a dense solve of K on QPs taken
from the oracle's MPC.  It also reports the per-lane storage and the flop / load counts per ADMM iteration
that DESIGN.md quotes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def chain_order(N, axis):
    """indices into z = [x_0 .. x_N (2 each), u_0 .. u_{N-1} (2 each)] of one axis chain,
    order x_N, u_{N-1}, x_{N-1}, ..., u_1, x_1, u_0, x_0 (mpc_chain.cu)"""
    xs = lambda k: 2 * k + axis
    us = lambda k: 2 * (N + 1) + 2 * k + axis
    c = [xs(N)]
    for k in range(N - 1, 0, -1):
        c += [us(k), xs(k)]
    c += [us(0), xs(0)]
    return c


class LaneFactor:
    """what one lane keeps in shared memory after the factorisation"""
    pass


def factor_upper(T, s):
    """LDL' of the leading s unknowns of the pentadiagonal T (top-down).  Returns the lane's arrays
    (l1, l2, dinv: one entry per unknown) and the 2x2 Schur complement it contributes to the separator
    rows (s, s+1) together with the two coupling multipliers rows that the right-hand side update needs."""
    l1 = np.zeros(s); l2 = np.zeros(s); d = np.zeros(s)
    # working copy of the three diagonals; the recurrences touch (j, j+1, j+2) only
    dg = np.array([T[j, j] for j in range(s + 2)])
    e1 = np.array([T[j + 1, j] for j in range(s + 1)])           # sub-diagonal, entry (j+1, j)
    e2 = np.array([T[j + 2, j] for j in range(s)])               # second sub-diagonal, entry (j+2, j)
    for j in range(s):
        d[j] = dg[j]
        a1 = e1[j] / d[j]                                        # multiplier of row j+1
        a2 = e2[j] / d[j]                                        # multiplier of row j+2
        l1[j], l2[j] = a1, a2
        dg[j + 1] -= a1 * e1[j]
        e1[j + 1] -= a2 * e1[j] if j + 1 < s + 1 else 0.0        # entry (j+2, j+1) -= l2_j * d_j * l1_j
        dg[j + 2] -= a2 * e2[j]
    # the separator block after the upper elimination: [[dg[s], e1[s]], [e1[s], dg[s+1]]]
    return dict(l1=l1, l2=l2, dinv=1.0 / d), (dg[s], e1[s], dg[s + 1])


def factor_lower(T, s, n):
    """Bottom-up LDL' of unknowns n-1 (x_0), then n-3, n-4, ..., s+2 with u_0 = unknown n-2 as a border.
    Lane arrays: l1, l2 (multipliers towards the two unknowns above), lb (multiplier towards u_0), dinv.
    Returns them plus the Schur contributions on (s, s+1, u_0)."""
    b = n - 2
    order = [n - 1] + list(range(n - 3, s + 1, -1))              # elimination order
    m = len(order)
    # dense working copy restricted to the lane's unknowns + separator + border (the prototype keeps it simple;
    # the checks below assert the band + border sparsity the kernel relies on)
    W = T.copy()
    l1 = np.zeros(m); l2 = np.zeros(m); lb = np.zeros(m); d = np.zeros(m)
    up = lambda j: [i for i in ([j - 1, j - 2] if j != n - 1 else [n - 3]) if i >= s]   # band neighbours above
    for t, j in enumerate(order):
        d[t] = W[j, j]
        nb = up(j)
        live = [i for i in range(n) if i != j and W[i, j] != 0.0 and i not in order[:t]]
        assert set(live) <= set(nb) | {b}, (j, live, nb)          # fill stays inside band + border
        mult = {i: W[i, j] / d[t] for i in nb + [b]}
        if j == n - 1:
            l1[t] = mult[n - 3]; l2[t] = 0.0
        else:
            l1[t] = mult.get(j - 1, 0.0); l2[t] = mult.get(j - 2, 0.0)
        lb[t] = mult[b]
        rows = nb + [b]
        for i in rows:
            for k in rows:
                W[i, k] -= mult[i] * W[j, k]
        W[j, :] = 0.0; W[:, j] = 0.0
    keep = [s, s + 1, b]
    return dict(order=order, l1=l1, l2=l2, lb=lb, dinv=1.0 / d), W[np.ix_(keep, keep)]


def solve_split(Tx, Ty, kxy, rx, ry, s):
    """Solve [[Tx, C], [C', Ty]] z = [rx; ry] with C = kxy e_b e_b' (b = u_0), the way four lanes would."""
    n = Tx.shape[0]
    b = n - 2
    lanes = []
    S6 = np.zeros((6, 6)); r6 = np.zeros(6)
    for a, (T, r) in enumerate(((Tx, rx), (Ty, ry))):
        fu, (sd0, se, sd1) = factor_upper(T, s)
        fl, Wl = factor_lower(T, s, n)
        # ---- forward sweeps (fused with the right-hand-side build in the kernel)
        tu = np.zeros(s)
        c1 = c2 = 0.0                                            # carried into rows j+1, j+2
        rs = np.array([r[s], r[s + 1]])
        acc = np.zeros(s + 2)
        for j in range(s):                                       # upper lane: t_j = r_j - l1_{j-1} t_{j-1} - l2_{j-2} t_{j-2}
            tj = r[j] - acc[j]
            tu[j] = tj
            acc[j + 1] += fu["l1"][j] * tj
            acc[j + 2] += fu["l2"][j] * tj
        rs -= acc[s:s + 2]
        order = fl["order"]
        tl = np.zeros(len(order))
        accl = {i: 0.0 for i in range(s, n)}
        rb = r[b]
        for t, j in enumerate(order):                            # lower lane, bottom-up, border accumulated in a register
            tj = r[j] - accl[j]
            tl[t] = tj
            if j == n - 1:
                accl[n - 3] += fl["l1"][t] * tj
            else:
                accl[j - 1] += fl["l1"][t] * tj
                accl[j - 2] += fl["l2"][t] * tj
            rb -= fl["lb"][t] * tj
        rs -= np.array([accl[s], accl[s + 1]])
        # ---- the lane pair's rows of the 6x6: separator block = T_SS minus both Schur complements
        Sup = np.array([[sd0, se], [se, sd1]])                   # T_SS - upper Schur
        Tss = T[np.ix_([s, s + 1], [s, s + 1])]
        blk = np.zeros((3, 3))
        blk[:2, :2] = Sup + (Wl[:2, :2] - Tss)                   # both corrections applied once
        blk[:2, 2] = Wl[:2, 2]; blk[2, :2] = Wl[2, :2]; blk[2, 2] = Wl[2, 2]
        idx = [2 * a, 2 * a + 1, 4 + a]
        S6[np.ix_(idx, idx)] = blk
        r6[idx] = [rs[0], rs[1], rb]
        lanes.append((fu, fl, tu, tl))
    S6[4, 5] = S6[5, 4] = kxy
    z6 = np.linalg.solve(S6, r6)                                 # dense 6x6 (LDL' held in registers in the kernel)
    out = []
    for a, T in enumerate((Tx, Ty)):
        fu, fl, tu, tl = lanes[a]
        z = np.zeros(n)
        z[s], z[s + 1], z[b] = z6[2 * a], z6[2 * a + 1], z6[4 + a]
        for j in range(s - 1, -1, -1):                           # upper lane: away from the separator
            z[j] = tu[j] * fu["dinv"][j] - fu["l1"][j] * z[j + 1] - fu["l2"][j] * z[j + 2]
        order = fl["order"]
        for t in range(len(order) - 1, -1, -1):                  # lower lane: away from the separator, border term added
            j = order[t]
            v = tl[t] * fl["dinv"][t] - fl["lb"][t] * z[b]
            if j == n - 1:
                v -= fl["l1"][t] * z[n - 3]
            else:
                v -= fl["l1"][t] * z[j - 1] + fl["l2"][t] * z[j - 2]
            z[j] = v
        out.append(z)
    return out


def reduced_kkt_from_oracle(N, seed=3, tick=40):
    """K and a right-hand side of one instance of the oracle's MPC (unscaled data, OSQP's rho vector)."""
    import importlib
    import torch  # noqa: F401  (synthetic inputs)
    from oracle import oracle as O
    syn = importlib.import_module("walking-controllers_b200.synthetic")
    B = 4
    g = syn.GaitBatch(B, tick + 2, N, seed=seed)
    traj = g.trajectory()
    mo = O.MpcBatch(O.mpc_default_config(N), B)
    for t in range(tick):
        f = g.feet(t)
        mo.tick(f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(),
                g.measured_dcm(t).numpy(), traj[2 * t: 2 * t + 2 * (N + 1)].contiguous().numpy())
    out = []
    for i in range(B):
        P, A, q, l, u = mo.qp_dense(i)
        rho = np.where(np.abs(l - u) < 1e-4, 1e3 * 0.1, 0.1)
        K = P + 1e-6 * np.eye(P.shape[0]) + A.T @ (rho[:, None] * A)
        out.append(K)
    return out


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    rng = np.random.default_rng(0)
    worst = 0.0
    for K in reduced_kkt_from_oracle(N):
        cx, cy = chain_order(N, 0), chain_order(N, 1)
        n = len(cx)
        Tx, Ty = K[np.ix_(cx, cx)], K[np.ix_(cy, cy)]
        Cxy = K[np.ix_(cx, cy)]
        b = n - 2
        kxy = Cxy[b, b]
        Cz = Cxy.copy(); Cz[b, b] = 0.0
        assert np.abs(Cz).max() == 0.0, "axes couple through u_0 only"
        for T in (Tx, Ty):
            assert np.abs(np.triu(T, 3)).max() == 0.0, "pentadiagonal in chain order"
        s = (n - 2) // 2 - 1                                     # separator in the middle of the 2N-1 chain unknowns
        r = rng.standard_normal(2 * n)
        zx, zy = solve_split(Tx, Ty, kxy, r[:n], r[n:], s)
        Kc = np.block([[Tx, Cxy], [Cxy.T, Ty]])
        zref = np.linalg.solve(Kc, r)
        err = np.abs(np.concatenate([zx, zy]) - zref).max() / np.abs(zref).max()
        worst = max(worst, err)
    n_up, n_lo = s, n - 3 - (s + 1)
    print(f"N = {N}: chain of {n} unknowns per axis, separator at {s},{s + 1}: upper lane {n_up}, lower lane {n_lo + 1} unknowns")
    print(f"  split solve vs dense solve, worst relative error over instances: {worst:.2e}")
    print(f"  factor storage per axis: now {3 * n} doubles; split {3 * n_up + 4 * (n_lo + 1)} + 6x6 LDL' (21 + 6) per instance")
    assert worst < 1e-9
    print("ok")


if __name__ == "__main__":
    main()
