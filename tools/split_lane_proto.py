"""Lane-accurate CPU prototype of mpc_split.cu (four lanes per instance, iterate in registers, unscaled variables).
Design / test infrastructure only: nothing in the product imports it.  Run:  python tools/split_lane_proto.py [N]

OSQP iterates on the scaled problem  Pb = c D P D, qb = c D q, Ab = E A D.  Writing X = D xb, Yh = E yb the
reduced KKT step  (Pb + sigma I + Ab' rho Ab) xt = sigma xb - qb + Ab'(rho zb - yb)  becomes, after a left
multiplication by D^-1,
    (c P + sigma D^-2 + A' (rho E^2) A) Xt = sigma D^-2 X - c q + A' E (rho zb - yb)
in which A has the constant coefficients (-1, a, b) of the LIPM recursion: no per-row coefficient arrays.

Per axis the unknowns x_N, u_{N-1}, ..., x_1, u_0, x_0 form a pentadiagonal chain.  Two lanes share it:
  top lane    slots t = 0..T-1  <->  (x_k, u_{k-1}), row k,   k = N - t;   slot T-1 is the separator (x_s, u_{s-1}), s = T_bot
  bottom lane slots t = off..T-1 <-> (x_j, u_j),     row j+1, j = t - off; eliminated bottom-up from x_0 with u_0 as a
              border column; its last slot holds x_{s-1} and a shadow of the separator's u_{s-1} / row s.
What remains is a dense 6x6 on (x_s, u_{s-1}, u_0) of both axes (the hull rows couple u_0x and u_0y).
The functions below run exactly the per-lane recurrences of the kernel and are checked against dense linear algebra
on the scaled problem.
"""
import sys

import numpy as np


class Problem:
    """one instance: LIPM constants, weights, scaling, hull rows (random but well-formed)"""

    def __init__(self, N, rng, nh=6):
        self.N = N
        self.a, self.b = 1.04396, -0.04396
        self.Q = np.array([7500.0, 7500.0]) * rng.uniform(0.5, 2, 2)
        self.R = np.array([9e6, 9e6]) * rng.uniform(0.5, 2, 2)
        self.sigma, self.alpha, self.rho = 1e-6, 1.6, 0.1
        self.c = rng.uniform(1e-4, 1e-3)
        self.Dx = rng.uniform(0.05, 0.5, (2, N + 1))      # D of x_k
        self.Du = rng.uniform(0.002, 0.02, (2, N))        # D of u_k
        self.E = rng.uniform(0.5, 3.0, (2, N + 1))        # E of equality row k
        self.nh = nh
        ang = np.sort(rng.uniform(0, 2 * np.pi, nh))
        self.H = np.stack([np.cos(ang), np.sin(ang)], 1)  # unscaled hull rows
        self.hb = rng.uniform(0.03, 0.08, nh)
        self.Eh = rng.uniform(0.5, 3.0, nh)
        self.q = rng.standard_normal((2, N + 1)) * 50     # unscaled gradient of the states
        self.qu0 = rng.standard_normal(2) * 1e4           # unscaled gradient of u_0 (-R u_prev)
        self.dcm = rng.standard_normal(2) * 0.05
        self.infty = 1e30

    # ---- dense scaled problem, variable order per axis a: x_0..x_N, u_0..u_{N-1}; axes stacked
    def dense(self):
        N = self.N
        nA = 2 * N + 1
        n = 2 * nA
        ix = lambda a, k: a * nA + k
        iu = lambda a, k: a * nA + N + 1 + k
        P = np.zeros((n, n)); q = np.zeros(n); D = np.zeros(n)
        for a in range(2):
            for k in range(N + 1):
                P[ix(a, k), ix(a, k)] = self.Q[a]; q[ix(a, k)] = self.q[a, k]; D[ix(a, k)] = self.Dx[a, k]
            for k in range(N):
                sc = 2.0 if k < N - 1 else 1.0
                P[iu(a, k), iu(a, k)] = sc * self.R[a]; D[iu(a, k)] = self.Du[a, k]
                if k >= 1:
                    P[iu(a, k), iu(a, k - 1)] = P[iu(a, k - 1), iu(a, k)] = -self.R[a]
            q[iu(a, 0)] = self.qu0[a]
        m = 2 * (N + 1) + self.nh
        A = np.zeros((m, n)); l = np.zeros(m); u = np.zeros(m); E = np.zeros(m)
        for a in range(2):
            r0 = a * (N + 1)
            A[r0, ix(a, 0)] = -1.0; l[r0] = u[r0] = -self.dcm[a]; E[r0] = self.E[a, 0]
            for k in range(1, N + 1):
                r = r0 + k
                A[r, ix(a, k)] = -1.0; A[r, ix(a, k - 1)] = self.a; A[r, iu(a, k - 1)] = self.b; E[r] = self.E[a, k]
        for h in range(self.nh):
            r = 2 * (N + 1) + h
            A[r, iu(0, 0)] = self.H[h, 0]; A[r, iu(1, 0)] = self.H[h, 1]
            l[r] = -self.infty; u[r] = self.hb[h]; E[r] = self.Eh[h]
        rho = np.where(l == u, 1e3 * self.rho, self.rho)
        return dict(P=P, q=q, A=A, l=l, u=u, D=D, E=E, rho=rho, ix=ix, iu=iu, m=m, n=n)


def dense_admm(pb, iters, x0, y0, z0):
    """OSQP's iteration on the scaled problem (dense linear algebra); returns the scaled iterate history"""
    d = pb.dense()
    D, E, c = d["D"], d["E"], pb.c
    Pb = c * D[:, None] * d["P"] * D[None, :]; qb = c * D * d["q"]; Ab = E[:, None] * d["A"] * D[None, :]
    lb, ub = E * d["l"], E * d["u"]
    rho = d["rho"]
    K = Pb + pb.sigma * np.eye(d["n"]) + Ab.T @ (rho[:, None] * Ab)
    x, y, z = x0.copy(), y0.copy(), z0.copy()
    hist = []
    for _ in range(iters):
        xt = np.linalg.solve(K, pb.sigma * x - qb + Ab.T @ (rho * z - y))
        zt = Ab @ xt
        xn = pb.alpha * xt + (1 - pb.alpha) * x
        zr = pb.alpha * zt + (1 - pb.alpha) * z
        zn = np.clip(zr + y / rho, lb, ub)
        y = y + rho * (zr - zn)
        x, z = xn, zn
        hist.append((x.copy(), y.copy(), z.copy()))
    return d, hist


# =====================================================================================================
# the lane program
# =====================================================================================================
class Lane:
    pass


def make_lanes(pb):
    """per-lane constant arrays (what stage-in builds in shared memory) for the four lanes of one instance"""
    N = pb.N
    Ttop = (N + 2) // 2
    Tbot = N + 1 - Ttop
    T = Ttop
    off = T - Tbot
    lanes = {}
    for a in range(2):
        for part in range(2):
            L = Lane(); L.a, L.part, L.T, L.off = a, part, T, off
            L.SX = np.zeros(T); L.SU = np.zeros(T); L.CQ = np.zeros(T); L.EE = np.zeros(T)
            L.stage = [None] * T
            for t in range(T):
                if part == 0:
                    k = N - t
                    L.stage[t] = (k, k - 1, k)             # (x index, u index, row index)
                elif t >= off:
                    j = t - off
                    L.stage[t] = (j, j, j + 1)
                if L.stage[t] is None:
                    continue
                kx, ku, kr = L.stage[t]
                L.SX[t] = pb.sigma / pb.Dx[a, kx] ** 2
                L.SU[t] = pb.sigma / pb.Du[a, ku] ** 2
                L.CQ[t] = pb.c * pb.q[a, kx]
                L.EE[t] = pb.E[a, kr] ** 2
            L.cy0, L.cy1 = (1.0, -pb.a) if part == 0 else (-pb.a, 1.0)
            L.EE0 = pb.E[a, 0] ** 2                        # bottom only: row 0
            lanes[(a, part)] = L
    return lanes, T, off


def factor_lane(pb, L, rho, hull_diag):
    """column LDL' of the lane's unknowns in elimination order with the border (bottom: u_0) carried;
    returns the Schur terms on (near, far, border)"""
    T, off, a_, b_ = L.T, L.off, pb.a, pb.b
    rho_e = 1e3 * rho
    cQ, cR = pb.c * pb.Q[L.a], pb.c * pb.R[L.a]
    top = L.part == 0
    z = np.zeros(T)
    L.l1x, L.l2x, L.dix, L.lbx, L.l1u, L.l2u, L.diu, L.lbu = (z.copy() for _ in range(8))
    # carried: multipliers and (multiplier x pivot) of the previous two unknowns
    l1p = m1p = l2p = m2p = lbp = 0.0          # previous unknown e-1: l1, m1 (= l1 d), l2, m2, lb
    l2pp = m2pp = lbpp = 0.0                   # unknown e-2: l2, m2, lb
    Sbb = 0.0
    Wprev = 0.0 if top else rho_e * L.EE0      # W of the row "before" the first real slot
    n_elim = T - 1                             # slots fully eliminated by both lanes
    out = {}
    for t in range(T):
        W = rho_e * L.EE[t]
        real = L.stage[t] is not None
        if real:
            kx, ku, kr = L.stage[t]
            sc = 2.0 if ku < pb.N - 1 else 1.0
            g_x = cQ + L.SX[t] + ((W + a_ * a_ * Wprev) if top else (Wprev + a_ * a_ * W))
            g_u = cR * sc + L.SU[t] + b_ * b_ * W
            s1x, s2x = ((-b_ * W, -a_ * W) if top else (a_ * b_ * W, -a_ * W))
            s1u, s2u = ((a_ * b_ * W) if top else (-b_ * W)), -cR
            bdx = bdu = 0.0
            u_dummy = False
            if not top and t == off:            # (x_0, border u_0): the u part is a dummy of the sequence
                s1x = 0.0; bdx = a_ * b_ * W; u_dummy = True
                g_u0 = g_u + hull_diag
            if not top and t == off + 1:
                bdx = -b_ * rho_e * L.EE[off]; bdu = -cR
        else:
            g_x = g_u = 1.0; s1x = s2x = s1u = s2u = bdx = bdu = 0.0; u_dummy = True
        if t == T - 1:                          # separator slot
            if top:
                out.update(g_xs=g_x, g_us=g_u, k_xu=s1x,
                           Snn=l1p * m1p + l2pp * m2pp, Sfn=l2p * m1p, Sff=l2p * m2p)
                break
            # bottom: eliminate x_{s-1} only
        # ---- x part
        d = g_x - l1p * m1p - l2pp * m2pp
        m1 = s1x - l2p * m1p
        m2 = s2x
        mb = bdx - lbp * m1p - lbpp * m2pp
        di = 1.0 / d if real else 0.0
        l1, l2, lb = m1 / d, m2 / d, mb / d
        L.l1x[t], L.l2x[t], L.dix[t], L.lbx[t] = l1, l2, di, lb
        Sbb += lb * mb
        l2pp, m2pp, lbpp = l2p, m2p, lbp
        l1p, m1p, l2p, m2p, lbp = l1, m1, l2, m2, lb
        if t == T - 1:                          # bottom's last: Schur terms, near = u_sep, far = x_sep
            out.update(Snn=l1p * m1p + l2pp * m2pp, Sfn=l2p * m1p, Sff=l2p * m2p,
                       Sbn=lbp * m1p + lbpp * m2pp, Sbf=lbp * m2p, Sbb=Sbb, g_u0=g_u0,
                       k_0u=(-cR if (T - 1 - off) == 1 else 0.0))
            break
        # ---- u part
        if u_dummy:
            d = 1.0; m1 = m2 = mb = 0.0; di = 0.0
            d = 1.0 - 0.0 * m1p                 # a dummy never couples
            l1 = l2 = lb = 0.0
            # the previous unknown's l1 pointed at this dummy: it is zero by construction (s1x = 0 above)
            assert abs(l1p) == 0.0
        else:
            d = g_u - l1p * m1p - l2pp * m2pp
            m1 = s1u - l2p * m1p
            m2 = s2u
            mb = bdu - lbp * m1p - lbpp * m2pp
            di = 1.0 / d
            l1, l2, lb = m1 / d, m2 / d, mb / d
        L.l1u[t], L.l2u[t], L.diu[t], L.lbu[t] = l1, l2, di, lb
        Sbb += lb * mb
        l2pp, m2pp, lbpp = l2p, m2p, lbp
        l1p, m1p, l2p, m2p, lbp = l1, m1, l2, m2, lb
        if real:
            Wprev = W
    return out


def build_factor(pb, lanes, rho):
    Hs = pb.Eh[:, None] * pb.H                               # half-scaled hull rows E_h * h
    hd = rho * (Hs ** 2).sum(0)
    kxy = rho * (Hs[:, 0] * Hs[:, 1]).sum()
    M = np.zeros((6, 6))
    for a in range(2):
        tp = factor_lane(pb, lanes[(a, 0)], rho, 0.0)
        bt = factor_lane(pb, lanes[(a, 1)], rho, hd[a])
        o = 3 * a
        M[o + 0, o + 0] = tp["g_xs"] - tp["Snn"] - bt["Sff"]
        M[o + 1, o + 1] = tp["g_us"] - tp["Sff"] - bt["Snn"]
        M[o + 0, o + 1] = M[o + 1, o + 0] = tp["k_xu"] - tp["Sfn"] - bt["Sfn"]
        M[o + 2, o + 0] = M[o + 0, o + 2] = -bt["Sbf"]
        M[o + 2, o + 1] = M[o + 1, o + 2] = bt["k_0u"] - bt["Sbn"]
        M[o + 2, o + 2] = bt["g_u0"] - bt["Sbb"]
    M[2, 5] = M[5, 2] = kxy
    return np.linalg.inv(M), Hs


def lane_iteration(pb, lanes, st, Minv, Hs, rho):
    """one ADMM iteration; st holds the 'registers': per lane X, U, Y (T each), bottom lanes U0-related scalars,
    hull z / y (scaled, as OSQP keeps them)"""
    T = lanes[(0, 0)].T; off = lanes[(0, 0)].off
    a_, b_, al, sg = pb.a, pb.b, pb.alpha, pb.sigma
    rho_e = 1e3 * rho
    fw = {}
    # hull contribution to the right-hand side of u_0 (carried from the previous iteration in the kernel)
    v = rho * st["zh"] - st["yh"]
    hull_r = Hs.T @ v
    for key, L in lanes.items():
        R = st[key]
        top = L.part == 0
        p1 = p2 = 0.0
        rb = 0.0
        tX = np.zeros(T); tU = np.zeros(T)
        Yprev = 0.0 if top else R["Yext"]
        for t in range(T):
            yprev = Yprev if t == (0 if top else off) else (R["Y"][t - 1] if t > 0 else 0.0)
            rx = L.SX[t] * R["X"][t] - L.CQ[t] + L.cy0 * R["Y"][t] + L.cy1 * yprev
            ru = L.SU[t] * R["U"][t] - b_ * R["Y"][t]
            if not top and t == off:
                rb = ru - pb.c * pb.qu0[L.a] + hull_r[L.a]
                ru = 0.0
            if t == T - 1:
                if top:
                    vx, vu = rx - p1, ru - p2
                    break
                tx = rx - p1
                tX[t] = tx
                p1, p2 = p2 + L.l1x[t] * tx, L.l2x[t] * tx
                rb -= L.lbx[t] * tx
                vx, vu = -p2, -p1
                break
            tx = rx - p1
            tX[t] = tx
            p1, p2 = p2 + L.l1x[t] * tx, L.l2x[t] * tx
            tu = ru - p1
            tU[t] = tu
            p1, p2 = p2 + L.l1u[t] * tu, L.l2u[t] * tu
            rb -= L.lbx[t] * tx + L.lbu[t] * tu
        fw[key] = (tX, tU, vx, vu, rb)
    r6 = np.zeros(6)
    for a in range(2):
        r6[3 * a + 0] = fw[(a, 0)][2] + fw[(a, 1)][2]
        r6[3 * a + 1] = fw[(a, 0)][3] + fw[(a, 1)][3]
        r6[3 * a + 2] = fw[(a, 1)][4]
    s6 = Minv @ r6
    # hull rows (scaled z, y; half-scaled rows on the unscaled u_0)
    zt = Hs @ np.array([s6[2], s6[5]])
    zr = al * zt + (1 - al) * st["zh"]
    zn = np.clip(zr + st["yh"] / rho, pb.Eh * (-pb.infty), pb.Eh * pb.hb)
    st["yh"] = st["yh"] + rho * (zr - zn)
    st["zh"] = zn
    xbot_last = {}
    for part in (1, 0):                                       # bottom first: the top's separator row needs x_{s-1}
        for a in range(2):
            L = lanes[(a, part)]; R = st[(a, part)]
            top = part == 0
            tX, tU = fw[(a, part)][0], fw[(a, part)][1]
            Xs, Us, U0 = s6[3 * a], s6[3 * a + 1], s6[3 * a + 2]
            ub = 0.0 if top else U0
            cx0, cx1 = -L.cy0, -L.cy1
            sx = su = 0.0
            for t in range(T - 1, -1, -1):
                if t == T - 1:
                    xu = Us
                    if top:
                        xx = Xs
                        sx_row = xbot_last[a]
                    else:
                        xx = tX[t] * L.dix[t] - L.l1x[t] * xu - L.l2x[t] * Xs - L.lbx[t] * ub
                        xbot_last[a] = xx
                        sx_row = Xs
                else:
                    xu = tU[t] * L.diu[t] - L.l1u[t] * sx - L.l2u[t] * su - L.lbu[t] * ub
                    if not top and t == off:
                        xu = U0
                    xx = tX[t] * L.dix[t] - L.l1x[t] * xu - L.l2x[t] * sx - L.lbx[t] * ub
                    sx_row = sx
                row = b_ * xu + (cx0 * xx + cx1 * sx_row)
                R["Y"][t] += (al * rho_e) * (L.EE[t] * row)
                R["X"][t] = al * xx + (1 - al) * R["X"][t]
                R["U"][t] = al * xu + (1 - al) * R["U"][t]
                sx, su = xx, xu
                if not top and t == off:                      # row 0: l = u = -dcm
                    W0 = rho_e * L.EE0
                    R["Y0"] += W0 * (-al * xx + (1 - al) * R["Z0"] + pb.dcm[a])
                    R["Z0"] = -pb.dcm[a]
                    R["Yext"] = R["Y0"] - W0 * R["Z0"]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    rng = np.random.default_rng(N)
    worst = 0.0
    for trial in range(4):
        pb = Problem(N, rng, nh=[4, 6, 8, 5][trial])
        d0 = pb.dense()
        n, m = d0["n"], d0["m"]
        x0 = rng.standard_normal(n) * 0.1; y0 = rng.standard_normal(m) * 0.1
        z0 = np.zeros(m)
        nR = N + 1
        z0[0] = rng.standard_normal() * 0.01; z0[nR] = rng.standard_normal() * 0.01     # warm z of row 0 (both axes)
        z0[2 * nR:] = np.minimum(rng.standard_normal(pb.nh) * 0.05, d0["E"][2 * nR:] * d0["u"][2 * nR:])
        iters = 30
        d, hist = dense_admm(pb, iters, x0, y0, z0)
        lanes, T, off = make_lanes(pb)
        Minv, Hs = build_factor(pb, lanes, pb.rho)
        D, E = d["D"], d["E"]
        Xun, Yh = D * x0, E * y0
        st = {"zh": z0[2 * nR:].copy(), "yh": y0[2 * nR:].copy()}
        for (a, part), L in lanes.items():
            R = {"X": np.zeros(T), "U": np.zeros(T), "Y": np.zeros(T)}
            for t in range(T):
                if L.stage[t] is None:
                    continue
                kx, ku, kr = L.stage[t]
                R["X"][t] = Xun[d["ix"](a, kx)]; R["U"][t] = Xun[d["iu"](a, ku)]; R["Y"][t] = Yh[a * nR + kr]
            if part == 1:
                R["Y0"] = Yh[a * nR]; R["Z0"] = z0[a * nR] / pb.E[a, 0]
                R["Yext"] = R["Y0"] - 1e3 * pb.rho * L.EE0 * R["Z0"]
            st[(a, part)] = R
        for it in range(iters):
            lane_iteration(pb, lanes, st, Minv, Hs, pb.rho)
            xs, ys, zs = hist[it]
            Xref, Yref = D * xs, E * ys
            err = 0.0
            for (a, part), L in lanes.items():
                R = st[(a, part)]
                for t in range(T):
                    if L.stage[t] is None:
                        continue
                    kx, ku, kr = L.stage[t]
                    err = max(err, abs(R["X"][t] - Xref[d["ix"](a, kx)]) / (1e-3 + abs(Xref[d["ix"](a, kx)])))
                    err = max(err, abs(R["U"][t] - Xref[d["iu"](a, ku)]) / (1e-3 + abs(Xref[d["iu"](a, ku)])))
                    err = max(err, abs(R["Y"][t] - Yref[a * nR + kr]) / (1e-3 + abs(Yref[a * nR + kr])))
                if part == 1:
                    err = max(err, abs(R["Y0"] - Yref[a * nR]) / (1e-3 + abs(Yref[a * nR])))
            err = max(err, np.abs(st["yh"] - ys[2 * nR:]).max(), np.abs(st["zh"] - zs[2 * nR:]).max())
            worst = max(worst, err)
            assert err < 1e-7, (trial, it, err)
    print(f"N = {N}: T = {T} slots per lane (bottom offset {off}); lane program vs dense scaled ADMM over 30 iterations: "
          f"worst relative deviation {worst:.2e}")
    print("ok")


if __name__ == "__main__":
    main()
