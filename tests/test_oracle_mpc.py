"""The MPC restatement (oracle/mpc_ref.c) checked by independent means."""
import numpy as np
import pytest
import scipy.linalg as sla


def rect(cx, cy, yaw, front=0.05, back=0.02, left=0.025, right=0.025):
    c, s = np.cos(yaw), np.sin(yaw)
    pts = np.array([[front, left], [-back, left], [-back, -right], [front, -right]])
    return pts @ np.array([[c, s], [-s, c]]) + np.array([cx, cy])


def test_hull_single_foot_is_the_rectangle(oracle):
    P = rect(0.3, -0.1, 0.4)
    A, b, v = oracle.hull_build(P)
    assert A.shape == (4, 2)
    for p in P:                                   # vertices lie on the boundary
        assert np.abs(np.max(A @ p - b)) < 1e-12
    assert (A @ np.array([0.3, -0.1]) < b).all()  # centre strictly inside
    # counter-clockwise ordering
    area = 0.5 * sum(v[i][0] * v[(i + 1) % 4][1] - v[(i + 1) % 4][0] * v[i][1] for i in range(4))
    assert abs(area - 0.07 * 0.05) < 1e-12


def test_hull_double_support_contains_both_feet(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        yaw = rng.uniform(-3, 3)
        L = rect(0, 0.08, yaw + rng.uniform(-0.2, 0.2)); Rr = rect(rng.uniform(0.05, 0.2), -0.08, yaw)
        A, b, v = oracle.hull_build(np.vstack([L, Rr]))
        assert 4 <= A.shape[0] <= 8
        for p in np.vstack([L, Rr]):
            assert (A @ p <= b + 1e-12).all()
        # every hull vertex is an input point
        for q in v:
            assert np.min(np.linalg.norm(np.vstack([L, Rr]) - q, axis=1)) < 1e-12


def test_hull_collinear_points_are_dropped(oracle):
    L = rect(0.0, 0.08, 0.0); Rr = rect(0.0, -0.08, 0.0)   # aligned feet: hull is a rectangle
    A, b, v = oracle.hull_build(np.vstack([L, Rr]))
    assert A.shape[0] == 4


def test_margin_sign_and_value(oracle):
    A, b, v = oracle.hull_build(rect(0, 0, 0))
    assert abs(oracle.hull_margin(A, b, v, [0.0, 0.0]) - 0.02) < 1e-12       # nearest edge: back (x=-0.02)
    assert abs(oracle.hull_margin(A, b, v, [0.10, 0.0]) + 0.05) < 1e-12      # 5 cm in front of the toe
    assert abs(oracle.hull_margin(A, b, v, [0.08, 0.065]) + 0.05) < 1e-12    # corner distance 3-4-5


def _tick_one(oracle, syn, N, T, seed=3, settings=None):
    g = syn.GaitBatch(1, T, N, seed=seed, random_phase=True)
    traj = g.trajectory().numpy()
    cfg = oracle.mpc_default_config(N)
    if settings:
        for k, v in settings.items():
            setattr(cfg.osqp, k, v)
    mb = oracle.MpcBatch(cfg, 1)
    outs = []
    for t in range(T):
        f = g.feet(t)
        out = mb.tick(f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(),
                      g.measured_dcm(t).numpy(), np.ascontiguousarray(traj[2 * t: 2 * t + 2 * (N + 1)]))
        outs.append(out)
    return mb, g, traj, outs


def test_qp_matrices_follow_the_reference_formulation(oracle, syn):
    """P = blkdiag(Q x (N+1), Theta' R~ Theta), A = [A_eq ; hull on u0]  (Controller.cpp:23-144, MPCSolver.cpp:76-123)"""
    N = 6
    mb, g, traj, outs = _tick_one(oracle, syn, N, 1)
    P, A, q, l, u = mb.qp_dense(0)
    n = 4 * N + 2
    Theta = np.eye(2 * N) - np.eye(2 * N, k=-2)
    Pref = sla.block_diag(7500.0 * np.eye(2 * (N + 1)), Theta.T @ (9e6 * np.eye(2 * N)) @ Theta)
    assert np.allclose(P, Pref)
    w = np.sqrt(9.81 / 0.53); a_, b_ = np.exp(w * 0.01), 1 - np.exp(w * 0.01)
    Aeq = -np.eye(2 * (N + 1), n)
    for i in range(N):
        Aeq[2 * i + 2: 2 * i + 4, 2 * i: 2 * i + 2] += a_ * np.eye(2)
        Aeq[2 * i + 2: 2 * i + 4, 2 * (N + 1) + 2 * i: 2 * (N + 1) + 2 * i + 2] = b_ * np.eye(2)
    assert np.allclose(A[: 2 * (N + 1)], Aeq)
    hA, hb = mb.hull(0)
    assert np.allclose(A[2 * (N + 1):, 2 * (N + 1): 2 * (N + 1) + 2], hA) and np.allclose(u[2 * (N + 1):], hb)
    assert (np.abs(A[2 * (N + 1):, :2 * (N + 1)]) == 0).all() and (A[2 * (N + 1):, 2 * (N + 1) + 2:] == 0).all()
    assert (l[2 * (N + 1):] == -1e30).all() and np.allclose(l[2:2 * (N + 1)], 0) and np.allclose(l[:2], u[:2])
    # gradient: -Q r_k on the states, -R u_prev on u_0 (u_prev = initial_zmp_position = 0), 0 elsewhere
    assert np.allclose(q[: 2 * (N + 1)], -7500.0 * traj[: 2 * (N + 1), 0])
    assert np.allclose(q[2 * (N + 1):], 0.0)


def test_solution_matches_exact_active_set_solution(oracle, syn):
    """tight OSQP tolerances: the oracle's output equals the exact optimum of the assembled QP"""
    from scipy.optimize import minimize
    N = 8
    mb, g, traj, outs = _tick_one(oracle, syn, N, 3, settings=dict(eps_abs=1e-10, eps_rel=1e-10, max_iter=50000))
    P, A, q, l, u = mb.qp_dense(0)
    primal = mb.solution()[0][:, 0]
    me = 2 * (N + 1)
    # exact: equality-constrained KKT if the hull is inactive, otherwise SLSQP polish
    K = np.block([[P, A[:me].T], [A[:me], np.zeros((me, me))]])
    x = np.linalg.solve(K, np.concatenate([-q, l[:me]]))[: 4 * N + 2]
    if (A[me:] @ x <= u[me:] + 1e-9).all():
        assert np.abs(x - primal).max() < 1e-6
    else:
        cons = [{"type": "eq", "fun": lambda z: A[:me] @ z - l[:me]}, {"type": "ineq", "fun": lambda z: u[me:] - A[me:] @ z}]
        res = minimize(lambda z: 0.5 * z @ P @ z + q @ z, primal, jac=lambda z: P @ z + q, constraints=cons, method="SLSQP",
                       options={"ftol": 1e-16, "maxiter": 300})
        assert np.abs(res.x - primal).max() < 1e-5


def test_gradient_shift_and_warm_start_semantics(oracle, syn):
    """warm-started ticks need the minimum of 25 iterations; a phase change builds a cold solver (flag 16);
    the shifted gradient equals a recomputed one when the reference deque advances by one sample"""
    N = 25
    mb, g, traj, outs = _tick_one(oracle, syn, N, 100, seed=5)
    flags = np.array([o[3][0] for o in outs]); iters = np.array([o[2][0] for o in outs])
    st = np.array([o[1][0] for o in outs])
    assert (st == 1).all()
    assert flags[0] == 16 and (flags & 16).sum() >= 2
    assert (iters % 25 == 0).all() and (iters[(flags & 16) == 0] == 25).mean() > 0.9
    P, A, q, l, u = mb.qp_dense(0)
    assert np.allclose(q[: 2 * (N + 1)], -7500.0 * traj[2 * 99: 2 * 99 + 2 * (N + 1), 0])
    zmp = np.array([o[0][:, 0] for o in outs])
    assert np.allclose(q[2 * (N + 1): 2 * (N + 1) + 2], -9e6 * zmp[-2])      # G u_prev with u_prev = previous output


def test_zmp_tracks_reference_and_stays_in_hull(oracle, syn):
    N = 25
    mb, g, traj, outs = _tick_one(oracle, syn, N, 60, seed=9)
    for t, o in enumerate(outs):
        assert (o[3][0] & 8) == 0                       # never outside the hull tolerance
    zmp = np.array([o[0][:, 0] for o in outs])
    ref = np.array([traj[2 * t: 2 * t + 2, 0] for t in range(60)])
    assert np.abs(zmp - ref).max() < 0.25               # ZMP stays within a footstep of the DCM reference


def test_no_contact_and_reset_paths(oracle, syn):
    N = 5
    g = syn.GaitBatch(2, 6, N, seed=2)
    traj = g.trajectory().numpy()
    mb = oracle.MpcBatch(oracle.mpc_default_config(N), 2)
    f = g.feet(0)
    args = lambda t, lc, rc: (f["lf_H"].numpy(), f["rf_H"].numpy(), lc, rc, g.measured_dcm(t).numpy(),
                              np.ascontiguousarray(traj[2 * t: 2 * t + 2 * (N + 1)]))
    z, st, it, fl = mb.tick(*args(0, np.array([0, 1], np.uint8), np.array([0, 1], np.uint8)))
    assert fl[0] == 1 and st[0] == -10 and it[0] == 0 and (z[:, 0] == 0).all()    # no foot in contact: flagged, nothing solved
    assert fl[1] == 16 and st[1] == 1
    z, st, it, fl = mb.tick(*args(1, np.array([1, 0], np.uint8), np.array([1, 0], np.uint8)))
    assert fl[0] == 16 and fl[1] == 1                                              # "None foot is in contact"
    mb.reset(np.array([1, 0], np.uint8))
    z, st, it, fl = mb.tick(*args(2, np.array([1, 1], np.uint8), np.array([1, 1], np.uint8)))
    assert fl[0] == 16 and fl[1] == 16


@pytest.mark.parametrize("checks", [0, 1])
def test_osqp_eigen_solve_semantics(oracle, syn, checks):
    """OsqpEigen::Solver::solve() under the two osqp-eigen behaviours (wc_osqp_settings.osqp_eigen_checks_status):
    0 = the reference's time, solve() true for SOLVED_INACCURATE / MAX_ITER_REACHED and m_output takes the last iterate;
    1 = osqp-eigen >= 0.4, solve() false unless OSQP_SOLVED and m_output (= the next tick's u_prev) is kept."""
    N, B, T = 10, 16, 6
    g = syn.GaitBatch(B, T, N, seed=9)
    traj = g.trajectory().numpy()
    cfg = oracle.mpc_default_config(N)
    cfg.osqp.eps_abs = cfg.osqp.eps_rel = 1e-5
    cfg.osqp.max_iter = 25
    cfg.osqp.osqp_eigen_checks_status = checks
    mb = oracle.MpcBatch(cfg, B)
    prev = None
    seen = set()
    for t in range(T):
        f = g.feet(t)
        z, st, it, fl = mb.tick(f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(),
                                g.measured_dcm(t).numpy(), np.ascontiguousarray(traj[2 * t: 2 * t + 2 * (N + 1)]))
        seen |= set(st.tolist())
        unsolved = st != 1
        if checks:
            assert ((fl & 4) != 0)[unsolved].all() and ((fl & 4) == 0)[~unsolved].all()
            if prev is not None:
                assert np.array_equal(z[:, unsolved], prev[:, unsolved])          # m_output kept
        else:
            assert ((fl & 4) == 0).all()                                          # solve() true for every status met here
            if prev is not None and unsolved.any():
                assert not np.array_equal(z[:, unsolved], prev[:, unsolved])      # the last iterate became m_output
        prev = z.copy()
    assert {-2, 2} <= seen
