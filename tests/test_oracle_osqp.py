"""The OSQP restatement (oracle/osqp_ref.c) checked by independent means: exact KKT solutions,
known infeasible problems, agreement of the two linear-system back ends."""
import numpy as np
import pytest
import scipy.linalg as sla
from scipy.optimize import lsq_linear, minimize


def exact_eq_qp(P, q, A, b):
    n, m = P.shape[0], A.shape[0]
    K = np.block([[P, A.T], [A, np.zeros((m, m))]])
    sol = np.linalg.solve(K, np.concatenate([-q, b]))
    return sol[:n], sol[n:]


def random_qp(rng, n, m_eq, m_in):
    M = rng.standard_normal((n, n))
    P = M @ M.T + 0.1 * np.eye(n)
    q = rng.standard_normal(n)
    A = rng.standard_normal((m_eq + m_in, n))
    x0 = rng.standard_normal(n)
    l = np.empty(m_eq + m_in); u = np.empty(m_eq + m_in)
    l[:m_eq] = u[:m_eq] = A[:m_eq] @ x0
    l[m_eq:] = A[m_eq:] @ x0 - rng.uniform(0.1, 1.0, m_in)
    u[m_eq:] = A[m_eq:] @ x0 + rng.uniform(0.1, 1.0, m_in)
    return P, q, A, l, u


def tight(oracle, **kw):
    s = oracle.osqp_default_settings()
    s.eps_abs = 1e-9; s.eps_rel = 1e-9; s.max_iter = 20000
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def test_equality_constrained_matches_kkt(oracle):
    rng = np.random.default_rng(0)
    for trial in range(5):
        P, q, A, l, u = random_qp(rng, 12, 5, 0)
        r = oracle.osqp_solve_dense(P, A, q, l, u, tight(oracle))
        x, _ = exact_eq_qp(P, q, A, l)
        assert r["status"] == 1
        assert np.abs(r["x"] - x).max() < 1e-6


def test_inequality_qp_satisfies_kkt(oracle):
    rng = np.random.default_rng(1)
    for trial in range(5):
        P, q, A, l, u = random_qp(rng, 10, 3, 8)
        r = oracle.osqp_solve_dense(P, A, q, l, u, tight(oracle))
        assert r["status"] == 1
        x, y = r["x"], r["y"]
        assert np.abs(P @ x + q + A.T @ y).max() < 1e-5           # stationarity
        Ax = A @ x
        assert (Ax >= l - 1e-6).all() and (Ax <= u + 1e-6).all()     # primal feasibility
        # complementarity: y>0 only at the upper bound, y<0 only at the lower bound
        assert (np.abs(y) * np.minimum(np.abs(Ax - l), np.abs(u - Ax)) < 1e-5).all()
        assert ((y <= 1e-7) | (np.abs(u - Ax) < 1e-5)).all() and ((y >= -1e-7) | (np.abs(Ax - l) < 1e-5)).all()
        # and it agrees with an independent solver
        res = minimize(lambda z: 0.5 * z @ P @ z + q @ z, x0=np.zeros(10), jac=lambda z: P @ z + q, method="SLSQP",
                       constraints=[{"type": "ineq", "fun": lambda z: A @ z - l, "jac": lambda z: A},
                                    {"type": "ineq", "fun": lambda z: u - A @ z, "jac": lambda z: -A}],
                       options={"ftol": 1e-14, "maxiter": 500})
        assert np.abs(res.x - x).max() < 1e-4


def test_default_tolerance_is_loose_but_consistent(oracle):
    """with osqp's default eps=1e-3 the iterate stops early: iterations are a multiple of 25 and the
    unscaled residuals satisfy the termination test"""
    rng = np.random.default_rng(2)
    P, q, A, l, u = random_qp(rng, 10, 3, 8)
    r = oracle.osqp_solve_dense(P, A, q, l, u)
    assert r["status"] == 1 and r["iter"] % 25 == 0
    x = r["x"]
    z = np.clip(A @ x, l, u)
    assert np.abs(A @ x - z).max() <= 1e-3 + 1e-3 * max(np.abs(A @ x).max(), np.abs(z).max()) + 1e-9


def test_primal_infeasible_detected(oracle):
    P = np.eye(2); q = np.zeros(2)
    A = np.array([[1.0, 0.0], [1.0, 0.0]])
    INF = 1e30  # OSQP_INFTY: the C API takes finite sentinels (OsqpEigen::INFTY in the reference, MPCSolver.cpp:49)
    r = oracle.osqp_solve_dense(P, A, q, np.array([1.0, -INF]), np.array([INF, 0.0]))
    assert r["status"] == -3 and np.isnan(r["x"]).all()


def test_dual_infeasible_detected(oracle):
    P = np.zeros((2, 2)); P[0, 0] = 1.0
    q = np.array([0.0, -1.0])
    A = np.array([[1.0, 0.0]])
    r = oracle.osqp_solve_dense(P, A, q, np.array([-1.0]), np.array([1.0]))
    assert r["status"] == -4


def test_max_iter_status(oracle):
    rng = np.random.default_rng(3)
    P, q, A, l, u = random_qp(rng, 10, 3, 8)
    s = tight(oracle, eps_abs=1e-14, eps_rel=1e-14, max_iter=30)
    r = oracle.osqp_solve_dense(P, A, q, l, u, s)
    assert r["status"] in (-2, 2) and r["iter"] == 30


def test_kkt_ldl_and_reduced_cholesky_backends_agree(oracle):
    """OSQP's quasi-definite KKT LDL^T (what QDLDL does) against the reduced SPD system the GPU kernel
    factorises: same iterates up to rounding -> same status / iteration count / solution"""
    rng = np.random.default_rng(4)
    for trial in range(4):
        P, q, A, l, u = random_qp(rng, 14, 4, 9)
        a = oracle.osqp_solve_dense(P, A, q, l, u, linsys=0)
        b = oracle.osqp_solve_dense(P, A, q, l, u, linsys=1)
        assert a["status"] == b["status"] and a["iter"] == b["iter"] and a["rho_updates"] == b["rho_updates"]
        assert np.abs(a["x"] - b["x"]).max() < 1e-9


def test_adaptive_rho_triggers_and_stays_correct(oracle):
    rng = np.random.default_rng(5)
    P, q, A, l, u = random_qp(rng, 12, 4, 6)
    P *= 1e4
    r = oracle.osqp_solve_dense(P, A, q, l, u, tight(oracle))
    assert r["status"] == 1 and r["rho_updates"] >= 1 and r["rho"] != 0.1
    assert np.abs(P @ r["x"] + q + A.T @ r["y"]).max() < 1e-4
