"""Kinematics restatement (finite differences, invariances) and the exact QP / QP-IK restatement (KKT)."""
import numpy as np
import pytest
import torch


def _H(x, y, z, yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([c, -s, 0, x, s, c, 0, y, 0, 0, 1, z])


def _rotvec(R0, R1):
    M = R1.reshape(3, 3) @ R0.reshape(3, 3).T
    return 0.5 * np.array([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])


@pytest.mark.parametrize("left", [1, 0])
def test_jacobians_by_finite_differences(oracle, left):
    m = oracle.model_icub_like()
    rng = np.random.default_rng(left)
    q = np.array(oracle.ik_default_config().joint_reg_rad[:23]) + rng.uniform(-0.3, 0.3, 23)
    bH = _H(0.4, -0.3, 0.0, 0.7)
    f = oracle.fk(m, q, bH, left)
    eps = 1e-6
    for j in range(23):
        qp = q.copy(); qp[j] += eps; qm = q.copy(); qm[j] -= eps
        fp, fm = oracle.fk(m, qp, bH, left), oracle.fk(m, qm, bH, left)
        for key, J in (("lf_H", "J_lf"), ("rf_H", "J_rf")):
            lin = (fp[key][[3, 7, 11]] - fm[key][[3, 7, 11]]) / (2 * eps)
            Rp = fp[key].reshape(3, 4)[:, :3]; Rm = fm[key].reshape(3, 4)[:, :3]
            ang = _rotvec(Rm, Rp) / (2 * eps)
            assert np.abs(lin - f[J][:3, 6 + j]).max() < 1e-7
            assert np.abs(ang - f[J][3:, 6 + j]).max() < 1e-7
        ang = _rotvec(fm["neck_R"], fp["neck_R"]) / (2 * eps)
        assert np.abs(ang - f["J_neck"][3:, 6 + j]).max() < 1e-7
        assert np.abs((fp["com"] - fm["com"]) / (2 * eps) - f["J_com"][:, 6 + j]).max() < 1e-7


def test_base_columns_and_stance_foot(oracle):
    """the stance sole coincides with the given base transform; its Jacobian has no joint columns;
    base columns are [I, -S(r); 0, I] (MIXED representation)"""
    m = oracle.model_icub_like()
    q = np.array(oracle.ik_default_config().joint_reg_rad[:23])
    bH = _H(1.0, 2.0, 0.0, -1.1)
    f = oracle.fk(m, q, bH, 1)
    assert np.abs(f["lf_H"] - bH).max() < 1e-14
    assert np.abs(f["J_lf"][:, 6:]).max() == 0.0
    assert np.allclose(f["J_lf"][:3, :3], np.eye(3)) and np.allclose(f["J_lf"][3:, 3:6], np.eye(3))
    assert np.allclose(f["J_com"][:, :3], np.eye(3))
    g = oracle.fk(m, q, bH, 0)
    assert np.abs(g["rf_H"] - bH).max() < 1e-14 and np.abs(g["J_rf"][:, 6:]).max() == 0.0
    # rigid motion of the base moves everything rigidly
    bH2 = _H(1.5, 2.0, 0.0, -1.1)
    f2 = oracle.fk(m, q, bH2, 1)
    assert np.allclose(f2["com"] - f["com"], [0.5, 0, 0]) and np.allclose(f2["rf_H"][[3, 7, 11]] - f["rf_H"][[3, 7, 11]], [0.5, 0, 0])


def kkt_check(H, g, A, b, lb, ub, x, tol=1e-7):
    """x is optimal iff it is feasible and the reduced gradient on the free variables vanishes with
    correctly signed multipliers on the active bounds"""
    n = len(g)
    assert np.abs(A @ x - b).max() < tol
    fin_l = lb > -1e19; fin_u = ub < 1e19
    assert (x[fin_l] >= lb[fin_l] - 1e-9).all() and (x[fin_u] <= ub[fin_u] + 1e-9).all()
    at_l = fin_l & (np.abs(x - lb) < 1e-9); at_u = fin_u & (np.abs(x - ub) < 1e-9)
    grad = H @ x + g
    # grad + A' lam - mu_l + mu_u = 0 with mu >= 0 supported on the active bounds
    act = np.nonzero(at_l | at_u)[0]
    Mx = np.hstack([A.T, np.eye(n)[:, act]])
    sol, *_ = np.linalg.lstsq(Mx, -grad, rcond=None)
    assert np.abs(Mx @ sol + grad).max() < 1e-6
    mu = sol[A.shape[0]:]
    for k, j in enumerate(act):
        if at_l[j] and not at_u[j]:
            assert mu[k] <= 1e-7     # multiplier of x_j >= lb_j enters with a minus sign
        if at_u[j] and not at_l[j]:
            assert mu[k] >= -1e-7


def test_exact_qp_random_problems(oracle):
    rng = np.random.default_rng(0)
    for trial in range(30):
        n, mE = 12, 5
        M = rng.standard_normal((n, n - 3)); H = M @ M.T            # singular PSD, like the IK Hessian
        A = rng.standard_normal((mE, n))
        Z = np.linalg.svd(A)[2][mE:].T
        if np.linalg.eigvalsh(Z.T @ H @ Z).min() < 1e-6:
            continue
        g = rng.standard_normal(n); x0 = rng.uniform(-0.5, 0.5, n); b = A @ x0
        lb = np.full(n, -1e308); ub = np.full(n, 1e308)
        lb[3:] = -rng.uniform(0.6, 1.5, n - 3); ub[3:] = rng.uniform(0.6, 1.5, n - 3)
        st, x, nwsr = oracle.qp_solve(H, g, A, b, lb, ub)
        assert st == 0
        kkt_check(H, g, A, b, lb, ub, x)


def test_exact_qp_infeasible_and_nonconvex(oracle):
    n = 4
    H = np.eye(n); g = np.zeros(n)
    A = np.array([[1.0, 1.0, 0, 0]]); b = np.array([5.0])
    lb = np.full(n, -1.0); ub = np.full(n, 1.0)
    st, x, _ = oracle.qp_solve(H, g, A, b, lb, ub)
    assert st == 2                                                   # bounds cannot meet x0+x1 = 5
    H2 = np.diag([1.0, 1.0, 0.0, 0.0])
    st, x, _ = oracle.qp_solve(H2, g, A, np.array([0.5]), np.full(n, -1e308), np.full(n, 1e308))
    assert st == 3                                                   # not positive definite on null(A)
    st, x, _ = oracle.qp_solve(H, g, np.array([[1.0, 0, 0, 0], [2.0, 0, 0, 0]]), np.array([1.0, 2.0]), lb, ub)
    assert st == 2                                                   # dependent equality rows are rejected


def _ik_case(oracle, syn, B, seed, perturb=1.0):
    m = oracle.model_icub_like(); cfg = oracle.ik_default_config()
    qreg = torch.tensor(list(cfg.joint_reg_rad[:23]), dtype=torch.float64)

    def fkb(ql, bH, lif):
        o = oracle.fk_batch(m, ql.numpy(), bH.numpy(), lif.numpy(), jacobians=False)
        return {k: torch.from_numpy(v) for k, v in o.items()}

    inp = {k: v.numpy() for k, v in syn.ik_inputs(B, seed, qreg, fkb, perturb=perturb).items()}
    return m, cfg, inp


def test_qpik_solution_satisfies_kkt_of_the_reference_qp(oracle, syn):
    """H = Lambda + Jn' W Jn, g, A = [J_lf; J_rf; J_com], b as in WalkingQPIK_qpOASES (:135-282)"""
    B = 24
    m, cfg, inp = _ik_case(oracle, syn, B, 4)
    for i in range(B):
        kin = oracle.fk(m, inp["q_lin"][:, i], inp["base_H"][:, i], inp["left_is_fixed"][i])
        r = oracle.ik_solve(cfg, inp["q_meas"][:, i], kin["lf_H"], kin["rf_H"], kin["neck_R"], inp["com_meas"][:, i],
                            kin["J_lf"], kin["J_rf"], kin["J_neck"], kin["J_com"], inp["lf_H_des"][:, i],
                            inp["rf_H_des"][:, i], inp["lf_twist"][:, i], inp["rf_twist"][:, i], inp["com_pos_des"][:, i],
                            inp["com_vel_des"][:, i], inp["neck_R_des"][:, i])
        assert r["status"] == 0
        H, g, A, b = r["H"], r["g"], r["A"], r["b"]
        # structure of the reference QP
        Jn = kin["J_neck"][3:]
        lam = np.concatenate([np.zeros(6), np.array(cfg.joint_reg_weight[:23])])
        assert np.allclose(H, np.diag(lam) + Jn.T @ (5.0 * np.eye(3)) @ Jn)
        assert np.allclose(A, np.vstack([kin["J_lf"], kin["J_rf"], kin["J_com"]]))
        assert np.abs(H[:3, :3]).max() == 0.0                        # singular base-linear block
        lb = np.concatenate([np.full(6, -1e308), np.array(cfg.min_vel[:23])])
        ub = np.concatenate([np.full(6, 1e308), np.array(cfg.max_vel[:23])])
        kkt_check(H, g, A, b, lb, ub, r["v"])
        assert np.allclose(r["dq"], r["v"][6:]) and np.abs(r["lf_err"]).max() < 1e-9   # feet rows are equalities


def test_qpik_foot_task_semantics(oracle, syn):
    """zero errors, zero twists, q = q_reg, neck aligned -> zero joint velocities"""
    m = oracle.model_icub_like(); cfg = oracle.ik_default_config()
    q = np.array(cfg.joint_reg_rad[:23]); bH = _H(0, 0, 0, 0.3)
    kin = oracle.fk(m, q, bH, 1)
    Radd = np.array(cfg.additional_rotation).reshape(3, 3)
    neck_des = kin["neck_R"].reshape(3, 3) @ Radd.T
    r = oracle.ik_solve(cfg, q, kin["lf_H"], kin["rf_H"], kin["neck_R"], kin["com"], kin["J_lf"], kin["J_rf"], kin["J_neck"],
                        kin["J_com"], kin["lf_H"], kin["rf_H"], np.zeros(6), np.zeros(6), kin["com"], np.zeros(3),
                        neck_des.reshape(-1))
    assert r["status"] == 0 and np.abs(r["v"]).max() < 1e-12
