"""WalkingQPIK_osqp (WalkingQPInverseKinematics_osqp.cpp:135-454 over osqp-eigen / OSQP).

CPU: the restatement (oracle/qpik_osqp_ref.c) — its QP build against the three documented divergences from the qpOASES variant,
its solution against the exact equality-constrained optimum, the solver life cycle (one osqp_setup, then updates; a new setup
when the stance foot — hence the Hessian's sparsity pattern — changes) and the fact that the empty velocity rows do NOT limit the
joint velocities.  Parity unpinned: neither OSQP nor osqp-eigen exists in this container and the reference holds no vectors.
GPU: wc_ik_set_solver(OSQP) against the restatement over multi-tick sequences with warm start and stance switches: identical
OSQP status, iteration count, flags and setup count on every instance and tick; joint velocities within 1e-5 rad/s
(north_star's IK tolerance; observed ~1e-9 — the two sides factor different but equivalent linear systems)."""
import numpy as np
import pytest
import torch


def _inputs(syn, oracle, B, seed, perturb=1.0):
    cfg, model = oracle.ik_default_config(), oracle.model_icub_like()
    qreg = torch.tensor(list(cfg.joint_reg_rad[:23]), dtype=torch.float64)

    def fkb(ql, bH, lif):
        o = oracle.fk_batch(model, ql.numpy(), bH.numpy(), lif.numpy(), jacobians=False)
        return {k: torch.from_numpy(v) for k, v in o.items()}

    inp = syn.ik_inputs(B, seed, qreg, fkb, perturb=perturb)
    return {k: v.numpy().copy() for k, v in inp.items()}


def _qp_of(oracle, cfg, model, ino, i, variant):
    nV = 29
    k = oracle.fk_batch(model, ino["q_lin"], ino["base_H"], ino["left_is_fixed"])
    col = lambda a: np.ascontiguousarray(a[:, i])
    args = [col(ino["q_meas"]), col(k["lf_H"]), col(k["rf_H"]), col(k["neck_R"]), col(ino["com_meas"]),
            col(k["J_lf"]).reshape(6, nV), col(k["J_rf"]).reshape(6, nV), col(k["J_neck"]).reshape(6, nV), col(k["J_com"]).reshape(3, nV),
            col(ino["lf_H_des"]), col(ino["rf_H_des"]), col(ino["lf_twist"]), col(ino["rf_twist"]), col(ino["com_pos_des"]),
            col(ino["com_vel_des"]), col(ino["neck_R_des"])]
    return oracle.ik_build_qp(cfg, variant, *args), k


def test_osqp_variant_qp_divergences(oracle, syn):
    cfg, model = oracle.ik_default_config(), oracle.model_icub_like()
    ino = _inputs(syn, oracle, 8, 4)
    ino["lf_twist"][:, 0] = 0.0                      # zero desired twist on the left foot of instance 0
    ino["lf_twist"][2:, 1] = [0.3, 0.0, 0.0, 0.1]; ino["lf_twist"][:2, 1] = 0.0   # twist(0) == twist(1) == 0 is enough (:286)
    for i in (0, 1, 2):
        (H0, g0, A0, b0), k = _qp_of(oracle, cfg, model, ino, i, 0)
        (H1, g1, A1, b1), _ = _qp_of(oracle, cfg, model, ino, i, 1)
        assert np.array_equal(H0, H1) and np.array_equal(A0, A1)
        # (1) neck gradient scaled by k_attFoot: g1 - g0 = (k_att - 1) * neck term, and the neck term is g0 minus the regularisation part
        reg = np.zeros(29)
        reg[6:] = -np.array(cfg.joint_reg_weight[:23]) * (np.array(cfg.joint_reg_gain[:23]) * (np.array(cfg.joint_reg_rad[:23]) - ino["q_meas"][:, i]))
        assert np.allclose(g1 - reg, cfg.k_att_foot * (g0 - reg), rtol=1e-12, atol=1e-14)
        # (2) bare twist as bound for a foot with twist(0) == twist(1) == 0
        if i in (0, 1):
            assert np.array_equal(b1[:6], ino["lf_twist"][:, i]) and not np.array_equal(b0[:6], b1[:6])
        else:
            assert np.array_equal(b0[:6], b1[:6])
        assert np.array_equal(b0[6:], b1[6:])


def test_osqp_variant_solution_and_life_cycle(oracle, syn):
    cfg, model = oracle.ik_default_config(), oracle.model_icub_like()
    B = 24
    ino = _inputs(syn, oracle, B, 6)
    for j in range(23):                               # (3) limits far below the optimum's velocities: they must not bind
        cfg.min_vel[j], cfg.max_vel[j] = -0.01, 0.01
    bt = oracle.IkOsqpBatch(cfg, model, B)
    rng = np.random.default_rng(0)
    for t in range(5):
        if t == 3:
            ino["left_is_fixed"] = (1 - ino["left_is_fixed"]).astype(np.uint8)   # stance switch: new Hessian pattern
        o = bt.tick(ino)
        assert (o["status"] == 1).all() and (o["flags"] == 0).all()
        assert (o["iters"] % 25 == 0).all() and (o["iters"] >= 25).all()
        assert (o["setups"] == (1 if t < 3 else 2)).all()
        for i in (0, 7, 23):
            (H, g, A, b), _ = _qp_of(oracle, cfg, model, ino, i, 1)
            K = np.block([[H, A.T], [A, np.zeros((15, 15))]])
            sol = np.linalg.solve(K, np.concatenate([-g, b]))
            # OSQP stops at eps_abs = eps_rel = 1e-3: agreement with the exact optimum at that level
            assert np.abs(sol[6:29] - o["dq"][:, i]).max() < 2e-3 * max(1.0, np.abs(sol[:29]).max())
            assert np.abs(sol[6:29]).max() > 0.05, "limits of 0.01 rad/s would bind if the velocity rows were not empty"
            # get*FootError = bound - J x: the residual of the equality rows, small at an OSQP solution
            assert np.abs(o["lf_err"][:, i]).max() < 5e-3 and np.abs(o["rf_err"][:, i]).max() < 5e-3
        ino["q_meas"] = ino["q_meas"] + 0.002 * rng.normal(size=ino["q_meas"].shape)


def _run_sequence(wc, oracle, syn, cuda_device, B, T, seed, com_in_cost=False, ext_kin=False):
    cfg_o, model_o = oracle.ik_default_config(), oracle.model_icub_like()
    cfg_g = wc.ik_default_config()
    if com_in_cost:
        cfg_o.use_com_as_constraint = 0; cfg_g.use_com_as_constraint = 0
    ino = _inputs(syn, oracle, B, seed)
    bo = oracle.IkOsqpBatch(cfg_o, model_o, B)
    ik = wc.IkBatch(cfg_g, wc.model_icub_like(), B)
    ik.set_solver(wc.SOLVER_OSQP)
    flags = torch.zeros(B, dtype=torch.int32, device=cuda_device); setups = torch.zeros(B, dtype=torch.int32, device=cuda_device)
    ik.osqp_outputs(flags, setups)
    rng = np.random.default_rng(seed)
    worst = 0.0
    for t in range(T):
        if t in (3, 6):   # stance switches on a third of the instances
            sw = rng.random(B) < 0.34
            ino["left_is_fixed"] = np.where(sw, 1 - ino["left_is_fixed"], ino["left_is_fixed"]).astype(np.uint8)
        if t == 4:        # a foot at rest: bare-twist branch
            ino["lf_twist"][:, : B // 2] = 0.0
        oo = bo.tick(ino)
        ing = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_device) for k, v in ino.items()}
        if ext_kin:
            kin = oracle.fk_batch(model_o, ino["q_lin"], ino["base_H"], ino["left_is_fixed"])
            og = ik.solve_with_jacobians({k: torch.from_numpy(v).to(cuda_device) for k, v in kin.items()}, ing)
        else:
            og = ik.tick(ing)
        torch.cuda.synchronize()
        assert np.array_equal(og["status"].cpu().numpy(), oo["status"]), f"tick {t}"
        assert np.array_equal(og["nwsr"].cpu().numpy(), oo["iters"]), f"tick {t}"
        assert np.array_equal(flags.cpu().numpy(), oo["flags"]) and np.array_equal(setups.cpu().numpy(), oo["setups"]), f"tick {t}"
        e = np.abs(og["dq"].cpu().numpy() - oo["dq"]).max()
        worst = max(worst, e)
        assert e <= 1e-5, f"tick {t}: dq differs by {e}"
        assert np.abs(og["lf_err"].cpu().numpy() - oo["lf_err"]).max() <= 1e-5 and np.abs(og["rf_err"].cpu().numpy() - oo["rf_err"]).max() <= 1e-5
        ino["q_meas"] = ino["q_meas"] + 0.002 * rng.normal(size=ino["q_meas"].shape)
        ino["com_vel_des"] = ino["com_vel_des"] + 0.005 * rng.normal(size=ino["com_vel_des"].shape)
    ik.close()
    return worst, oo


@pytest.mark.gpu
def test_ik_osqp_gpu_parity_sequence(wc, oracle, syn, cuda_device):
    worst, last = _run_sequence(wc, oracle, syn, cuda_device, B=300, T=9, seed=12)
    assert (last["setups"] >= 1).all() and last["setups"].max() >= 2
    print(f"worst |dq_gpu - dq_oracle| = {worst:.3e} rad/s")


@pytest.mark.gpu
def test_ik_osqp_gpu_parity_com_in_cost(wc, oracle, syn, cuda_device):
    _run_sequence(wc, oracle, syn, cuda_device, B=96, T=5, seed=13, com_in_cost=True)


@pytest.mark.gpu
def test_ik_osqp_gpu_external_jacobians_and_reset(wc, oracle, syn, cuda_device):
    """the WalkingQPIK drop-in path (set*Jacobian + setRobotState) with the osqp solver; reset_solver = a new object"""
    _run_sequence(wc, oracle, syn, cuda_device, B=64, T=4, seed=14, ext_kin=True)
    B = 32
    cfg_o, model_o = oracle.ik_default_config(), oracle.model_icub_like()
    ino = _inputs(syn, oracle, B, 15)
    ik = wc.IkBatch(wc.ik_default_config(), wc.model_icub_like(), B)
    ik.set_solver(wc.SOLVER_OSQP)
    ing = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_device) for k, v in ino.items()}
    first = {k: v.clone() for k, v in ik.tick(ing).items()}
    second = {k: v.clone() for k, v in ik.tick(ing).items()}      # warm start: same problem, different iterates
    ik.reset_solver()
    again = ik.tick(ing)
    torch.cuda.synchronize()
    assert torch.equal(first["dq"], again["dq"]) and torch.equal(first["nwsr"], again["nwsr"])
    assert not torch.equal(first["dq"], second["dq"])
    # back to the qpOASES variant on the same handle
    ik.set_solver(wc.SOLVER_QPOASES)
    oq = ik.tick(ing)
    ref = oracle.ik_batch_tick(cfg_o, model_o, ino)
    torch.cuda.synchronize()
    assert np.array_equal(oq["status"].cpu().numpy(), ref["status"]) and np.abs(oq["dq"].cpu().numpy() - ref["dq"]).max() <= 1e-8
    ik.close()


def _ext_kin_pair(wc, oracle, syn, cuda_device, B, seed, settings_o=None, settings_g=None, mutate=None):
    """one tick of both sides on caller-supplied (possibly doctored) kinematics"""
    cfg_o, model_o = oracle.ik_default_config(), oracle.model_icub_like()
    ino = _inputs(syn, oracle, B, seed)
    kin = oracle.fk_batch(model_o, ino["q_lin"], ino["base_H"], ino["left_is_fixed"])
    if mutate:
        mutate(kin, ino)
    out_o = dict(status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32), flags=np.zeros(B, np.int32), dq=np.zeros((23, B)))
    for i in range(B):
        col = lambda a: np.ascontiguousarray(a[:, i])
        L = oracle.lib()
        h = L.wco_ikosqp_create
        h.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_void_p]; h.restype = __import__("ctypes").c_void_p
        C = __import__("ctypes")
        obj = L.wco_ikosqp_create(C.byref(cfg_o), None if settings_o is None else C.byref(settings_o))
        dp = C.POINTER(C.c_double)
        L.wco_ikosqp_solve.argtypes = [C.c_void_p] + [dp] * 16 + [dp] * 3 + [C.POINTER(C.c_int)] * 3
        L.wco_ikosqp_free.argtypes = [C.c_void_p]
        args = [col(ino["q_meas"]), col(kin["lf_H"]), col(kin["rf_H"]), col(kin["neck_R"]), col(ino["com_meas"]), col(kin["J_lf"]),
                col(kin["J_rf"]), col(kin["J_neck"]), col(kin["J_com"]), col(ino["lf_H_des"]), col(ino["rf_H_des"]), col(ino["lf_twist"]),
                col(ino["rf_twist"]), col(ino["com_pos_des"]), col(ino["com_vel_des"]), col(ino["neck_R_des"])]
        dq = np.zeros(23); le = np.zeros(6); re = np.zeros(6)
        st, it, fl = C.c_int(0), C.c_int(0), C.c_int(0)
        L.wco_ikosqp_solve(obj, *[a.ctypes.data_as(dp) for a in args], dq.ctypes.data_as(dp), le.ctypes.data_as(dp),
                           re.ctypes.data_as(dp), C.byref(st), C.byref(it), C.byref(fl))
        L.wco_ikosqp_free(obj)
        out_o["status"][i], out_o["iters"][i], out_o["flags"][i] = st.value, it.value, fl.value
        out_o["dq"][:, i] = dq
    ik = wc.IkBatch(wc.ik_default_config(), wc.model_icub_like(), B)
    ik.set_solver(wc.SOLVER_OSQP, settings_g)
    flags = torch.zeros(B, dtype=torch.int32, device=cuda_device)
    ik.osqp_outputs(flags, None)
    og = ik.solve_with_jacobians({k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_device) for k, v in kin.items()},
                                 {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_device) for k, v in ino.items()})
    torch.cuda.synchronize()
    out_g = dict(status=og["status"].cpu().numpy(), iters=og["nwsr"].cpu().numpy(), flags=flags.cpu().numpy(), dq=og["dq"].cpu().numpy())
    ik.close()
    return out_o, out_g


@pytest.mark.gpu
def test_ik_osqp_gpu_adaptive_rho(wc, oracle, syn, cuda_device, rho0=1e-4):
    """a badly chosen initial rho (1e-4) with the rho update every 10 iterations: the estimate moves rho by far more than the 5x
    tolerance, so the re-factorisation path (new K, Cholesky, L^-1 inside the loop) runs — with it OSQP needs ~30 iterations,
    without it > 1000; checks every 5 iterations, eps 1e-5 (well above the 1e-10 noise between the two linear-system methods)"""
    def settings(mod):
        c = mod.mpc_default_config(25)
        s = c.osqp
        s.eps_abs = 1e-5; s.eps_rel = 1e-5; s.check_termination = 5; s.adaptive_rho_interval = 10; s.rho = rho0
        return c, s
    (co, so), (cg, sg) = settings(oracle), settings(wc)
    o, g = _ext_kin_pair(wc, oracle, syn, cuda_device, 48, 21, so, sg)
    co2, so2 = settings(oracle); so2.adaptive_rho = 0
    o_fixed, _ = _ext_kin_pair(wc, oracle, syn, cuda_device, 4, 21, so2, sg)
    assert (o_fixed["iters"] > 10 * o["iters"][:4]).all(), "rho was never updated: the test does not reach the path it is for"
    assert (o["status"] == 1).all()
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["iters"], o["iters"]) and np.array_equal(g["flags"], o["flags"])
    assert np.abs(g["dq"] - o["dq"]).max() <= 1e-8


@pytest.mark.gpu
def test_ik_osqp_gpu_max_iter_and_inaccurate(wc, oracle, syn, cuda_device):
    """max_iter below the first termination check: OSQP_MAX_ITER_REACHED or SOLVED_INACCURATE, decided by the 10x tolerances"""
    so, sg = oracle.mpc_default_config(25).osqp, wc.mpc_default_config(25).osqp
    for s in (so, sg):
        s.max_iter = 7; s.eps_abs = 1e-6; s.eps_rel = 1e-6
    o, g = _ext_kin_pair(wc, oracle, syn, cuda_device, 48, 22, so, sg)
    assert set(np.unique(o["status"])) <= {2, -2} and (o["iters"] == 7).all()
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["iters"], o["iters"])
    assert np.abs(g["dq"] - o["dq"]).max() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("checks", [0, 1])
def test_ik_osqp_gpu_solve_semantics(wc, oracle, syn, cuda_device, checks):
    """m_optimizerSolver->solve() (..._osqp.cpp:377) under the two osqp-eigen behaviours (osqp_eigen_checks_status, the same
    switch and rule as the MPC): flag bit 0 is set for every non-SOLVED status only when the status is checked"""
    so, sg = oracle.mpc_default_config(25).osqp, wc.mpc_default_config(25).osqp
    for s in (so, sg):
        s.max_iter = 7; s.eps_abs = 1e-6; s.eps_rel = 1e-6; s.osqp_eigen_checks_status = checks
    o, g = _ext_kin_pair(wc, oracle, syn, cuda_device, 32, 22, so, sg)
    assert set(np.unique(o["status"])) <= {2, -2}
    assert ((o["flags"] & 1) == checks).all()
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["flags"], o["flags"])


@pytest.mark.gpu
def test_ik_osqp_gpu_primal_infeasible(wc, oracle, syn, cuda_device):
    """identical feet Jacobians with different right-hand sides: contradictory equality rows -> OSQP_PRIMAL_INFEASIBLE"""
    def mutate(kin, ino):
        kin["J_rf"] = kin["J_lf"].copy()
        ino["rf_twist"] = ino["lf_twist"] + 1.0
    o, g = _ext_kin_pair(wc, oracle, syn, cuda_device, 32, 23, mutate=mutate)
    assert (o["status"] == -3).all(), np.unique(o["status"])
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["iters"], o["iters"]) and np.array_equal(g["flags"], o["flags"])
    assert np.isnan(g["dq"]).all() and np.isnan(o["dq"]).all()
