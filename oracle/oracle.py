"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so, the CPU restatement of the reference hot path
(see oracle/README.md).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.

PARITY UNPINNED: the reference holds no tests/golden vectors for this path and its solvers
(OSQP, qpOASES, iDynTree) are not available offline; see DESIGN.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

WCO_MAX_HULL = 8
WC_MAX_LINKS = 40
WC_MAX_DOFS = 32


class OsqpSettings(C.Structure):
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
                ("eps_abs", C.c_double), ("eps_rel", C.c_double),
                ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
                ("adaptive_rho_tolerance", C.c_double), ("infty", C.c_double),
                ("max_iter", C.c_int32), ("scaling", C.c_int32), ("check_termination", C.c_int32),
                ("adaptive_rho", C.c_int32), ("adaptive_rho_interval", C.c_int32),
                ("warm_start", C.c_int32), ("osqp_eigen_checks_status", C.c_int32), ("reserved1", C.c_int32)]


class MpcConfig(C.Structure):
    _fields_ = [("horizon", C.c_int32), ("reserved", C.c_int32),
                ("sampling_time", C.c_double), ("com_height", C.c_double), ("gravity", C.c_double),
                ("Q", C.c_double * 4), ("R", C.c_double * 4),
                ("foot_x", C.c_double * 2), ("foot_y", C.c_double * 2),
                ("initial_zmp", C.c_double * 2), ("hull_tolerance", C.c_double),
                ("osqp", OsqpSettings)]


class Model(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("n_dofs", C.c_int32),
                ("parent", C.c_int32 * WC_MAX_LINKS), ("dof", C.c_int32 * WC_MAX_LINKS),
                ("joint_R", (C.c_double * 9) * WC_MAX_LINKS),
                ("joint_p", (C.c_double * 3) * WC_MAX_LINKS),
                ("axis", (C.c_double * 3) * WC_MAX_LINKS),
                ("mass", C.c_double * WC_MAX_LINKS),
                ("com", (C.c_double * 3) * WC_MAX_LINKS),
                ("l_sole_link", C.c_int32), ("r_sole_link", C.c_int32), ("neck_link", C.c_int32),
                ("reserved", C.c_int32),
                ("l_sole_R", C.c_double * 9), ("l_sole_p", C.c_double * 3),
                ("r_sole_R", C.c_double * 9), ("r_sole_p", C.c_double * 3),
                ("neck_R", C.c_double * 9), ("neck_p", C.c_double * 3)]


class IkConfig(C.Structure):
    _fields_ = [("n_dofs", C.c_int32), ("use_com_as_constraint", C.c_int32),
                ("com_weight", C.c_double * 9), ("neck_weight", C.c_double * 9),
                ("additional_rotation", C.c_double * 9),
                ("joint_reg_rad", C.c_double * WC_MAX_DOFS),
                ("joint_reg_weight", C.c_double * WC_MAX_DOFS),
                ("joint_reg_gain", C.c_double * WC_MAX_DOFS),
                ("k_pos_foot", C.c_double), ("k_att_foot", C.c_double),
                ("k_neck", C.c_double), ("k_pos_com", C.c_double),
                ("min_vel", C.c_double * WC_MAX_DOFS), ("max_vel", C.c_double * WC_MAX_DOFS),
                ("nwsr", C.c_int32), ("reserved", C.c_int32)]


def build(force=False):
    """Compile oracle/liboracle.so (gcc, plain C) if missing or stale."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "wc_capi.h"))
    srcs.append(os.path.join(_HERE, "..", "walking-controllers_b200", "csrc", "wc_defaults.c"))
    stale = force or not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        up = C.POINTER(C.c_uint8)
        L.wc_mpc_default_config.argtypes = [C.POINTER(MpcConfig), C.c_int32]
        L.wc_osqp_default_settings.argtypes = [C.POINTER(OsqpSettings)]
        L.wc_ik_default_config.argtypes = [C.POINTER(IkConfig)]
        L.wc_model_icub_like.argtypes = [C.POINTER(Model)]
        L.wco_hull_build.argtypes = [dp, C.c_int, dp, dp, dp]
        L.wco_hull_build.restype = C.c_int
        L.wco_hull_margin.argtypes = [dp, dp, dp, C.c_int, dp]
        L.wco_hull_margin.restype = C.c_double
        L.wco_mpc_batch_create.argtypes = [C.POINTER(MpcConfig), C.c_int]
        L.wco_mpc_batch_create.restype = C.c_void_p
        L.wco_mpc_batch_free.argtypes = [C.c_void_p]
        L.wco_mpc_batch_get.argtypes = [C.c_void_p, C.c_int]
        L.wco_mpc_batch_get.restype = C.c_void_p
        L.wco_mpc_batch_set_weights.argtypes = [C.c_void_p, dp, dp]
        L.wco_mpc_batch_reset.argtypes = [C.c_void_p, up]
        L.wco_mpc_batch_tick.argtypes = [C.c_void_p, dp, dp, up, up, dp, dp, up, dp, ip, ip, ip, C.c_int]
        L.wco_mpc_batch_get_solution.argtypes = [C.c_void_p, dp, dp, ip]
        L.wco_mpc_set_linsys.argtypes = [C.c_void_p, C.c_int]
        L.wco_mpc_n_hull.argtypes = [C.c_void_p]
        L.wco_mpc_n_hull.restype = C.c_int
        L.wco_mpc_get_hull.argtypes = [C.c_void_p, dp, dp]
        L.wco_mpc_get_qp_dense.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
        L.wco_mpc_get_qp_dense.restype = C.c_int
        L.wco_osqp_solve_dense.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp,
                                           C.POINTER(OsqpSettings), C.c_int, dp, dp, dp]
        L.wco_osqp_solve_dense.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def mpc_default_config(horizon=25):
    cfg = MpcConfig()
    lib().wc_mpc_default_config(C.byref(cfg), horizon)
    return cfg


def osqp_default_settings():
    s = OsqpSettings()
    lib().wc_osqp_default_settings(C.byref(s))
    return s


def ik_default_config():
    cfg = IkConfig()
    lib().wc_ik_default_config(C.byref(cfg))
    return cfg


def model_icub_like():
    m = Model()
    lib().wc_model_icub_like(C.byref(m))
    return m


def hull_build(points):
    pts = _f64(points).reshape(-1, 2)
    A = np.zeros((WCO_MAX_HULL, 2)); b = np.zeros(WCO_MAX_HULL); v = np.zeros((WCO_MAX_HULL, 2))
    nh = lib().wco_hull_build(_dp(pts), pts.shape[0], _dp(A), _dp(b), _dp(v))
    return A[:nh].copy(), b[:nh].copy(), v[:nh].copy()


def hull_margin(A, b, v, z):
    A = _f64(A); b = _f64(b); v = _f64(v); z = _f64(z)
    return lib().wco_hull_margin(_dp(A), _dp(b), _dp(v), A.shape[0], _dp(z))


def osqp_solve_dense(P, A, q, l, u, settings=None, linsys=0):
    P = _f64(P); A = _f64(A); q = _f64(q); l = _f64(l); u = _f64(u)
    n, m = P.shape[0], A.shape[0]
    s = settings or osqp_default_settings()
    x = np.zeros(n); y = np.zeros(max(m, 1)); info = np.zeros(5)
    st = lib().wco_osqp_solve_dense(n, m, _dp(P), _dp(A), _dp(q), _dp(l), _dp(u), C.byref(s), linsys,
                                    _dp(x), _dp(y), _dp(info))
    return dict(status=st, x=x, y=y[:m], iter=int(info[0]), rho_updates=int(info[1]), rho=info[2],
                pri_res=info[3], dua_res=info[4])


class MpcBatch:
    """Batched oracle MPC with the same SoA layout as wc_mpc_tick_host (include/wc_capi.h)."""

    def __init__(self, cfg, B, linsys=0, hull_variant=0):
        self.cfg, self.B, self.N = cfg, B, cfg.horizon
        self.h = lib().wco_mpc_batch_create(C.byref(cfg), B)
        if linsys:
            for i in range(B):
                lib().wco_mpc_set_linsys(lib().wco_mpc_batch_get(self.h, i), linsys)
        if hull_variant:  # sensitivity studies only: bit 0 = unit-normal hull rows, bits 1.. = cyclic rotation of their order
            lib().wco_mpc_set_hull_variant.argtypes = [C.c_void_p, C.c_int]
            for i in range(B):
                lib().wco_mpc_set_hull_variant(lib().wco_mpc_batch_get(self.h, i), hull_variant)

    def __del__(self):
        if getattr(self, "h", None):
            lib().wco_mpc_batch_free(self.h)
            self.h = None

    def set_weights(self, q_diag, r_diag):
        q = _f64(q_diag); r = _f64(r_diag)
        lib().wco_mpc_batch_set_weights(self.h, _dp(q), _dp(r))

    def reset(self, mask=None):
        m = _u8(mask) if mask is not None else None
        lib().wco_mpc_batch_reset(self.h, _up(m))

    def tick(self, lf_H, rf_H, l_contact, r_contact, dcm, ref, reset=None, threads=1):
        B = self.B
        lf_H = _f64(lf_H); rf_H = _f64(rf_H); dcm = _f64(dcm); ref = _f64(ref)
        lc = _u8(l_contact); rc = _u8(r_contact)
        rs = _u8(reset) if reset is not None else None
        assert lf_H.size == 12 * B and ref.size == 2 * (self.N + 1) * B
        zmp = np.zeros((2, B)); st = np.zeros(B, np.int32); it = np.zeros(B, np.int32); fl = np.zeros(B, np.int32)
        lib().wco_mpc_batch_tick(self.h, _dp(lf_H), _dp(rf_H), _up(lc), _up(rc), _dp(dcm), _dp(ref), _up(rs),
                                 _dp(zmp), _ip(st), _ip(it), _ip(fl), threads)
        return zmp, st, it, fl

    def solution(self):
        n = 4 * self.N + 2
        primal = np.zeros((n, self.B)); rho = np.zeros(self.B); ru = np.zeros(self.B, np.int32)
        lib().wco_mpc_batch_get_solution(self.h, _dp(primal), _dp(rho), _ip(ru))
        return primal, rho, ru

    def qp_dense(self, i):
        n = 4 * self.N + 2
        mmax = 2 * (self.N + 1) + WCO_MAX_HULL
        P = np.zeros((n, n)); A = np.zeros((mmax, n)); q = np.zeros(n); l = np.zeros(mmax); u = np.zeros(mmax)
        inst = lib().wco_mpc_batch_get(self.h, i)
        m = lib().wco_mpc_get_qp_dense(inst, _dp(P), _dp(A), _dp(q), _dp(l), _dp(u))
        A = A.reshape(-1)[: m * n].reshape(m, n).copy()
        return P, A, q, l[:m].copy(), u[:m].copy()

    def hull(self, i):
        inst = lib().wco_mpc_batch_get(self.h, i)
        nh = lib().wco_mpc_n_hull(inst)
        A = np.zeros((WCO_MAX_HULL, 2)); b = np.zeros(WCO_MAX_HULL)
        lib().wco_mpc_get_hull(inst, _dp(A), _dp(b))
        return A[:nh].copy(), b[:nh].copy()


# ------------------------------------------------------------------------------------ kinematics / QP-IK
def _bind_ik():
    L = lib()
    if getattr(L, "_ik_bound", False):
        return L
    dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int32); up = C.POINTER(C.c_uint8)
    L.wco_fk.argtypes = [C.POINTER(Model), dp, dp, C.c_int] + [dp] * 8
    L.wco_fk_batch.argtypes = [C.POINTER(Model), C.c_int, dp, dp, up] + [dp] * 8
    L.wco_qp_solve.argtypes = [C.c_int, dp, dp, C.c_int, dp, dp, dp, dp, C.c_int, dp, ip]
    L.wco_qp_solve.restype = C.c_int
    L.wco_ik_solve.argtypes = [C.POINTER(IkConfig)] + [dp] * 16 + [dp, dp, dp, ip] + [dp] * 5
    L.wco_ik_solve.restype = C.c_int
    L.wco_ik_batch_tick.argtypes = [C.POINTER(IkConfig), C.POINTER(Model), C.c_int, dp, dp, dp, up] + [dp] * 8 + \
                                   [dp, dp, dp, dp, dp, ip, ip, C.c_int]
    L._ik_bound = True
    return L


def fk(model, q, base_H, left_is_fixed):
    """single instance; returns dict of row-major arrays"""
    L = _bind_ik()
    nV = model.n_dofs + 6
    q = _f64(q); base_H = _f64(base_H)
    o = dict(lf_H=np.zeros(12), rf_H=np.zeros(12), neck_R=np.zeros(9), com=np.zeros(3),
             J_lf=np.zeros((6, nV)), J_rf=np.zeros((6, nV)), J_neck=np.zeros((6, nV)), J_com=np.zeros((3, nV)))
    L.wco_fk(C.byref(model), _dp(q), _dp(base_H), int(left_is_fixed), _dp(o["lf_H"]), _dp(o["rf_H"]), _dp(o["neck_R"]),
             _dp(o["com"]), _dp(o["J_lf"]), _dp(o["J_rf"]), _dp(o["J_neck"]), _dp(o["J_com"]))
    return o


def fk_batch(model, q, base_H, left_is_fixed, jacobians=True):
    """SoA batch: q [nJ,B], base_H [12,B], left_is_fixed [B] -> dict of [rows,B] arrays"""
    L = _bind_ik()
    q = _f64(q); base_H = _f64(base_H); lif = _u8(left_is_fixed)
    B = q.shape[1]; nV = model.n_dofs + 6
    o = dict(lf_H=np.zeros((12, B)), rf_H=np.zeros((12, B)), neck_R=np.zeros((9, B)), com=np.zeros((3, B)))
    if jacobians:
        o.update(J_lf=np.zeros((6 * nV, B)), J_rf=np.zeros((6 * nV, B)), J_neck=np.zeros((6 * nV, B)),
                 J_com=np.zeros((3 * nV, B)))
    L.wco_fk_batch(C.byref(model), B, _dp(q), _dp(base_H), _up(lif), _dp(o["lf_H"]), _dp(o["rf_H"]), _dp(o["neck_R"]),
                   _dp(o["com"]), _dp(o.get("J_lf")), _dp(o.get("J_rf")), _dp(o.get("J_neck")), _dp(o.get("J_com")))
    return o


def qp_solve(H, g, A, b, lb, ub, nwsr_max=100):
    L = _bind_ik()
    H = _f64(H); g = _f64(g); A = _f64(A); b = _f64(b); lb = _f64(lb); ub = _f64(ub)
    n = H.shape[0]; mE = A.shape[0] if A.size else 0
    x = np.zeros(n); nw = np.zeros(1, np.int32)
    st = L.wco_qp_solve(n, _dp(H), _dp(g), mE, _dp(A), _dp(b), _dp(lb), _dp(ub), nwsr_max, _dp(x), _ip(nw))
    return st, x, int(nw[0])


def ik_solve(cfg, q_meas, lf_H, rf_H, neck_R, com_meas, J_lf, J_rf, J_neck, J_com, lf_H_des, rf_H_des, lf_twist,
             rf_twist, com_pos_des, com_vel_des, neck_R_des):
    """single instance with caller-supplied kinematics; also returns the assembled QP"""
    L = _bind_ik()
    nJ = cfg.n_dofs; nV = nJ + 6; nC = 15 if cfg.use_com_as_constraint else 12
    a = [_f64(v) for v in (q_meas, lf_H, rf_H, neck_R, com_meas, J_lf, J_rf, J_neck, J_com, lf_H_des, rf_H_des,
                           lf_twist, rf_twist, com_pos_des, com_vel_des, neck_R_des)]
    dq = np.zeros(nJ); le = np.zeros(6); re = np.zeros(6); nw = np.zeros(1, np.int32)
    H = np.zeros((nV, nV)); g = np.zeros(nV); A = np.zeros((nC, nV)); b = np.zeros(nC); v = np.zeros(nV)
    st = L.wco_ik_solve(C.byref(cfg), *[_dp(x) for x in a], _dp(dq), _dp(le), _dp(re), _ip(nw),
                        _dp(H), _dp(g), _dp(A), _dp(b), _dp(v))
    return dict(status=st, dq=dq, lf_err=le, rf_err=re, nwsr=int(nw[0]), H=H, g=g, A=A, b=b, v=v)


def ik_batch_tick(cfg, model, inp, kpos=None, katt=None, threads=1):
    """inp: dict with the SoA arrays of wc_ik_tick_host (numpy); returns dict"""
    L = _bind_ik()
    B = inp["q_meas"].shape[1]; nJ = cfg.n_dofs
    keys = ("q_meas", "q_lin", "base_H")
    a = {k: _f64(inp[k]) for k in ("q_meas", "q_lin", "base_H", "lf_H_des", "rf_H_des", "lf_twist", "rf_twist",
                                   "com_pos_des", "com_vel_des", "com_meas", "neck_R_des")}
    lif = _u8(inp["left_is_fixed"])
    kp = _f64(kpos) if kpos is not None else None
    ka = _f64(katt) if katt is not None else None
    dq = np.zeros((nJ, B)); le = np.zeros((6, B)); re = np.zeros((6, B))
    st = np.zeros(B, np.int32); nw = np.zeros(B, np.int32)
    L.wco_ik_batch_tick(C.byref(cfg), C.byref(model), B, _dp(a["q_meas"]), _dp(a["q_lin"]), _dp(a["base_H"]), _up(lif),
                        _dp(a["lf_H_des"]), _dp(a["rf_H_des"]), _dp(a["lf_twist"]), _dp(a["rf_twist"]),
                        _dp(a["com_pos_des"]), _dp(a["com_vel_des"]), _dp(a["com_meas"]), _dp(a["neck_R_des"]),
                        _dp(kp), _dp(ka), _dp(dq), _dp(le), _dp(re), _ip(st), _ip(nw), threads)
    return dict(dq=dq, lf_err=le, rf_err=re, status=st, nwsr=nw)


# ---- closed-loop glue (StableDCMModel + WalkingZMPController + joint integrator) -------------------------------
class GlueConfig(C.Structure):
    _fields_ = [("n_dofs", C.c_int32), ("use_gain_scheduling", C.c_int32), ("sampling_time", C.c_double),
                ("com_height", C.c_double), ("gravity", C.c_double), ("k_zmp", C.c_double), ("k_com", C.c_double),
                ("smoothing_time", C.c_double), ("k_zmp_stance", C.c_double), ("k_com_stance", C.c_double)]


def glue_default_config():
    L = lib()
    L.wc_glue_default_config.argtypes = [C.POINTER(GlueConfig)]
    cfg = GlueConfig()
    L.wc_glue_default_config(C.byref(cfg))
    return cfg


class GlueBatch:
    """numpy SoA batch: the checker of walking-controllers_b200.GlueBatch"""

    def __init__(self, cfg, B):
        L = lib()
        dp = C.POINTER(C.c_double)
        L.wco_glue_create.argtypes = [C.POINTER(GlueConfig), C.c_int]; L.wco_glue_create.restype = C.c_void_p
        L.wco_glue_free.argtypes = [C.c_void_p]
        L.wco_glue_reset.argtypes = [C.c_void_p, dp, dp]
        L.wco_glue_com_reference.argtypes = [C.c_void_p] + [dp] * 10
        L.wco_glue_integrate_joints.argtypes = [C.c_void_p, dp, dp]
        self.L, self.cfg, self.B, self.nJ = L, cfg, B, cfg.n_dofs
        self.h = L.wco_glue_create(C.byref(cfg), B)

    def __del__(self):
        try:
            if self.h:
                self.L.wco_glue_free(self.h); self.h = None
        except Exception:
            pass

    def reset(self, com_xy=None, q=None):
        a = _f64(com_xy) if com_xy is not None else None
        b = _f64(q) if q is not None else None
        self.L.wco_glue_reset(self.h, _dp(a), _dp(b))

    def set_phase(self, dcm_vel_des=None, stance=None):
        """WalkingZMPController::setPhase: stance flags [B], or the |dcm_vel_des| < 0.001 rule of updateModule"""
        self.L.wco_glue_set_phase.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint8)]
        v = None if dcm_vel_des is None else _f64(dcm_vel_des)
        s = None if stance is None else np.ascontiguousarray(stance, dtype=np.uint8)
        self.L.wco_glue_set_phase(self.h, _dp(v), None if s is None else s.ctypes.data_as(C.POINTER(C.c_uint8)))

    def com_reference(self, dcm_des, zmp_des, zmp_meas, com_meas, com_z_des=None, com_z_vel_des=None, k_zmp=None, k_com=None):
        arrs = [None if x is None else _f64(x) for x in (dcm_des, zmp_des, zmp_meas, com_meas, com_z_des, com_z_vel_des, k_zmp, k_com)]
        pos = np.zeros((3, self.B)); vel = np.zeros((3, self.B))
        self.L.wco_glue_com_reference(self.h, *[_dp(x) for x in arrs], _dp(pos), _dp(vel))
        return dict(com_pos_des=pos, com_vel_des=vel)

    def integrate_joints(self, dq):
        dq = _f64(dq)
        out = np.zeros((self.nJ, self.B))
        self.L.wco_glue_integrate_joints(self.h, _dp(dq), _dp(out))
        return out

    def reset_com_filter(self):
        self.L.wco_glue_reset_com_filter.argtypes = [C.c_void_p]
        self.L.wco_glue_reset_com_filter(self.h)

    def evaluate_com(self, com, J_com, dq, cut_frequency=0.0, use_filters=False):
        """WalkingFK::evaluateCoM + evaluateDCM -> dict(com, com_vel [3,B], dcm [2,B])"""
        dp = C.POINTER(C.c_double)
        self.L.wco_glue_evaluate_com.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, C.c_int, dp, dp, dp]
        a, b, c = _f64(com), _f64(J_com), _f64(dq)
        o = dict(com=np.zeros((3, self.B)), com_vel=np.zeros((3, self.B)), dcm=np.zeros((2, self.B)))
        self.L.wco_glue_evaluate_com(self.h, _dp(a), _dp(b), _dp(c), cut_frequency, int(use_filters), _dp(o["com"]), _dp(o["com_vel"]), _dp(o["dcm"]))
        return o


# ---- rank 4: WalkingDCMReactiveController, evaluateDCM, evaluateZMP, logger row ---------------------------------
LOG_COLUMNS = 53
LOG_FIELDS = ("dcm", "dcm_des", "dcm_vel_des", "zmp", "zmp_des", "com", "com_des_xy", "com_vel_des_xy",
              "lf_H", "rf_H", "lf_H_des", "rf_H_des", "lf_err", "rf_err")


class LogInputs(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_double)) for k in LOG_FIELDS]


def reactive_dcm(omega, k_dcm, dcm_des, dcm_vel_des, dcm_meas):
    """k_dcm: float or [B] array"""
    L = lib(); dp = C.POINTER(C.c_double)
    L.wco_reactive_dcm.argtypes = [C.c_int, C.c_double, C.c_double, dp, dp, dp, dp, dp]
    a, b, c = _f64(dcm_des), _f64(dcm_vel_des), _f64(dcm_meas)
    B = a.shape[1]
    karr = None if np.isscalar(k_dcm) else _f64(k_dcm)
    out = np.zeros((2, B))
    L.wco_reactive_dcm(B, omega, float(k_dcm) if karr is None else 0.0, _dp(karr), _dp(a), _dp(b), _dp(c), _dp(out))
    return out


def evaluate_dcm(omega, com, com_vel):
    L = lib(); dp = C.POINTER(C.c_double)
    L.wco_evaluate_dcm.argtypes = [C.c_int, C.c_double, dp, dp, dp]
    a, b = _f64(com), _f64(com_vel)
    out = np.zeros((2, a.shape[1]))
    L.wco_evaluate_dcm(a.shape[1], omega, _dp(a), _dp(b), _dp(out))
    return out


def evaluate_zmp(l_wrench, r_wrench, lf_H, rf_H):
    L = lib(); dp = C.POINTER(C.c_double)
    L.wco_evaluate_zmp.argtypes = [C.c_int, dp, dp, dp, dp, dp, C.POINTER(C.c_int)]
    a, b, c, d = _f64(l_wrench), _f64(r_wrench), _f64(lf_H), _f64(rf_H)
    B = a.shape[1]
    out = np.zeros((2, B)); ok = np.zeros(B, dtype=np.int32)
    L.wco_evaluate_zmp(B, _dp(a), _dp(b), _dp(c), _dp(d), _dp(out), ok.ctypes.data_as(C.POINTER(C.c_int)))
    return out, ok


def log_pack(B, **fields):
    L = lib()
    L.wco_log_pack.argtypes = [C.c_int, C.POINTER(LogInputs), C.POINTER(C.c_double)]
    keep = {k: _f64(v) for k, v in fields.items() if v is not None}
    li = LogInputs(**{k: _dp(keep.get(k)) for k in LOG_FIELDS})
    out = np.zeros((LOG_COLUMNS, B))
    L.wco_log_pack(B, C.byref(li), _dp(out))
    return out


# ---- rank 3: synthetic producer of the path's inputs (CPU twin of wc_traj_*) ---------------------------------------
class TrajConfig(C.Structure):
    _fields_ = [("step_ticks", C.c_int32), ("ds_ticks", C.c_int32), ("n_steps", C.c_int32), ("reserved", C.c_int32),
                ("sampling_time", C.c_double), ("com_height", C.c_double), ("gravity", C.c_double),
                ("nominal_width", C.c_double), ("step_height", C.c_double), ("zmp_delta", C.c_double * 2)]


class TrajFeet(C.Structure):
    _fields_ = [("lf_H", C.c_void_p), ("rf_H", C.c_void_p), ("lf_twist", C.c_void_p), ("rf_twist", C.c_void_p),
                ("l_contact", C.c_void_p), ("r_contact", C.c_void_p), ("left_is_fixed", C.c_void_p),
                ("com_height", C.c_void_p), ("com_height_vel", C.c_void_p)]


def traj_default_config(n_ticks, horizon):
    L = lib()
    L.wc_traj_default_config.argtypes = [C.POINTER(TrajConfig), C.c_int32, C.c_int32]
    cfg = TrajConfig()
    L.wc_traj_default_config(C.byref(cfg), n_ticks, horizon)
    return cfg


def philox4x32_10(counter, key):
    L = lib()
    L.wco_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]
    c = (C.c_uint32 * 4)(*counter)
    L.wco_philox4x32_10(c, key[0], key[1])
    return list(c)


class TrajBatch:
    def __init__(self, cfg, B):
        L = lib()
        L.wco_traj_create.argtypes = [C.POINTER(TrajConfig), C.c_int]; L.wco_traj_create.restype = C.c_void_p
        L.wco_traj_free.argtypes = [C.c_void_p]
        L.wco_traj_set_gait.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.wco_traj_dcm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.wco_traj_measured_dcm.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_void_p]
        L.wco_traj_feet.argtypes = [C.c_void_p, C.c_int, C.POINTER(TrajFeet)]
        self.L, self.cfg, self.B = L, cfg, B
        self.h = L.wco_traj_create(C.byref(cfg), B)

    def __del__(self):
        try:
            if self.h:
                self.L.wco_traj_free(self.h); self.h = None
        except Exception:
            pass

    def set_gait(self, step_len, phase=None, dtheta=None, theta0=None, first_left=None):
        p = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)
        arrs = [p(phase, np.int32), p(step_len, np.float64), p(dtheta, np.float64), p(theta0, np.float64), p(first_left, np.uint8)]
        self.L.wco_traj_set_gait(self.h, *[None if a is None else a.ctypes.data for a in arrs])

    def dcm(self, first_sample, n_samples):
        pos = np.zeros((2 * n_samples, self.B)); vel = np.zeros((2 * n_samples, self.B))
        self.L.wco_traj_dcm(self.h, first_sample, n_samples, pos.ctypes.data, vel.ctypes.data)
        return pos, vel

    def measured_dcm(self, tick, sigma, seed):
        out = np.zeros((2, self.B))
        self.L.wco_traj_measured_dcm(self.h, tick, sigma, seed, out.ctypes.data)
        return out

    def feet(self, tick):
        B = self.B
        o = dict(lf_H=np.zeros((12, B)), rf_H=np.zeros((12, B)), lf_twist=np.zeros((6, B)), rf_twist=np.zeros((6, B)),
                 l_contact=np.zeros(B, np.uint8), r_contact=np.zeros(B, np.uint8), left_is_fixed=np.zeros(B, np.uint8),
                 com_height=np.zeros(B), com_height_vel=np.zeros(B))
        f = TrajFeet(**{k: v.ctypes.data for k, v in o.items()})
        self.L.wco_traj_feet(self.h, tick, C.byref(f))
        return o


# ---- rank 2: WalkingQPIK_osqp (persistent per-instance OSQP solvers) ----------------------------------------------
IK_KEYS = ("q_meas", "q_lin", "base_H", "lf_H_des", "rf_H_des", "lf_twist", "rf_twist", "com_pos_des", "com_vel_des",
           "com_meas", "neck_R_des")


def ik_build_qp(cfg, variant, q_meas, lf_H, rf_H, neck_R, com_meas, J_lf, J_rf, J_neck, J_com, lf_H_des, rf_H_des, lf_twist,
                rf_twist, com_pos_des, com_vel_des, neck_R_des):
    """H, g, A, b of the QP-IK (variant 0 = qpOASES, 1 = osqp) for one instance"""
    L = _bind_ik()
    nV = cfg.n_dofs + 6; nC = 15 if cfg.use_com_as_constraint else 12
    a = [_f64(v) for v in (q_meas, lf_H, rf_H, neck_R, com_meas, J_lf, J_rf, J_neck, J_com, lf_H_des, rf_H_des,
                           lf_twist, rf_twist, com_pos_des, com_vel_des, neck_R_des)]
    H = np.zeros((nV, nV)); g = np.zeros(nV); A = np.zeros((nC, nV)); b = np.zeros(nC)
    L.wco_ik_build_qp.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_double)] * 20
    L.wco_ik_build_qp(C.byref(cfg), variant, *[_dp(x) for x in a], _dp(H), _dp(g), _dp(A), _dp(b))
    return H, g, A, b


class IkOsqpBatch:
    def __init__(self, cfg, model, B, settings=None):
        L = _bind_ik()
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.wco_ikosqp_batch_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]; L.wco_ikosqp_batch_create.restype = C.c_void_p
        L.wco_ikosqp_batch_free.argtypes = [C.c_void_p]
        L.wco_ikosqp_batch_tick.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_ubyte)] + [dp] * 8 + [dp] * 3 + [ip] * 4
        self.L, self.cfg, self.B = L, cfg, B
        self.h = L.wco_ikosqp_batch_create(C.byref(cfg), None if settings is None else C.byref(settings), C.byref(model), B)

    def __del__(self):
        try:
            if self.h:
                self.L.wco_ikosqp_batch_free(self.h); self.h = None
        except Exception:
            pass

    def tick(self, inp):
        B, nJ = self.B, self.cfg.n_dofs
        a = {k: _f64(inp[k]) for k in IK_KEYS}
        lif = _u8(inp["left_is_fixed"])
        dq = np.zeros((nJ, B)); le = np.zeros((6, B)); re = np.zeros((6, B))
        st = np.zeros(B, np.int32); it = np.zeros(B, np.int32); fl = np.zeros(B, np.int32); su = np.zeros(B, np.int32)
        self.L.wco_ikosqp_batch_tick(self.h, _dp(a["q_meas"]), _dp(a["q_lin"]), _dp(a["base_H"]), _up(lif),
                                     _dp(a["lf_H_des"]), _dp(a["rf_H_des"]), _dp(a["lf_twist"]), _dp(a["rf_twist"]),
                                     _dp(a["com_pos_des"]), _dp(a["com_vel_des"]), _dp(a["com_meas"]), _dp(a["neck_R_des"]),
                                     _dp(dq), _dp(le), _dp(re), _ip(st), _ip(it), _ip(fl), _ip(su))
        return dict(dq=dq, lf_err=le, rf_err=re, status=st, iters=it, flags=fl, setups=su)


class ClosedLoop:
    """CHECKER of walking-controllers_b200.LoopBatch: the same closed-loop tick (WalkingModule::updateModule order,
    WalkingModule.cpp:541-745) stitched from the CPU restatements -- producer twin, kin_ref FK, mpc_ref, glue_ref (with the
    gain schedule), qpik_ref -- with the same synthetic plant (q_meas = q_des, measured DCM = reference + noise (+ push),
    measured ZMP = desired + offset)."""

    def __init__(self, mpc_cfg, ik_cfg, model, glue_cfg, traj, B, sigma=0.002, seed=0x5EED, zmp_offset=0.001, threads=1):
        self.mpc = MpcBatch(mpc_cfg, B)
        self.glue = GlueBatch(glue_cfg, B)
        self.ik_cfg, self.model, self.traj, self.B, self.N = ik_cfg, model, traj, B, mpc_cfg.horizon
        self.sigma, self.seed, self.off, self.threads = sigma, seed, zmp_offset, threads
        self.t = 0
        self.push_tick = np.full(B, -1, np.int32); self.push_dxy = np.zeros((2, B))
        self.ik_gains = None  # per-instance (k_posFoot, k_attFoot) of a gain sweep

    @staticmethod
    def _pre(f):
        lif = f["left_is_fixed"]
        base = np.where(lif[None, :] != 0, f["lf_H"], f["rf_H"])
        yl = np.arctan2(f["lf_H"][4], f["lf_H"][0]); yr = np.arctan2(f["rf_H"][4], f["rf_H"][0])
        my = np.arctan2(np.sin(yl) + np.sin(yr), np.cos(yl) + np.cos(yr))
        c, s = np.cos(my), np.sin(my); z, o = np.zeros_like(c), np.ones_like(c)
        neck = np.stack([c, -s, z, s, c, z, z, z, o])
        return np.ascontiguousarray(base), neck

    def reset(self, q0):
        self.q = np.array(q0, dtype=np.float64, copy=True)
        self.mpc.reset()
        f = self.traj.feet(0)
        base, _ = self._pre(f)
        com = fk_batch(self.model, self.q, base, f["left_is_fixed"], jacobians=False)["com"]
        self.glue.reset(com[:2], self.q)
        self.t = 0

    def set_push(self, push_tick, push_dxy):
        self.push_tick = np.asarray(push_tick, np.int32).copy(); self.push_dxy = np.asarray(push_dxy, np.float64).copy()

    def tick(self):
        t, N = self.t, self.N
        f = self.traj.feet(t)
        ref, vel = self.traj.dcm(t, N + 1)
        dcm = self.traj.measured_dcm(t, self.sigma, self.seed)
        hit = self.push_tick == t
        dcm[:, hit] += self.push_dxy[:, hit]
        base, neck = self._pre(f)
        com = fk_batch(self.model, self.q, base, f["left_is_fixed"], jacobians=False)["com"]
        z, st, it, fl = self.mpc.tick(f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], dcm, ref, threads=self.threads)
        self.glue.set_phase(dcm_vel_des=vel[0:2])
        co = self.glue.com_reference(ref[0:2], z, z + self.off, com, f["com_height"], f["com_height_vel"])
        inp = dict(q_meas=self.q, q_lin=self.q, base_H=base, left_is_fixed=f["left_is_fixed"], lf_H_des=f["lf_H"], rf_H_des=f["rf_H"],
                   lf_twist=f["lf_twist"], rf_twist=f["rf_twist"], com_pos_des=co["com_pos_des"], com_vel_des=co["com_vel_des"],
                   com_meas=com, neck_R_des=neck)
        kp, ka = self.ik_gains if self.ik_gains is not None else (None, None)
        oo = ik_batch_tick(self.ik_cfg, self.model, inp, kpos=kp, katt=ka, threads=self.threads)
        self.q = self.glue.integrate_joints(oo["dq"])
        self.t += 1
        return dict(zmp=z, osqp_status=st, iters=it, flags=fl, dq=oo["dq"], qp_status=oo["status"], nwsr=oo["nwsr"], q_des=self.q,
                    com_pos_des=co["com_pos_des"], com_vel_des=co["com_vel_des"], com_meas=com, dcm_meas=dcm)
