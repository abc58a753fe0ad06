"""walking-controllers_b200 — B200-native batched MPC + QP-IK hot path of walking-controllers.

Python here is plumbing only: a ctypes view of the C ABI in include/wc_capi.h (the product is
walking-controllers_b200/libwc_b200.so, hand-written CUDA for sm_100a) plus torch tensors for
device memory / streams.  The directory name is not a Python identifier; import it with
    importlib.import_module("walking-controllers_b200")
There is NO CPU fallback: every compute entry point raises if the CUDA library is missing.
"""
import collections
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WC_LIB_PATH") or os.path.join(_HERE, "libwc_b200.so")  # override: kernel-tuning sweeps only
CSRC = os.path.join(_HERE, "csrc")

WC_MAX_LINKS = 40
WC_MAX_DOFS = 32

# status / flag constants (include/wc_capi.h)
OSQP_SOLVED = 1
OSQP_SOLVED_INACCURATE = 2
OSQP_MAX_ITER_REACHED = -2
OSQP_PRIMAL_INFEASIBLE = -3
OSQP_DUAL_INFEASIBLE = -4
OSQP_UNSOLVED = -10
MPC_FLAG_NO_CONTACT, MPC_FLAG_NO_SOLVER, MPC_FLAG_NOT_SOLVED, MPC_FLAG_OUTSIDE_HULL, MPC_FLAG_NEW_SOLVER = 1, 2, 4, 8, 16
QP_SUCCESSFUL_RETURN, QP_MAX_NWSR_REACHED, QP_INFEASIBLE, QP_NOT_CONVEX = 0, 1, 2, 3


class TickHostIn(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("lf_H", "rf_H", "l_contact", "r_contact", "left_is_fixed", "reset", "dcm_meas",
                                          "ref_newest", "ref_window", "q_meas", "q_lin", "base_H", "lf_twist", "rf_twist",
                                          "com_pos_des", "com_vel_des", "com_meas", "neck_R_des")]


class TickHostOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("zmp", "osqp_status", "iters", "flags", "dq", "lf_err", "rf_err", "qp_status", "nwsr")]


class LoopConfig(C.Structure):
    _fields_ = [("dcm_noise_sigma", C.c_double), ("noise_seed", C.c_uint64), ("zmp_meas_offset", C.c_double),
                ("use_graph", C.c_int32), ("reserved", C.c_int32)]


LOOP_VIEW = (("zmp", 2, "f8"), ("dq", "nJ", "f8"), ("q_des", "nJ", "f8"), ("com_pos_des", 3, "f8"), ("com_vel_des", 3, "f8"),
             ("com_meas", 3, "f8"), ("dcm_meas", 2, "f8"), ("lf_err", 6, "f8"), ("rf_err", 6, "f8"), ("lf_H", 12, "f8"),
             ("rf_H", 12, "f8"), ("ref_pos", "2N", "f8"), ("osqp_status", 1, "i4"), ("iters", 1, "i4"), ("flags", 1, "i4"),
             ("qp_status", 1, "i4"), ("nwsr", 1, "i4"), ("lf_twist", 6, "f8"), ("rf_twist", 6, "f8"), ("base_H", 12, "f8"),
             ("neck_R_des", 9, "f8"), ("l_contact", 1, "u1"), ("r_contact", 1, "u1"), ("left_is_fixed", 1, "u1"))


class LoopView(C.Structure):
    _fields_ = [(k, C.c_void_p) for k, _, _ in LOOP_VIEW]


class GatherIO(C.Structure):
    _fields_ = [("zmp", C.c_void_p), ("zmp_all", C.c_void_p), ("dq", C.c_void_p), ("dq_all", C.c_void_p),
                ("osqp_status", C.c_void_p), ("osqp_status_all", C.c_void_p), ("iters", C.c_void_p), ("iters_all", C.c_void_p),
                ("qp_status", C.c_void_p), ("qp_status_all", C.c_void_p), ("nwsr", C.c_void_p), ("nwsr_all", C.c_void_p),
                ("stats_sum", C.c_void_p), ("n_stats_sum", C.c_int32), ("stats_max", C.c_void_p), ("n_stats_max", C.c_int32)]


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


class GlueConfig(C.Structure):
    _fields_ = [("n_dofs", C.c_int32), ("use_gain_scheduling", C.c_int32), ("sampling_time", C.c_double),
                ("com_height", C.c_double), ("gravity", C.c_double), ("k_zmp", C.c_double), ("k_com", C.c_double),
                ("smoothing_time", C.c_double), ("k_zmp_stance", C.c_double), ("k_com_stance", C.c_double)]


LOG_COLUMNS = 53
LOG_FIELDS = ("dcm", "dcm_des", "dcm_vel_des", "zmp", "zmp_des", "com", "com_des_xy", "com_vel_des_xy",
              "lf_H", "rf_H", "lf_H_des", "rf_H_des", "lf_err", "rf_err")
# column names of the logger file, WalkingModule.cpp:1231-1250
LOG_NAMES = ("dcm_x dcm_y dcm_des_x dcm_des_y dcm_des_dx dcm_des_dy zmp_x zmp_y zmp_des_x zmp_des_y com_x com_y com_z "
             "com_des_x com_des_y com_des_dx com_des_dy lf_x lf_y lf_z lf_roll lf_pitch lf_yaw rf_x rf_y rf_z rf_roll "
             "rf_pitch rf_yaw lf_des_x lf_des_y lf_des_z lf_des_roll lf_des_pitch lf_des_yaw rf_des_x rf_des_y rf_des_z "
             "rf_des_roll rf_des_pitch rf_des_yaw lf_err_x lf_err_y lf_err_z lf_err_roll lf_err_pitch lf_err_yaw rf_err_x "
             "rf_err_y rf_err_z rf_err_roll rf_err_pitch rf_err_yaw").split()


class TrajConfig(C.Structure):
    _fields_ = [("step_ticks", C.c_int32), ("ds_ticks", C.c_int32), ("n_steps", C.c_int32), ("reserved", C.c_int32),
                ("sampling_time", C.c_double), ("com_height", C.c_double), ("gravity", C.c_double),
                ("nominal_width", C.c_double), ("step_height", C.c_double), ("zmp_delta", C.c_double * 2)]


class TrajFeet(C.Structure):
    _fields_ = [("lf_H", C.c_void_p), ("rf_H", C.c_void_p), ("lf_twist", C.c_void_p), ("rf_twist", C.c_void_p),
                ("l_contact", C.c_void_p), ("r_contact", C.c_void_p), ("left_is_fixed", C.c_void_p),
                ("com_height", C.c_void_p), ("com_height_vel", C.c_void_p)]


class LogInputs(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_double)) for k in LOG_FIELDS]


def write_dataset(path, times, rows):
    """Text file in the WalkingLogger module's format (WalkingLoggerModule.cpp:62-84 header, :163-168 rows):
    `time name ... ` then one `t v0 v1 ... ` line per tick, C++ ostream default formatting (%g, 6 digits).
    rows: [T, 53] host array of one instance."""
    assert len(LOG_NAMES) == LOG_COLUMNS
    with open(path, "w") as f:
        f.write("time " + "".join(n + " " for n in LOG_NAMES) + "\n")
        for t, r in zip(times, rows):
            f.write("%g " % t + "".join("%g " % float(v) for v in r) + "\n")


def build(verbose=False):
    """Compile libwc_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-j8", "-C", CSRC], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libwc_b200.so failed")
    return LIB_PATH


_LIB = None
_VP = C.c_void_p


def lib():
    """The C-ABI library.  Raises if it has not been built: there is no fallback path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the product has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    i32, i64 = C.c_int32, C.c_int64
    L.wc_version.restype = C.c_char_p
    L.wc_osqp_default_settings.argtypes = [C.POINTER(OsqpSettings)]
    L.wc_mpc_default_config.argtypes = [C.POINTER(MpcConfig), i32]
    L.wc_ik_default_config.argtypes = [C.POINTER(IkConfig)]
    L.wc_model_icub_like.argtypes = [C.POINTER(Model)]
    L.wc_mpc_create.argtypes = [C.POINTER(MpcConfig), i32, i32, C.POINTER(_VP)]
    L.wc_mpc_create.restype = i32
    L.wc_mpc_destroy.argtypes = [_VP]
    L.wc_mpc_batch.argtypes = [_VP]; L.wc_mpc_batch.restype = i32
    L.wc_mpc_horizon.argtypes = [_VP]; L.wc_mpc_horizon.restype = i32
    L.wc_mpc_set_weights_host.argtypes = [_VP, _VP, _VP]; L.wc_mpc_set_weights_host.restype = i32
    L.wc_mpc_reset.argtypes = [_VP, _VP]; L.wc_mpc_reset.restype = i32
    L.wc_mpc_tick_dev.argtypes = [_VP] + [_VP] * 11 + [_VP]; L.wc_mpc_tick_dev.restype = i32
    L.wc_mpc_tick_ring_dev.argtypes = [_VP] + [_VP] * 12 + [_VP]; L.wc_mpc_tick_ring_dev.restype = i32
    L.wc_tick_host_async.argtypes = [_VP, _VP, C.POINTER(TickHostIn), C.POINTER(TickHostOut)]; L.wc_tick_host_async.restype = i32
    L.wc_mpc_tick_host.argtypes = [_VP] + [_VP] * 11; L.wc_mpc_tick_host.restype = i32
    L.wc_mpc_tick_host_async.argtypes = [_VP] + [_VP] * 11; L.wc_mpc_tick_host_async.restype = i32
    L.wc_mpc_sync.argtypes = [_VP]; L.wc_mpc_sync.restype = i32
    L.wc_mpc_sync_one.argtypes = [_VP]; L.wc_mpc_sync_one.restype = i32
    L.wc_mpc_debug_get_host.argtypes = [_VP, _VP, _VP, _VP]; L.wc_mpc_debug_get_host.restype = i32
    L.wc_mpc_launch_count.argtypes = [_VP]; L.wc_mpc_launch_count.restype = i64
    L.wc_mpc_enable_profiling.argtypes = [_VP, i32]; L.wc_mpc_enable_profiling.restype = i32
    L.wc_mpc_kernel_times_ms.argtypes = [_VP, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
    L.wc_mpc_kernel_times_ms.restype = i32
    L.wc_mpc_kernel_times_detail_ms.argtypes = [_VP] + [C.POINTER(C.c_double)] * 3
    L.wc_mpc_kernel_times_detail_ms.restype = i32
    L.wc_mpc_path.argtypes = [_VP]; L.wc_mpc_path.restype = i32
    L.wc_glue_default_config.argtypes = [C.POINTER(GlueConfig)]
    L.wc_glue_create.argtypes = [C.POINTER(GlueConfig), i32, i32, C.POINTER(_VP)]; L.wc_glue_create.restype = i32
    L.wc_glue_destroy.argtypes = [_VP]
    L.wc_glue_reset_dev.argtypes = [_VP, _VP, _VP, _VP]; L.wc_glue_reset_dev.restype = i32
    L.wc_glue_com_reference_dev.argtypes = [_VP] + [_VP] * 10 + [_VP]; L.wc_glue_com_reference_dev.restype = i32
    L.wc_glue_set_phase_dev.argtypes = [_VP, _VP, _VP, _VP]; L.wc_glue_set_phase_dev.restype = i32
    L.wc_glue_integrate_joints_dev.argtypes = [_VP, _VP, _VP, _VP]; L.wc_glue_integrate_joints_dev.restype = i32
    L.wc_glue_launch_count.argtypes = [_VP]; L.wc_glue_launch_count.restype = i64
    L.wc_traj_default_config.argtypes = [C.POINTER(TrajConfig), i32, i32]
    L.wc_traj_create.argtypes = [C.POINTER(TrajConfig), i32, i32, C.POINTER(_VP)]; L.wc_traj_create.restype = i32
    L.wc_traj_destroy.argtypes = [_VP]
    L.wc_traj_set_gait_dev.argtypes = [_VP] * 7; L.wc_traj_set_gait_dev.restype = i32
    L.wc_traj_dcm_dev.argtypes = [_VP, i32, i32, _VP, _VP, _VP]; L.wc_traj_dcm_dev.restype = i32
    L.wc_traj_measured_dcm_dev.argtypes = [_VP, i32, C.c_double, C.c_uint64, _VP, _VP]; L.wc_traj_measured_dcm_dev.restype = i32
    L.wc_traj_feet_dev.argtypes = [_VP, i32, C.POINTER(TrajFeet), _VP]; L.wc_traj_feet_dev.restype = i32
    L.wc_traj_generate_host.argtypes = [_VP] * 6 + [i32, i32] + [_VP] * 11; L.wc_traj_generate_host.restype = i32
    L.wc_traj_launch_count.argtypes = [_VP]; L.wc_traj_launch_count.restype = i64
    L.wc_glue_reactive_dcm_dev.argtypes = [_VP, C.c_double] + [_VP] * 6; L.wc_glue_reactive_dcm_dev.restype = i32
    L.wc_glue_reactive_dcm_host.argtypes = [_VP, C.c_double] + [_VP] * 5; L.wc_glue_reactive_dcm_host.restype = i32
    L.wc_glue_evaluate_dcm_dev.argtypes = [_VP] * 5; L.wc_glue_evaluate_dcm_dev.restype = i32
    L.wc_glue_evaluate_com_dev.argtypes = [_VP, _VP, _VP, _VP, C.c_double, i32, _VP, _VP, _VP, _VP]; L.wc_glue_evaluate_com_dev.restype = i32
    L.wc_glue_evaluate_com_host.argtypes = [_VP, _VP, _VP, _VP, C.c_double, i32, _VP, _VP, _VP]; L.wc_glue_evaluate_com_host.restype = i32
    L.wc_glue_reset_com_filter_dev.argtypes = [_VP, _VP]; L.wc_glue_reset_com_filter_dev.restype = i32
    L.wc_glue_evaluate_zmp_dev.argtypes = [_VP] * 8; L.wc_glue_evaluate_zmp_dev.restype = i32
    L.wc_glue_log_pack_dev.argtypes = [_VP, C.POINTER(LogInputs), _VP, _VP]; L.wc_glue_log_pack_dev.restype = i32
    L.wc_measure_fp64_peak.argtypes = [i32, C.POINTER(C.c_double)]; L.wc_measure_fp64_peak.restype = i32
    L.wc_comm_unique_id.argtypes = [_VP]; L.wc_comm_unique_id.restype = i32
    L.wc_comm_create.argtypes = [_VP, i32, i32, i32, C.POINTER(_VP)]; L.wc_comm_create.restype = i32
    L.wc_comm_destroy.argtypes = [_VP]
    L.wc_comm_world.argtypes = [_VP]; L.wc_comm_world.restype = i32
    L.wc_comm_rank.argtypes = [_VP]; L.wc_comm_rank.restype = i32
    L.wc_gather_async.argtypes = [_VP, C.POINTER(GatherIO), i32, i32, _VP]; L.wc_gather_async.restype = i32
    L.wc_gather_wait.argtypes = [_VP, _VP, i32]; L.wc_gather_wait.restype = i32
    L.wc_loop_default_config.argtypes = [C.POINTER(LoopConfig)]
    L.wc_loop_create.argtypes = [_VP, _VP, _VP, _VP, C.POINTER(LoopConfig), i32, C.POINTER(_VP)]; L.wc_loop_create.restype = i32
    L.wc_loop_destroy.argtypes = [_VP]
    L.wc_loop_reset_dev.argtypes = [_VP, _VP, _VP]; L.wc_loop_reset_dev.restype = i32
    L.wc_loop_set_push_dev.argtypes = [_VP, _VP, _VP, _VP]; L.wc_loop_set_push_dev.restype = i32
    L.wc_loop_run_dev.argtypes = [_VP, i32, _VP]; L.wc_loop_run_dev.restype = i32
    L.wc_loop_tick_count.argtypes = [_VP]; L.wc_loop_tick_count.restype = i64
    L.wc_loop_view_get.argtypes = [_VP, C.POINTER(LoopView)]; L.wc_loop_view_get.restype = i32
    L.wc_mpc_set_tick_source_dev.argtypes = [_VP, _VP]; L.wc_mpc_set_tick_source_dev.restype = i32
    L.wc_traj_set_tick_source_dev.argtypes = [_VP, _VP]; L.wc_traj_set_tick_source_dev.restype = i32
    if hasattr(L, "wc_ik_create"):
        L.wc_ik_create.argtypes = [C.POINTER(IkConfig), C.POINTER(Model), i32, i32, C.POINTER(_VP)]
        L.wc_ik_create.restype = i32
        L.wc_ik_destroy.argtypes = [_VP]
        L.wc_ik_batch.argtypes = [_VP]; L.wc_ik_batch.restype = i32
        L.wc_ik_dofs.argtypes = [_VP]; L.wc_ik_dofs.restype = i32
        L.wc_ik_set_gains_host.argtypes = [_VP, _VP, _VP]; L.wc_ik_set_gains_host.restype = i32
        L.wc_ik_tick_dev.argtypes = [_VP] + [_VP] * 17 + [_VP]; L.wc_ik_tick_dev.restype = i32
        L.wc_ik_tick_part_dev.argtypes = [_VP, i32] + [_VP] * 12 + [_VP] + [_VP] * 5 + [_VP]; L.wc_ik_tick_part_dev.restype = i32
        L.wc_ik_tick_is_split.argtypes = [_VP]; L.wc_ik_tick_is_split.restype = i32
        L.wc_ik_tick_host.argtypes = [_VP] + [_VP] * 17; L.wc_ik_tick_host.restype = i32
        L.wc_ik_tick_host_async.argtypes = [_VP] + [_VP] * 17; L.wc_ik_tick_host_async.restype = i32
        L.wc_ik_sync.argtypes = [_VP]; L.wc_ik_sync.restype = i32
        L.wc_ik_sync_one.argtypes = [_VP]; L.wc_ik_sync_one.restype = i32
        L.wc_fk_dev.argtypes = [_VP] + [_VP] * 11 + [_VP]; L.wc_fk_dev.restype = i32
        L.wc_ik_solve_with_jacobians_dev.argtypes = [_VP] + [_VP] * 21 + [_VP]
        L.wc_ik_solve_with_jacobians_dev.restype = i32
        L.wc_ik_launch_count.argtypes = [_VP]; L.wc_ik_launch_count.restype = i64
        L.wc_ik_set_solver.argtypes = [_VP, i32, C.POINTER(OsqpSettings)]; L.wc_ik_set_solver.restype = i32
        L.wc_ik_reset_solver.argtypes = [_VP]; L.wc_ik_reset_solver.restype = i32
        L.wc_ik_osqp_outputs_dev.argtypes = [_VP, _VP, _VP]; L.wc_ik_osqp_outputs_dev.restype = i32
        L.wc_ik_osqp_info_host.argtypes = [_VP, _VP, _VP]; L.wc_ik_osqp_info_host.restype = i32
        L.wc_ik_enable_profiling.argtypes = [_VP, i32]; L.wc_ik_enable_profiling.restype = i32
        L.wc_ik_kernel_times_ms.argtypes = [_VP, C.POINTER(C.c_double), C.POINTER(i64)]
        L.wc_ik_kernel_times_ms.restype = i32
        L.wc_ik_kernel_times_detail_ms.argtypes = [_VP] + [C.POINTER(C.c_double)] * 2
        L.wc_ik_kernel_times_detail_ms.restype = i32
        L.wc_ik_path_counts_host.argtypes = [_VP, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.wc_ik_path_counts_host.restype = i32
    _LIB = L
    return L


def mpc_default_config(horizon=25):
    cfg = MpcConfig()
    lib().wc_mpc_default_config(C.byref(cfg), horizon)
    return cfg


def ik_default_config():
    cfg = IkConfig()
    lib().wc_ik_default_config(C.byref(cfg))
    return cfg


def model_icub_like():
    m = Model()
    lib().wc_model_icub_like(C.byref(m))
    return m


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed with wc_status {rc}")


def _ptr(t):
    """device/host pointer of a torch tensor or numpy array (None -> NULL)"""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


_KINDS = {"f64": ("float64", 8), "u8": ("uint8", 1), "i32": ("int32", 4)}


def _arg(t, kind, n, where, name):
    """Validated pointer for the C ABI (which takes raw pointers and sizes and cannot check anything): dtype `kind`
    ('f64' | 'u8' | 'i32'; bool counts as u8), C-contiguous, at least n elements, and on the side of the bus the entry point
    expects (where = 'cuda' | 'host').  A wrong dtype, a transposed view, a CPU tensor handed to a *_dev call or a short
    output buffer raises here instead of reading or writing out of bounds on the device.  None -> NULL."""
    if t is None:
        return None
    dname, _ = _KINDS[kind]
    if hasattr(t, "data_ptr"):  # torch tensor
        dt = str(t.dtype).replace("torch.", "")
        if dt == "bool" and kind == "u8":
            dt = "uint8"
        if dt != dname:
            raise TypeError(f"{name}: dtype {dt}, the C ABI expects {dname}")
        if not t.is_contiguous():
            raise ValueError(f"{name}: not contiguous (a transposed / strided view); pass .contiguous()")
        if t.numel() < n:
            raise ValueError(f"{name}: {t.numel()} elements, the C ABI reads / writes {n}")
        on_cuda = t.is_cuda
        if where == "cuda" and not on_cuda:
            raise ValueError(f"{name}: a CPU tensor was passed to a device-pointer (*_dev) entry point")
        if where == "host" and on_cuda:
            raise ValueError(f"{name}: a CUDA tensor was passed to a host-pointer entry point")
        return t.data_ptr()
    dt = str(t.dtype)
    if dt == "bool" and kind == "u8":
        dt = "uint8"
    if dt != dname:
        raise TypeError(f"{name}: dtype {dt}, the C ABI expects {dname}")
    if not t.flags["C_CONTIGUOUS"]:
        raise ValueError(f"{name}: not C-contiguous; pass np.ascontiguousarray(...)")
    if t.size < n:
        raise ValueError(f"{name}: {t.size} elements, the C ABI reads / writes {n}")
    if where == "cuda":
        raise ValueError(f"{name}: a numpy array was passed to a device-pointer (*_dev) entry point")
    return t.ctypes.data


class MpcBatch:
    """Batched DCM-MPC handle (mirrors WalkingController for B instances).

    tick(...) takes torch CUDA tensors (field-major SoA, float64 / uint8) and launches on the
    current torch stream; tick_host(...) takes numpy arrays / CPU tensors and copies inside."""

    def __init__(self, cfg, B, device=0):
        self.cfg, self.B, self.N, self.device = cfg, B, cfg.horizon, device
        h = _VP()
        _check(lib().wc_mpc_create(C.byref(cfg), B, device, C.byref(h)), "wc_mpc_create")
        self.h = h
        self._inflight = collections.deque()  # buffers of the host ticks in flight (kept alive until retired)

    def close(self):
        if getattr(self, "h", None):
            lib().wc_mpc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, q_diag, r_diag):
        _check(lib().wc_mpc_set_weights_host(self.h, _ptr(q_diag), _ptr(r_diag)), "wc_mpc_set_weights_host")

    def reset(self, mask=None):
        _check(lib().wc_mpc_reset(self.h, _ptr(mask)), "wc_mpc_reset")

    def tick(self, lf_H, rf_H, l_contact, r_contact, dcm, ref, reset=None, out=None, stream=None):
        import torch
        B = self.B
        if out is None:
            dev = lf_H.device
            out = dict(zmp=torch.empty(2, B, dtype=torch.float64, device=dev),
                       status=torch.empty(B, dtype=torch.int32, device=dev),
                       iters=torch.empty(B, dtype=torch.int32, device=dev),
                       flags=torch.empty(B, dtype=torch.int32, device=dev))
        if stream is None:
            stream = torch.cuda.current_stream(lf_H.device).cuda_stream
        a = self._args(lf_H, rf_H, l_contact, r_contact, dcm, ref, reset, out, "cuda")
        _check(lib().wc_mpc_tick_dev(self.h, *a, stream), "wc_mpc_tick_dev")
        return out

    def _args(self, lf_H, rf_H, l_contact, r_contact, dcm, ref, reset, out, where, ref_rows=None):
        B, N = self.B, self.N
        return (_arg(lf_H, "f64", 12 * B, where, "lf_H"), _arg(rf_H, "f64", 12 * B, where, "rf_H"),
                _arg(l_contact, "u8", B, where, "l_contact"), _arg(r_contact, "u8", B, where, "r_contact"),
                _arg(dcm, "f64", 2 * B, where, "dcm"),
                _arg(ref, "f64", (2 * (N + 1) if ref_rows is None else ref_rows) * B, where, "ref"),
                _arg(reset, "u8", B, where, "reset"), _arg(out["zmp"], "f64", 2 * B, where, "out.zmp"),
                _arg(out["status"], "i32", B, where, "out.status"), _arg(out["iters"], "i32", B, where, "out.iters"),
                _arg(out["flags"], "i32", B, where, "out.flags"))

    def tick_ring(self, lf_H, rf_H, l_contact, r_contact, dcm, ref_newest=None, ref_window=None, reset=None, out=None,
                  stream=None):
        """wc_mpc_tick_ring_dev: the DCM window lives in a device-side ring; per tick only its newest sample
        (ref_newest [2, B]) is passed, the whole window (ref_window) on the first tick and after a re-plan"""
        import torch
        B, N = self.B, self.N
        if out is None:
            dev = lf_H.device
            out = dict(zmp=torch.empty(2, B, dtype=torch.float64, device=dev), status=torch.empty(B, dtype=torch.int32, device=dev),
                       iters=torch.empty(B, dtype=torch.int32, device=dev), flags=torch.empty(B, dtype=torch.int32, device=dev))
        if stream is None:
            stream = torch.cuda.current_stream(lf_H.device).cuda_stream
        a = self._args(lf_H, rf_H, l_contact, r_contact, dcm, ref_newest, reset, out, "cuda", ref_rows=2)
        _check(lib().wc_mpc_tick_ring_dev(self.h, *a[:6], _arg(ref_window, "f64", 2 * (N + 1) * B, "cuda", "ref_window"),
                                          *a[6:], stream), "wc_mpc_tick_ring_dev")
        return out

    def tick_host(self, lf_H, rf_H, l_contact, r_contact, dcm, ref, reset=None, out=None):
        import numpy as np
        B = self.B
        if out is None:
            out = dict(zmp=np.empty((2, B)), status=np.empty(B, np.int32), iters=np.empty(B, np.int32),
                       flags=np.empty(B, np.int32))
        a = self._args(lf_H, rf_H, l_contact, r_contact, dcm, ref, reset, out, "host")
        _check(lib().wc_mpc_tick_host(self.h, *a), "wc_mpc_tick_host")
        return out

    def tick_host_async(self, lf_H, rf_H, l_contact, r_contact, dcm, ref, out, reset=None):
        """enqueue copies + kernels on the handle's stream; the handle keeps the buffers referenced until the tick is
        retired by sync_one() / sync() (a temporary freed before its DMA ran would be read or written after the free)"""
        a = self._args(lf_H, rf_H, l_contact, r_contact, dcm, ref, reset, out, "host")
        _check(lib().wc_mpc_tick_host_async(self.h, *a), "wc_mpc_tick_host_async")
        self._inflight.append((lf_H, rf_H, l_contact, r_contact, dcm, ref, reset, out))
        return out

    def tick_fused_host_async(self, ik, inp, out):
        """wc_tick_host_async: one upload for both blocks.  inp: dict of host (pinned) arrays with the keys of
        wc_tick_host_in (absent / None = NULL where the C ABI allows it); out: dict with the keys of wc_tick_host_out."""
        B, N, nJ = self.B, self.N, ik.nJ
        f = TickHostIn(); o = TickHostOut()
        rows_in = dict(lf_H=12, rf_H=12, dcm_meas=2, ref_newest=2, ref_window=2 * (N + 1), q_meas=nJ, q_lin=nJ, base_H=12,
                       lf_twist=6, rf_twist=6, com_pos_des=3, com_vel_des=3, com_meas=3, neck_R_des=9)
        for k, r in rows_in.items():
            setattr(f, k, _arg(inp.get(k), "f64", r * B, "host", k))
        for k in ("l_contact", "r_contact", "left_is_fixed", "reset"):
            setattr(f, k, _arg(inp.get(k), "u8", B, "host", k))
        rows_out = dict(zmp=2, dq=nJ, lf_err=6, rf_err=6)
        for k, r in rows_out.items():
            setattr(o, k, _arg(out.get(k), "f64", r * B, "host", "out." + k))
        for k in ("osqp_status", "iters", "flags", "qp_status", "nwsr"):
            setattr(o, k, _arg(out.get(k), "i32", B, "host", "out." + k))
        _check(lib().wc_tick_host_async(self.h, ik.h, C.byref(f), C.byref(o)), "wc_tick_host_async")
        self._inflight.append((inp, out))
        return out

    def sync(self):
        _check(lib().wc_mpc_sync(self.h), "wc_mpc_sync")
        self._inflight.clear()

    def sync_one(self):
        """retire the oldest host tick in flight (up to three may be)"""
        _check(lib().wc_mpc_sync_one(self.h), "wc_mpc_sync_one")
        if self._inflight:
            self._inflight.popleft()

    def debug(self):
        import numpy as np
        n = 4 * self.N + 2
        primal = np.empty((n, self.B)); rho = np.empty(self.B); ru = np.empty(self.B, np.int32)
        _check(lib().wc_mpc_debug_get_host(self.h, _ptr(primal), _ptr(rho), _ptr(ru)), "wc_mpc_debug_get_host")
        return primal, rho, ru

    @property
    def launches(self):
        return lib().wc_mpc_launch_count(self.h)

    def enable_profiling(self, on=True):
        _check(lib().wc_mpc_enable_profiling(self.h, 1 if on else 0), "wc_mpc_enable_profiling")

    def kernel_times_ms(self):
        """(prepare_ms, admm_ms, ticks) accumulated since the last call (device-synchronising)"""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        _check(lib().wc_mpc_kernel_times_ms(self.h, C.byref(a), C.byref(b), C.byref(n)), "wc_mpc_kernel_times_ms")
        return a.value, b.value, n.value

    def kernel_times_detail_ms(self):
        """(setup_ms, phase1_ms, phase2_ms) of the window closed by the last kernel_times_ms() call"""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _check(lib().wc_mpc_kernel_times_detail_ms(self.h, C.byref(a), C.byref(b), C.byref(c)), "wc_mpc_kernel_times_detail_ms")
        return a.value, b.value, c.value

    @property
    def path(self):
        return {2: "split", 1: "chain"}.get(lib().wc_mpc_path(self.h), "coop")


IK_INPUT_KEYS = ("q_meas", "q_lin", "base_H", "left_is_fixed", "lf_H_des", "rf_H_des", "lf_twist", "rf_twist",
                 "com_pos_des", "com_vel_des", "com_meas", "neck_R_des")


SOLVER_QPOASES, SOLVER_OSQP = 0, 1


class IkBatch:
    """Batched FK + QP-IK handle (mirrors WalkingFK + WalkingQPIK_qpOASES for B instances)."""

    def __init__(self, cfg, model, B, device=0):
        self.cfg, self.model, self.B, self.nJ, self.device = cfg, model, B, cfg.n_dofs, device
        h = _VP()
        _check(lib().wc_ik_create(C.byref(cfg), C.byref(model), B, device, C.byref(h)), "wc_ik_create")
        self.h = h
        self._inflight = collections.deque()  # buffers of the host ticks in flight (kept alive until retired)

    def close(self):
        if getattr(self, "h", None):
            lib().wc_ik_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_gains(self, k_pos_foot, k_att_foot):
        _check(lib().wc_ik_set_gains_host(self.h, _ptr(k_pos_foot), _ptr(k_att_foot)), "wc_ik_set_gains_host")

    def set_solver(self, solver, settings=None):
        """SOLVER_QPOASES (WalkingQPIK_qpOASES, default) or SOLVER_OSQP (WalkingQPIK_osqp, `use_osqp`): with the latter the
        tick outputs mean status = OSQP status, nwsr = OSQP iterations, and the OSQP workspaces persist between ticks."""
        _check(lib().wc_ik_set_solver(self.h, solver, None if settings is None else C.byref(settings)), "wc_ik_set_solver")

    def reset_solver(self):
        _check(lib().wc_ik_reset_solver(self.h), "wc_ik_reset_solver")

    def osqp_outputs(self, flags=None, setups=None):
        """device int32 [B] tensors written by the following ticks of the osqp variant (None = not wanted)"""
        self._osqp_keep = (flags, setups)
        _check(lib().wc_ik_osqp_outputs_dev(self.h, _ptr(flags), _ptr(setups)), "wc_ik_osqp_outputs_dev")

    def _outs(self, dev):
        import torch
        B, nJ = self.B, self.nJ
        return dict(dq=torch.empty(nJ, B, dtype=torch.float64, device=dev),
                    lf_err=torch.empty(6, B, dtype=torch.float64, device=dev),
                    rf_err=torch.empty(6, B, dtype=torch.float64, device=dev),
                    status=torch.empty(B, dtype=torch.int32, device=dev),
                    nwsr=torch.empty(B, dtype=torch.int32, device=dev))

    def tick(self, inp, out=None, stream=None):
        """inp: dict of CUDA tensors keyed by IK_INPUT_KEYS"""
        import torch
        dev = inp["q_meas"].device
        if out is None:
            out = self._outs(dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        _check(lib().wc_ik_tick_dev(self.h, *self._args(inp, out, "cuda"), stream), "wc_ik_tick_dev")
        return out

    def tick_kinematics(self, inp, com_out=None, stream=None):
        """first half of tick() (wc_ik_tick_part_dev, WC_IK_PART_KINEMATICS): everything that does not need the desired CoM of the
        tick -- inp may leave com_pos_des / com_vel_des / com_meas out; com_out [3, B] receives the CoM of the posture"""
        import torch
        dev = inp["q_meas"].device
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        B, nJ = self.B, self.nJ
        rows = dict(q_meas=nJ, q_lin=nJ, base_H=12, lf_H_des=12, rf_H_des=12, lf_twist=6, rf_twist=6, com_pos_des=3,
                    com_vel_des=3, com_meas=3, neck_R_des=9)
        a = []
        for k in IK_INPUT_KEYS:
            if k in ("com_pos_des", "com_vel_des", "com_meas") and inp.get(k) is None:
                a.append(None)
            elif k == "left_is_fixed":
                a.append(_arg(inp.get(k), "u8", B, "cuda", k))
            else:
                a.append(_arg(inp.get(k), "f64", rows[k] * B, "cuda", k))
        co = _arg(com_out, "f64", 3 * B, "cuda", "com_out") if com_out is not None else None
        _check(lib().wc_ik_tick_part_dev(self.h, 1, *a, co, None, None, None, None, None, stream), "wc_ik_tick_part_dev")
        return com_out

    def tick_solve(self, inp, out=None, stream=None):
        """second half of tick() (WC_IK_PART_SOLVE): the CoM term, the QP solves, the outputs; together with tick_kinematics() on the
        same inputs it gives the bits of tick()"""
        import torch
        dev = inp["q_meas"].device
        if out is None:
            out = self._outs(dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        a = self._args(inp, out, "cuda")
        _check(lib().wc_ik_tick_part_dev(self.h, 2, *a[:12], None, *a[12:], stream), "wc_ik_tick_part_dev")
        return out

    @property
    def tick_is_split(self):
        return bool(lib().wc_ik_tick_is_split(self.h))

    def _args(self, inp, out, where):
        B, nJ = self.B, self.nJ
        rows = dict(q_meas=nJ, q_lin=nJ, base_H=12, lf_H_des=12, rf_H_des=12, lf_twist=6, rf_twist=6, com_pos_des=3,
                    com_vel_des=3, com_meas=3, neck_R_des=9)
        a = [_arg(inp.get(k), "u8", B, where, k) if k == "left_is_fixed" else _arg(inp.get(k), "f64", rows[k] * B, where, k)
             for k in IK_INPUT_KEYS]
        return a + [_arg(out["dq"], "f64", nJ * B, where, "out.dq"), _arg(out["lf_err"], "f64", 6 * B, where, "out.lf_err"),
                    _arg(out["rf_err"], "f64", 6 * B, where, "out.rf_err"), _arg(out["status"], "i32", B, where, "out.status"),
                    _arg(out["nwsr"], "i32", B, where, "out.nwsr")]

    def tick_host(self, inp, out=None):
        import numpy as np
        B, nJ = self.B, self.nJ
        if out is None:
            out = dict(dq=np.empty((nJ, B)), lf_err=np.empty((6, B)), rf_err=np.empty((6, B)),
                       status=np.empty(B, np.int32), nwsr=np.empty(B, np.int32))
        _check(lib().wc_ik_tick_host(self.h, *self._args(inp, out, "host")), "wc_ik_tick_host")
        return out

    def tick_host_async(self, inp, out):
        """buffers stay referenced by the handle until sync_one() / sync() retires the tick"""
        _check(lib().wc_ik_tick_host_async(self.h, *self._args(inp, out, "host")), "wc_ik_tick_host_async")
        self._inflight.append((inp, out))
        return out

    def sync(self):
        _check(lib().wc_ik_sync(self.h), "wc_ik_sync")
        self._inflight.clear()

    def sync_one(self):
        _check(lib().wc_ik_sync_one(self.h), "wc_ik_sync_one")
        if self._inflight:
            self._inflight.popleft()

    def fk(self, q, base_H, left_is_fixed, jacobians=True, stream=None):
        import torch
        dev, B, nV = q.device, self.B, self.nJ + 6
        f64 = torch.float64
        o = dict(lf_H=torch.empty(12, B, dtype=f64, device=dev), rf_H=torch.empty(12, B, dtype=f64, device=dev),
                 neck_R=torch.empty(9, B, dtype=f64, device=dev), com=torch.empty(3, B, dtype=f64, device=dev))
        if jacobians:
            o.update(J_lf=torch.empty(6 * nV, B, dtype=f64, device=dev), J_rf=torch.empty(6 * nV, B, dtype=f64, device=dev),
                     J_neck=torch.empty(6 * nV, B, dtype=f64, device=dev), J_com=torch.empty(3 * nV, B, dtype=f64, device=dev))
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        _arg(q, "f64", self.nJ * B, "cuda", "q"); _arg(base_H, "f64", 12 * B, "cuda", "base_H")
        _arg(left_is_fixed, "u8", B, "cuda", "left_is_fixed")
        _check(lib().wc_fk_dev(self.h, _ptr(q), _ptr(base_H), _ptr(left_is_fixed), _ptr(o["lf_H"]), _ptr(o["rf_H"]),
                               _ptr(o["neck_R"]), _ptr(o["com"]), _ptr(o.get("J_lf")), _ptr(o.get("J_rf")),
                               _ptr(o.get("J_neck")), _ptr(o.get("J_com")), stream), "wc_fk_dev")
        return o

    def solve_with_jacobians(self, kin, inp, out=None, stream=None):
        """kin: dict lf_H, rf_H, neck_R, J_lf, J_rf, J_neck, J_com (CUDA tensors); inp as for tick()"""
        import torch
        dev = inp["q_meas"].device
        if out is None:
            out = self._outs(dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        _check(lib().wc_ik_solve_with_jacobians_dev(
            self.h, _ptr(inp["q_meas"]), _ptr(kin["lf_H"]), _ptr(kin["rf_H"]), _ptr(kin["neck_R"]), _ptr(inp["com_meas"]),
            _ptr(kin["J_lf"]), _ptr(kin["J_rf"]), _ptr(kin["J_neck"]), _ptr(kin["J_com"]),
            _ptr(inp["lf_H_des"]), _ptr(inp["rf_H_des"]), _ptr(inp["lf_twist"]), _ptr(inp["rf_twist"]),
            _ptr(inp["com_pos_des"]), _ptr(inp["com_vel_des"]), _ptr(inp["neck_R_des"]),
            _ptr(out["dq"]), _ptr(out["lf_err"]), _ptr(out["rf_err"]), _ptr(out["status"]), _ptr(out["nwsr"]), stream),
            "wc_ik_solve_with_jacobians_dev")
        return out

    @property
    def launches(self):
        return lib().wc_ik_launch_count(self.h)

    def enable_profiling(self, on=True):
        _check(lib().wc_ik_enable_profiling(self.h, 1 if on else 0), "wc_ik_enable_profiling")

    def kernel_times_ms(self):
        a, n = C.c_double(), C.c_int64()
        _check(lib().wc_ik_kernel_times_ms(self.h, C.byref(a), C.byref(n)), "wc_ik_kernel_times_ms")
        return a.value, n.value

    def kernel_times_detail_ms(self):
        """(fast_ms, general_ms) of the window closed by the last kernel_times_ms() call"""
        a, b = C.c_double(), C.c_double()
        _check(lib().wc_ik_kernel_times_detail_ms(self.h, C.byref(a), C.byref(b)), "wc_ik_kernel_times_detail_ms")
        return a.value, b.value

    def path_counts(self):
        """(second_pass, general): instances of the last fused tick that took the fast kernel's second pass / the general kernel"""
        a, b = C.c_int32(), C.c_int32()
        _check(lib().wc_ik_path_counts_host(self.h, C.byref(a), C.byref(b)), "wc_ik_path_counts_host")
        return a.value, b.value


def glue_default_config():
    cfg = GlueConfig()
    lib().wc_glue_default_config(C.byref(cfg))
    return cfg


class GlueBatch:
    """Batched StableDCMModel + WalkingZMPController + joint integrator (torch CUDA tensors, field-major SoA)."""

    def __init__(self, cfg, B, device=0):
        self.cfg, self.B, self.nJ, self.device = cfg, B, cfg.n_dofs, device
        h = _VP()
        _check(lib().wc_glue_create(C.byref(cfg), B, device, C.byref(h)), "wc_glue_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().wc_glue_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream(stream):
        import torch
        return (stream or torch.cuda.current_stream()).cuda_stream

    def reset(self, com_xy=None, q=None, stream=None):
        _check(lib().wc_glue_reset_dev(self.h, _ptr(com_xy), _ptr(q), self._stream(stream)), "wc_glue_reset_dev")

    def set_phase(self, dcm_vel_des=None, stance=None, stream=None):
        """WalkingZMPController::setPhase (gain scheduling): stance flags [B] uint8, or the |dcm_vel_des| < 0.001 rule"""
        _check(lib().wc_glue_set_phase_dev(self.h, _arg(dcm_vel_des, "f64", 2 * self.B, "cuda", "dcm_vel_des"),
                                           _arg(stance, "u8", self.B, "cuda", "stance"), self._stream(stream)),
               "wc_glue_set_phase_dev")

    def com_reference(self, dcm_des, zmp_des, zmp_meas, com_meas, com_z_des=None, com_z_vel_des=None, k_zmp=None,
                      k_com=None, out=None, stream=None):
        import torch
        if out is None:
            out = dict(com_pos_des=torch.empty(3, self.B, dtype=torch.float64, device=dcm_des.device),
                       com_vel_des=torch.empty(3, self.B, dtype=torch.float64, device=dcm_des.device))
        _check(lib().wc_glue_com_reference_dev(self.h, _ptr(dcm_des), _ptr(zmp_des), _ptr(zmp_meas), _ptr(com_meas),
                                               _ptr(com_z_des), _ptr(com_z_vel_des), _ptr(k_zmp), _ptr(k_com),
                                               _ptr(out["com_pos_des"]), _ptr(out["com_vel_des"]), self._stream(stream)),
               "wc_glue_com_reference_dev")
        return out

    def integrate_joints(self, dq, out=None, stream=None):
        import torch
        if out is None:
            out = torch.empty(self.nJ, self.B, dtype=torch.float64, device=dq.device)
        _check(lib().wc_glue_integrate_joints_dev(self.h, _ptr(dq), _ptr(out), self._stream(stream)),
               "wc_glue_integrate_joints_dev")
        return out

    # ---- SURVEY §8(f) rank 4
    def reactive_dcm(self, k_dcm, dcm_des, dcm_vel_des, dcm_meas, out=None, stream=None):
        """WalkingDCMReactiveController::evaluateControl; k_dcm: float or [B] tensor"""
        import torch
        if out is None:
            out = torch.empty(2, self.B, dtype=torch.float64, device=dcm_des.device)
        kt = k_dcm if isinstance(k_dcm, torch.Tensor) else None
        _check(lib().wc_glue_reactive_dcm_dev(self.h, 0.0 if kt is not None else float(k_dcm), _ptr(kt), _ptr(dcm_des),
                                              _ptr(dcm_vel_des), _ptr(dcm_meas), _ptr(out), self._stream(stream)),
               "wc_glue_reactive_dcm_dev")
        return out

    def evaluate_dcm(self, com, com_vel, out=None, stream=None):
        import torch
        if out is None:
            out = torch.empty(2, self.B, dtype=torch.float64, device=com.device)
        _check(lib().wc_glue_evaluate_dcm_dev(self.h, _ptr(com), _ptr(com_vel), _ptr(out), self._stream(stream)),
               "wc_glue_evaluate_dcm_dev")
        return out

    def reset_com_filter(self, stream=None):
        _check(lib().wc_glue_reset_com_filter_dev(self.h, self._stream(stream)), "wc_glue_reset_com_filter_dev")

    def evaluate_com(self, com, J_com, dq, cut_frequency=0.0, use_filters=False, stream=None):
        """WalkingFK::evaluateCoM + evaluateDCM: CoM velocity from the CoM Jacobian and the joint velocities, the two
        low-pass filters, the measured DCM -> dict(com, com_vel [3,B], dcm [2,B])"""
        import torch
        o = dict(com=torch.empty(3, self.B, dtype=torch.float64, device=com.device),
                 com_vel=torch.empty(3, self.B, dtype=torch.float64, device=com.device),
                 dcm=torch.empty(2, self.B, dtype=torch.float64, device=com.device))
        _check(lib().wc_glue_evaluate_com_dev(self.h, _ptr(com), _ptr(J_com), _ptr(dq), float(cut_frequency), int(use_filters),
                                              _ptr(o["com"]), _ptr(o["com_vel"]), _ptr(o["dcm"]), self._stream(stream)),
               "wc_glue_evaluate_com_dev")
        return o

    def evaluate_zmp(self, l_wrench, r_wrench, lf_H, rf_H, stream=None):
        import torch
        zmp = torch.empty(2, self.B, dtype=torch.float64, device=lf_H.device)
        ok = torch.empty(self.B, dtype=torch.int32, device=lf_H.device)
        _check(lib().wc_glue_evaluate_zmp_dev(self.h, _ptr(l_wrench), _ptr(r_wrench), _ptr(lf_H), _ptr(rf_H), _ptr(zmp),
                                              _ptr(ok), self._stream(stream)), "wc_glue_evaluate_zmp_dev")
        return zmp, ok

    def log_pack(self, out=None, stream=None, **fields):
        """the 53 logger columns of one tick -> [53, B] (fields: LOG_FIELDS, missing ones give zero columns)"""
        import torch
        dev = next(v.device for v in fields.values() if v is not None)
        if out is None:
            out = torch.empty(LOG_COLUMNS, self.B, dtype=torch.float64, device=dev)
        li = LogInputs(**{k: C.cast(_ptr(fields.get(k)), C.POINTER(C.c_double)) for k in LOG_FIELDS})
        _check(lib().wc_glue_log_pack_dev(self.h, C.byref(li), _ptr(out), self._stream(stream)), "wc_glue_log_pack_dev")
        return out

    @property
    def launches(self):
        return lib().wc_glue_launch_count(self.h)


def traj_default_config(n_ticks, horizon):
    cfg = TrajConfig()
    lib().wc_traj_default_config(C.byref(cfg), n_ticks, horizon)
    return cfg


class TrajBatch:
    """On-device producer of the path's per-tick inputs for B periodic gaits (torch CUDA tensors,
    field-major SoA): DCM reference window / velocity, measured DCM, feet poses / twists / contacts, CoM height."""

    def __init__(self, cfg, B, device=0):
        self.cfg, self.B, self.device = cfg, B, device
        h = _VP()
        _check(lib().wc_traj_create(C.byref(cfg), B, device, C.byref(h)), "wc_traj_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().wc_traj_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream(stream):
        import torch
        return (stream or torch.cuda.current_stream()).cuda_stream

    def set_gait(self, step_len, phase=None, dtheta=None, theta0=None, first_left=None, stream=None):
        """step_len [B] f64; phase [B] int32; dtheta, theta0 [B] f64; first_left [B] uint8 (device tensors)"""
        self._keep = (step_len, phase, dtheta, theta0, first_left)
        _check(lib().wc_traj_set_gait_dev(self.h, _ptr(phase), _ptr(step_len), _ptr(dtheta), _ptr(theta0), _ptr(first_left),
                                          self._stream(stream)), "wc_traj_set_gait_dev")

    def dcm(self, first_sample, n_samples, pos=None, vel=None, want_vel=True, stream=None):
        import torch
        dev = torch.device("cuda", self.device)
        if pos is None:
            pos = torch.empty(2 * n_samples, self.B, dtype=torch.float64, device=dev)
        if vel is None and want_vel:
            vel = torch.empty(2 * n_samples, self.B, dtype=torch.float64, device=dev)
        _check(lib().wc_traj_dcm_dev(self.h, first_sample, n_samples, _ptr(pos), _ptr(vel), self._stream(stream)), "wc_traj_dcm_dev")
        return pos, vel

    def measured_dcm(self, tick, sigma, seed, out=None, stream=None):
        import torch
        if out is None:
            out = torch.empty(2, self.B, dtype=torch.float64, device=torch.device("cuda", self.device))
        _check(lib().wc_traj_measured_dcm_dev(self.h, tick, sigma, seed, _ptr(out), self._stream(stream)), "wc_traj_measured_dcm_dev")
        return out

    def feet(self, tick, out=None, stream=None):
        import torch
        dev = torch.device("cuda", self.device)
        if out is None:
            f64, u8, B = torch.float64, torch.uint8, self.B
            out = dict(lf_H=torch.empty(12, B, dtype=f64, device=dev), rf_H=torch.empty(12, B, dtype=f64, device=dev),
                       lf_twist=torch.empty(6, B, dtype=f64, device=dev), rf_twist=torch.empty(6, B, dtype=f64, device=dev),
                       l_contact=torch.empty(B, dtype=u8, device=dev), r_contact=torch.empty(B, dtype=u8, device=dev),
                       left_is_fixed=torch.empty(B, dtype=u8, device=dev), com_height=torch.empty(B, dtype=f64, device=dev),
                       com_height_vel=torch.empty(B, dtype=f64, device=dev))
        f = TrajFeet(**{k: _ptr(out.get(k)) for k, _ in TrajFeet._fields_})
        _check(lib().wc_traj_feet_dev(self.h, tick, C.byref(f), self._stream(stream)), "wc_traj_feet_dev")
        return out

    @property
    def launches(self):
        return lib().wc_traj_launch_count(self.h)


def measure_fp64_peak(device=0):
    """measured FP64 FMA throughput of the device in GFLOP/s"""
    g = C.c_double()
    _check(lib().wc_measure_fp64_peak(device, C.byref(g)), "wc_measure_fp64_peak")
    return g.value


def loop_default_config():
    c = LoopConfig()
    lib().wc_loop_default_config(C.byref(c))
    return c


class _DevArray:
    """a raw device pointer as a CUDA-array-interface object (torch.as_tensor wraps it without a copy)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": ("|" if typestr == "u1" else "<") + typestr,
                                         "data": (ptr, False), "version": 2}


class LoopBatch:
    """The closed-loop control tick as one call (wc_loop_*): producer -> measured CoM -> DCM-MPC -> ZMP-CoM law with gain
    scheduling -> FK + QP-IK -> joint integrator, device resident, CUDA-graph replayed.  Borrows the four block handles."""

    def __init__(self, mpc, ik, glue, traj, cfg=None, device=0):
        self.mpc, self.ik, self.glue, self.traj = mpc, ik, glue, traj  # keep the borrowed handles alive
        self.cfg = cfg if cfg is not None else loop_default_config()
        self.B, self.N, self.nJ, self.device = mpc.B, mpc.N, ik.nJ, device
        h = _VP()
        _check(lib().wc_loop_create(mpc.h, ik.h, glue.h, traj.h, C.byref(self.cfg), device, C.byref(h)), "wc_loop_create")
        self.h = h
        self._view = None

    def close(self):
        if getattr(self, "h", None):
            lib().wc_loop_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self, stream):
        import torch
        return torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream

    def reset(self, q0, stream=None):
        _check(lib().wc_loop_reset_dev(self.h, _arg(q0, "f64", self.nJ * self.B, "cuda", "q0"), self._stream(stream)), "wc_loop_reset_dev")

    def set_push(self, push_tick, push_dxy, stream=None):
        _check(lib().wc_loop_set_push_dev(self.h, _arg(push_tick, "i32", self.B, "cuda", "push_tick"),
                                          _arg(push_dxy, "f64", 2 * self.B, "cuda", "push_dxy"), self._stream(stream)),
               "wc_loop_set_push_dev")

    def run(self, n_ticks=1, stream=None):
        _check(lib().wc_loop_run_dev(self.h, int(n_ticks), self._stream(stream)), "wc_loop_run_dev")

    @property
    def tick(self):
        return lib().wc_loop_tick_count(self.h)

    def view(self):
        """dict of torch tensors aliasing the handle's device buffers (signals of the last tick run; no copy)"""
        import torch
        if self._view is None:
            v = LoopView()
            _check(lib().wc_loop_view_get(self.h, C.byref(v)), "wc_loop_view_get")
            out = {}
            for k, rows, ts in LOOP_VIEW:
                r = self.nJ if rows == "nJ" else (2 * (self.N + 1) if rows == "2N" else rows)
                shape = (self.B,) if ts in ("i4", "u1") else (r, self.B)
                out[k] = torch.as_tensor(_DevArray(getattr(v, k), shape, ts), device=f"cuda:{self.device}")
            self._view = out
        return self._view


def host_tick_arena_in(B, N, nJ, window=False, base=False, reset=False):
    """Pinned host buffers for one tick's inputs of wc_tick_host_async, carved out of ONE arena in the packed order the C ABI
    documents, so that the upload is two copies (the doubles, the byte flags) instead of sixteen.  Returns a dict of views."""
    import torch
    rows = [("lf_H", 12), ("rf_H", 12), ("dcm_meas", 2), ("ref_newest", 2), ("q_meas", nJ), ("q_lin", nJ), ("lf_twist", 6),
            ("rf_twist", 6), ("com_pos_des", 3), ("com_vel_des", 3), ("com_meas", 3), ("neck_R_des", 9)]
    if window:  # the whole window instead of its newest sample: ref_newest stays NULL, ref_window follows the run
        rows = [r for r in rows if r[0] != "ref_newest"]
    tail = ([("base_H", 12)] if base else []) + ([("ref_window", 2 * (N + 1))] if window else [])
    total = sum(r for _, r in rows + tail)
    arena = torch.empty(total, B, dtype=torch.float64).pin_memory()
    out, k = {}, 0
    for name, r in rows + tail:
        out[name] = arena[k:k + r]; k += r
    Bp = (B + 7) // 8 * 8
    flags = torch.zeros(4 if reset else 3, Bp, dtype=torch.uint8).pin_memory()
    for j, name in enumerate(("l_contact", "r_contact", "left_is_fixed") + (("reset",) if reset else ())):
        out[name] = flags[j, :B]
    out["_arena"] = (arena, flags)
    return out


def host_tick_arena_out(B, nJ):
    """Pinned host buffers for one tick's results of wc_tick_host_async in ONE arena (one download)."""
    import torch
    nd = (2 + nJ + 12) * B
    ni = (B * 4 + 7) // 8 * 8
    buf = torch.empty(nd * 8 + 5 * ni, dtype=torch.uint8).pin_memory()
    dbl = buf[:nd * 8].view(torch.float64).view(2 + nJ + 12, B)
    out = dict(zmp=dbl[0:2], dq=dbl[2:2 + nJ], lf_err=dbl[2 + nJ:8 + nJ], rf_err=dbl[8 + nJ:14 + nJ])
    for j, name in enumerate(("osqp_status", "iters", "flags", "qp_status", "nwsr")):
        out[name] = buf[nd * 8 + j * ni: nd * 8 + j * ni + B * 4].view(torch.int32)
    out["_arena"] = buf
    return out


COMM_ID_BYTES = 128


def comm_unique_id():
    """bytes of a fresh NCCL unique id (rank 0 creates it and hands it to the other ranks)"""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    _check(lib().wc_comm_unique_id(buf), "wc_comm_unique_id")
    return bytes(buf)


class Comm:
    """wc_comm_*: NCCL result collection (all-gather of per-instance results + all-reduce of statistics) on a side stream"""

    def __init__(self, uid, world, rank, device=0):
        self.world, self.rank, self.device = world, rank, device
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(uid)
        h = _VP()
        _check(lib().wc_comm_create(buf, world, rank, device, C.byref(h)), "wc_comm_create")
        self.h = h
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            lib().wc_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def gather(self, B, n_dofs, zmp=None, dq=None, osqp_status=None, iters=None, qp_status=None, nwsr=None, stats_sum=None,
               stats_max=None, stream=None, wait="stream"):
        """all-gather the given CUDA tensors ([rows, B] blocks -> [world, rows, B]) and all-reduce stats_sum / stats_max in
        place, after the work on `stream`; wait = "stream" (stream waits), "host" (blocks) or None (call wait() later)"""
        import torch
        dev = torch.device(f"cuda:{self.device}")
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        io = GatherIO(); out = {}
        spec = dict(zmp=(zmp, 2, "f64", torch.float64), dq=(dq, n_dofs, "f64", torch.float64), osqp_status=(osqp_status, 1, "i32", torch.int32),
                    iters=(iters, 1, "i32", torch.int32), qp_status=(qp_status, 1, "i32", torch.int32), nwsr=(nwsr, 1, "i32", torch.int32))
        for k, (t, rows, kind, dt) in spec.items():
            if t is None:
                continue
            shape = (self.world, B) if rows == 1 else (self.world, rows, B)
            out[k] = torch.empty(shape, dtype=dt, device=dev)
            setattr(io, k, _arg(t, kind, rows * B, "cuda", k)); setattr(io, k + "_all", out[k].data_ptr())
        if stats_sum is not None:
            io.stats_sum = _arg(stats_sum, "f64", stats_sum.numel(), "cuda", "stats_sum"); io.n_stats_sum = stats_sum.numel()
        if stats_max is not None:
            io.stats_max = _arg(stats_max, "f64", stats_max.numel(), "cuda", "stats_max"); io.n_stats_max = stats_max.numel()
        self._keep = (zmp, dq, osqp_status, iters, qp_status, nwsr, stats_sum, stats_max, out)
        _check(lib().wc_gather_async(self.h, C.byref(io), B, n_dofs, stream), "wc_gather_async")
        if wait == "stream":
            _check(lib().wc_gather_wait(self.h, stream, 0), "wc_gather_wait")
        elif wait == "host":
            _check(lib().wc_gather_wait(self.h, None, 1), "wc_gather_wait")
        return out

    def wait(self, stream=None):
        _check(lib().wc_gather_wait(self.h, stream, 1 if stream is None else 0), "wc_gather_wait")
