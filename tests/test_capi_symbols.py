"""The C-ABI library loads and exports every entry point include/wc_capi.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "wc_capi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(wc_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("wc_mpc_create", "wc_mpc_tick_dev", "wc_mpc_tick_host", "wc_ik_create", "wc_ik_tick_dev",
                 "wc_ik_tick_host", "wc_fk_dev", "wc_ik_solve_with_jacobians_dev", "wc_mpc_reset"):
        assert must in names


def test_library_exports_every_declared_symbol(wc):
    wc.build()
    lib = C.CDLL(wc.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in wc_capi.h but not exported: {missing}"


def test_struct_layouts_and_defaults(wc, oracle):
    assert C.sizeof(wc.MpcConfig) == C.sizeof(oracle.MpcConfig)
    assert C.sizeof(wc.IkConfig) == C.sizeof(oracle.IkConfig)
    assert C.sizeof(wc.Model) == C.sizeof(oracle.Model)
    cfg = wc.mpc_default_config(25)
    # app/robots/iCubGazeboV2_5/controllerParams.ini + dcmWalkingCoordinator.ini
    assert cfg.horizon == 25 and cfg.sampling_time == 0.01 and cfg.com_height == 0.53
    assert list(cfg.Q) == [7500.0, 0.0, 0.0, 7500.0] and list(cfg.R) == [9e6, 0.0, 0.0, 9e6]
    assert list(cfg.foot_x) == [-0.02, 0.05] and list(cfg.foot_y) == [-0.025, 0.025]
    assert cfg.hull_tolerance == 0.05
    s = cfg.osqp  # osqp 0.6 defaults
    assert (s.rho, s.sigma, s.alpha, s.eps_abs, s.eps_rel) == (0.1, 1e-6, 1.6, 1e-3, 1e-3)
    assert (s.max_iter, s.scaling, s.check_termination, s.adaptive_rho, s.warm_start) == (4000, 10, 25, 1, 1)
    ik = wc.ik_default_config()
    assert ik.n_dofs == 23 and ik.use_com_as_constraint == 1 and ik.nwsr == 100
    assert (ik.k_pos_com, ik.k_pos_foot, ik.k_att_foot, ik.k_neck) == (1.0, 4.0, 2.0, 1.0)
    m = wc.model_icub_like()
    assert m.n_dofs == 23 and m.n_links == 25
    assert sorted(d for d in list(m.dof)[: m.n_links] if d >= 0) == list(range(23))


def test_product_has_no_cpu_fallback(wc):
    """compute entry points must fail loudly without a CUDA device instead of falling back"""
    import torch
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    cfg = wc.mpc_default_config(25)
    rc = wc.lib().wc_mpc_create(C.byref(cfg), 4, 0, C.byref(h))
    assert rc != 0 and not h.value


def test_sass_of_the_shipped_library(wc):
    """What DESIGN.md / profiles/r2d_sass.md claim about the machine code, checked on the library as built here (cuobjdump needs
    no GPU): sm_100a only, the thread-per-instance IK kernel fetches its model table with a bulk asynchronous copy (UBLKCP)
    completing on an mbarrier (SYNCS), and the benchmark kernels issue fp64 FMAs (DFMA), not tensor-core instructions."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    wc.build()
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "walking-controllers_b200", "libwc_b200.so")
    elfs = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in elfs and "sm_90" not in elfs and "sm_80" not in elfs
    names = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    fn = [l.split("Function : ")[1].strip() for l in names.splitlines() if "Function : " in l]
    tpi = [f for f in fn if "ik_tpi_kernelILi0E" in f]  # the whole tick in one launch (ILi1E / ILi2E: its two parts)
    split = [f for f in fn if "mpc_split_kernelILi0ELi25E" in f]
    assert len(tpi) == 1 and len(split) == 1
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", tpi[0], so], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass and "DFMA" in sass
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", split[0], so], capture_output=True, text=True).stdout
    assert sass.count("DFMA") > 1000 and "HMMA" not in sass and "UTCMMA" not in sass
    assert "LDGSTS" in sass  # the stage-in and the epilogue's scaling factors travel by cp.async
