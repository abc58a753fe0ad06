"""C++ drop-in host classes (walking-controllers_b200/host): parameter parsing of the reference's shipped
.ini files (CPU) and the per-tick call sequence of WalkingModule::updateModule on the GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "walking-controllers_b200", "host")
REF_INI = "/root/reference/app/robots/iCubGazeboV2_5"
FROZEN = os.path.join(ROOT, "tests", "golden", "iCubGazeboV2_5_params.txt")


def _parse(text):
    out = {}
    for line in text.strip().splitlines():
        p = line.split()
        out[p[0]] = [float(v) for v in p[1:]]
    return out


def _build(target):
    subprocess.check_call(["make", "-C", HOST, target], stdout=subprocess.DEVNULL)


def test_shipped_ini_values_match_defaults(wc):
    """tests/golden/iCubGazeboV2_5_params.txt is ini_dump's reading of the reference's own
    controllerParams.ini / qpInverseKinematics.ini / dcmWalkingCoordinator.ini (regenerated and re-checked
    whenever /root/reference is present); the library defaults must equal the shipped values."""
    if os.path.isdir(REF_INI):
        _build("ini_dump")
        txt = subprocess.check_output([os.path.join(HOST, "ini_dump")] + [os.path.join(REF_INI, f) for f in
                                      ("controllerParams.ini", "qpInverseKinematics.ini", "dcmWalkingCoordinator.ini")], text=True)
        assert _parse(txt) == _parse(open(FROZEN).read()), "frozen parameter fixture is stale"
    p = _parse(open(FROZEN).read())
    cfg = wc.mpc_default_config(200)
    assert p["sampling_time"] == [cfg.sampling_time] and p["com_height"] == [cfg.com_height]
    assert round(p["controllerHorizon"][0] / p["sampling_time"][0]) == 200          # shipped horizon: 2 s / 0.01 s
    assert p["stateWeightTriplets"] == [0, 0, cfg.Q[0], 1, 1, cfg.Q[3]]
    assert p["inputWeightTriplets"] == [0, 0, cfg.R[0], 1, 1, cfg.R[3]]
    assert p["foot_size"] == list(cfg.foot_x) + list(cfg.foot_y)
    assert p["initial_zmp_position"] == list(cfg.initial_zmp) and p["convex_hull_tolerance"] == [cfg.hull_tolerance]
    ik = wc.ik_default_config()
    assert p["useCoMAsConstraint"] == [ik.use_com_as_constraint]
    assert p["neckWeightTriplets"] == [0, 0, ik.neck_weight[0], 1, 1, ik.neck_weight[4], 2, 2, ik.neck_weight[8]]
    assert p["additional_rotation"] == list(ik.additional_rotation)
    assert np.allclose(np.deg2rad(p["jointRegularization"]), list(ik.joint_reg_rad)[:23], rtol=0, atol=1e-15)
    assert p["jointRegularizationWeights"] == list(ik.joint_reg_weight)[:23]
    assert p["jointRegularizationGains"] == list(ik.joint_reg_gain)[:23]
    assert [p["k_posCom"][0], p["k_posFoot"][0], p["k_attFoot"][0], p["k_neck"][0]] == [ik.k_pos_com, ik.k_pos_foot, ik.k_att_foot, ik.k_neck]


def test_host_classes_compile(wc, oracle):
    wc.build(); oracle.build()
    _build("all")
    assert os.path.exists(os.path.join(HOST, "test_dropin"))


@pytest.mark.gpu
@pytest.mark.parametrize("horizon,ticks", [(25, 300), (200, 60)])
def test_dropin_classes_follow_the_reference_call_sequence(cuda_device, horizon, ticks):
    """WalkingController / WalkingFK / WalkingQPIK_qpOASES driven as in WalkingModule::updateModule (B = 1),
    checked against the oracle inside the binary (rel 1e-6 on the ZMP, 1e-5 rad/s on dq, same status)."""
    exe = os.path.join(HOST, "test_dropin")
    if not os.path.exists(exe):
        _build("all")
    out = subprocess.run([exe, str(horizon), str(ticks)], capture_output=True, text=True, timeout=600)
    print(out.stdout, out.stderr[-2000:])
    assert out.returncode == 0, out.stdout + out.stderr[-2000:]


def test_real_idyntree_binding_compiles(wc):
    """`make check_real_api`: every drop-in class and a caller of the reference's call sequence compiled with
    -DWC_USE_REAL_IDYNTREE against host/mock_include, a header set with ONLY genuine iDynTree 0.11 / YARP 3.1 members
    (WM/include/WalkingQPInverseKinematics.hpp:91-207 for the signatures) -- the binding INTEGRATION.md describes."""
    wc.build()
    subprocess.check_call(["make", "-C", HOST, "check_real_api"], stdout=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(HOST, "real_api_caller"))
    # the mock must not leak the stand-in API: no member of that name anywhere in it
    for root, _, files in os.walk(os.path.join(HOST, "mock_include")):
        for f in files:
            if f.endswith(".h"):
                txt = open(os.path.join(root, f)).read()
                for banned in ("toH12", "fromH12", "fromIniFile", " list(", " number("):
                    assert banned not in txt, (f, banned)


@pytest.mark.gpu
def test_real_api_caller_runs_one_tick(cuda_device):
    """the caller written against genuine iDynTree / YARP signatures runs one MPC + QP-IK + reactive-controller tick on the GPU"""
    exe = os.path.join(HOST, "real_api_caller")
    if not os.path.exists(exe):
        _build("check_real_api")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr[-2000:])
    assert out.returncode == 0, out.stdout + out.stderr[-2000:]
