"""The C++ host shim (CudaQuatMpc : LeggedMpc) compiled against stub reference headers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from quaternion_mpc_b200 import abi
from quaternion_mpc_b200.config import default_config
from quaternion_mpc_b200.workloads import random_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "stubs", "libshim_test.so")


@pytest.fixture(scope="module")
def shim():
    import __graft_entry__ as g
    g.build()
    pkg = os.path.join(ROOT, "quaternion_mpc_b200")
    srcs = [os.path.join(pkg, "shim", "CudaQuatMpc.cpp"), os.path.join(pkg, "shim", "CudaConvexMpc.cpp"),
            os.path.join(ROOT, "tests", "stubs", "shim_driver.cpp")]
    deps = srcs + [os.path.join(pkg, "shim", "CudaQuatMpc.h"), os.path.join(pkg, "shim", "CudaConvexMpc.h"),
                   os.path.join(ROOT, "tests", "stubs", "mpc", "LeggedMpc.h"), os.path.join(ROOT, "include", "qmpc.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO] + srcs + [
            "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
            "-I", os.path.join(pkg, "shim"), "-L", pkg, "-lqmpc_b200", f"-Wl,-rpath,{pkg}"])
    return C.CDLL(SO)


def test_shim_builds_and_fails_loudly_without_gpu(shim):
    import torch
    err = C.create_string_buffer(256)
    rc = shim.shim_construct_only(0, err, 256)
    if torch.cuda.is_available():
        assert rc == 0
    else:
        assert rc == 1 and b"qmpc_create failed" in err.value


def test_convex_shim_builds_and_fails_loudly_without_gpu(shim):
    import torch
    err = C.create_string_buffer(256)
    rc = shim.shim_convex_construct_only(0, err, 256)
    if torch.cuda.is_available():
        assert rc == 0
    else:
        assert rc == 1 and b"CudaConvexMpc: qmpc_create failed" in err.value


@pytest.mark.gpu
@pytest.mark.parametrize("walking", [0, 1])
def test_convex_shim_tick_matches_oracle(shim, oracle, walking):
    """CudaConvexMpc::update through the LeggedMpc pointer: the problem it packs solves to what it wrote back;
    walking = the per-knot schedule option (the TODO at ConvexMpc.cpp:82) through the FSM stub."""
    from quaternion_mpc_b200.workloads import random_convex_batch
    p = random_convex_batch(1, seed=31, gait="stand")
    out_p = np.zeros(1, dtype=abi.CONVEX_PROBLEM_DTYPE)
    sched = np.zeros((1, abi.QMPC_MAX_HORIZON), dtype=np.uint8)
    gb, st6, si = np.zeros(12), np.zeros(6), np.zeros(2, dtype=np.int32)
    shim.shim_convex_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    rc = shim.shim_convex_run(p.ctypes.data, 10, 3, walking, 0.43, out_p.ctypes.data, sched.ctypes.data,
                              gb.ctypes.data, st6.ctypes.data, si.ctypes.data)
    assert rc == 0
    assert np.allclose(out_p["torso_euler"], p["torso_euler"]) and np.allclose(out_p["foot_pos_abs_com"], p["foot_pos_abs_com"])
    assert abs(out_p["torso_lin_vel_d_world"][0, 0] - 3 * 0.005) < 1e-12      # 1 m/s^2 ramp, three 5 ms ticks
    assert out_p["torso_lin_vel_d_world"][0, 1] == -0.05 and out_p["yaw_rate_d"][0] == 0.2
    assert out_p["torso_pos_d_world"][0, 2] == 0.29
    assert np.allclose(st6[:3], out_p["torso_pos_d_world"][0]) and np.allclose(st6[3:], [0.01, 0.02, 0.03])
    cfg = default_config(2, 10)
    if walking:
        # FSM stub: trot from phase 0.43 at 2.2 Hz, 5 ms knots: FL+RR (1001b) until the phase passes 0.5, then FR+RL
        assert sched[0, :10].tolist() == [9] * 7 + [6] * 3 and (sched[0, 10:] == 0).all()
        ref = oracle.solve_batch_convex_sched(cfg, out_p, sched)
    else:
        assert (out_p["plan_contacts"] == 1).all()
        ref = oracle.solve_batch_convex(cfg, out_p)
    assert np.abs(ref["grf_body"][0] - gb).max() < 1e-4
    assert si[1] == ref["iterations"][0]


@pytest.mark.gpu
def test_shim_tick_matches_oracle(shim, oracle):
    p = random_batch(1, seed=21, gait="stand")
    out_p = np.zeros(1, dtype=abi.PROBLEM_DTYPE)
    gb, gw, si = np.zeros(12), np.zeros(12), np.zeros(2, dtype=np.int32)
    rc = shim.shim_run(C.c_void_p(p.ctypes.data), 10, 3, C.c_void_p(out_p.ctypes.data),
                       C.c_void_p(gb.ctypes.data), C.c_void_p(gw.ctypes.data), C.c_void_p(si.ctypes.data))
    assert rc == 0
    # the shim packed: measured state from the input, desired quantities from its own filters
    assert np.allclose(out_p["torso_quat"], p["torso_quat"])
    assert (out_p["plan_contacts"] == 1).all()
    assert abs(out_p["torso_lin_vel_d_body"][0, 0]) > 0          # joystick velocity went through the filter
    ref = oracle.solve_batch(default_config(0, 10), out_p)
    assert np.abs(ref["grf_body"][0] - gb).max() < 1e-4
    assert np.abs(ref["grf_world"][0] - gw).max() < 1e-4
    assert si[1] == ref["iterations"][0]


@pytest.mark.gpu
def test_shim_contact_schedule_mode(shim, oracle):
    """enable_contact_schedule(true): the shim asks the (stub of the) reference's own
    LeggedContactFSM::predict_contact_state for every knot and solves through the schedule entry point."""
    p = random_batch(1, seed=22, gait="stand")
    out_p = np.zeros(1, dtype=abi.PROBLEM_DTYPE)
    sched = np.zeros((1, abi.QMPC_MAX_HORIZON), dtype=np.uint8)
    gb, gw, si = np.zeros(12), np.zeros(12), np.zeros(2, dtype=np.int32)
    shim.shim_run_sched.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
    rc = shim.shim_run_sched(p.ctypes.data, 10, 2, 0.43, out_p.ctypes.data, sched.ctypes.data, gb.ctypes.data,
                             gw.ctypes.data, si.ctypes.data)
    assert rc == 0
    # trot from phase 0.43 at 2.2 Hz, 10 ms knots: FL+RR (1001b) until the phase passes 0.5, then FR+RL (0110b)
    assert sched[0, :4].tolist() == [9, 9, 9, 9] and sched[0, 4:10].tolist() == [6] * 6 and (sched[0, 10:] == 0).all()
    ref = oracle.solve_batch_sched(default_config(0, 10), out_p, sched)
    assert np.abs(ref["grf_body"][0] - gb).max() < 1e-4 and si[1] == ref["iterations"][0]


@pytest.mark.gpu
@pytest.mark.parametrize("policy", [0, 1, 2])
def test_shim_failed_solve_is_reported_and_follows_the_policy(shim, policy):
    """A solve the solver cannot do (NaN state -> status NONFINITE) is counted, flagged and described; update() still
    returns true (the reference's contract) and the GRFs follow the explicit policy: hold the last tick's values,
    the weight share u_ref, or zero."""
    p = random_batch(1, seed=23, gait="stand")
    g1, g2, info = np.zeros(12), np.zeros(12), np.zeros(4, dtype=np.int32)
    err = C.create_string_buffer(256)
    shim.shim_failure_policy.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
    assert shim.shim_failure_policy(p.ctypes.data, policy, g1.ctypes.data, g2.ctypes.data, info.ctypes.data, err, 256) == 0
    assert info.tolist() == [1, 1, 4, 1]
    assert b"nonfinite" in err.value
    assert np.isfinite(g1).all() and g1[2::3].sum() > 40.0        # a real solve: the feet carry the robot
    if policy == 0:
        assert np.array_equal(g2, g1)
    elif policy == 1:
        want = np.zeros(12); want[2::3] = 12.84 * 9.81 / 4
        assert np.allclose(g2, want, atol=1e-12)
    else:
        assert (g2 == 0).all()


@pytest.mark.gpu
def test_shim_sine_attitude_trajectory(shim):
    """joy.sin_ang_vel (QuatMpc.cpp:139-146): the desired attitude is euler_to_quat of the sine Euler trajectory,
    not the integrated joystick rate."""
    p = random_batch(1, seed=24, gait="stand")
    packed, state_q, eul = np.zeros(4), np.zeros(4), np.zeros(3)
    ticks = 37
    assert shim.shim_sin_ang_vel(C.c_void_p(p.ctypes.data), ticks, C.c_void_p(packed.ctypes.data), C.c_void_p(state_q.ctypes.data),
                                 C.c_void_p(eul.ctypes.data)) == 0
    e = 3.14 / 8 * np.sin(2 * 3.14 / 900 * (ticks - 1))
    assert np.allclose(eul, e, atol=1e-15)
    c, s_ = np.cos(e / 2), np.sin(e / 2)
    want = np.array([c * c * c + s_ * s_ * s_, c * c * s_ - s_ * s_ * c, c * s_ * c + s_ * c * s_, s_ * c * c - c * s_ * s_])   # Utils.cpp:94-97
    assert np.abs(packed - want).max() < 1e-15 and np.abs(state_q - want).max() < 1e-12
