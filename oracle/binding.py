"""ctypes binding of oracle/libqmpc_oracle.so (TEST INFRASTRUCTURE ONLY).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm —
never by quaternion_mpc_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from quaternion_mpc_b200.abi import (FOOT_UPDATE_INPUT_DTYPE, FOOT_UPDATE_OUTPUT_DTYPE, CONVEX_PROBLEM_DTYPE, GAIT_STATE_DTYPE, GOAL_INPUT_DTYPE, QmpcRaibertParams, PROBLEM_DTYPE, QMPC_MAX_HORIZON,
                                     RESULT_DTYPE, WARM_DTYPE, QmpcConfig, QmpcLegParams)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("status", C.c_int), ("ls_trials", C.c_int),
                ("cost", C.c_double), ("max_violation", C.c_double), ("stationarity", C.c_double),
                ("penalty", C.c_double), ("pivot_ratio", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libqmpc_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        _LIB.qmpc_ref_solve_batch.argtypes = [C.POINTER(QmpcConfig), vp, C.c_int, vp, C.c_int]
        _LIB.qmpc_ref_solve_batch_convex.argtypes = [C.POINTER(QmpcConfig), vp, C.c_int, vp, C.c_int]
        _LIB.qmpc_ref_solve_batch_sched.argtypes = [C.POINTER(QmpcConfig), vp, vp, C.c_int, vp, C.c_int]
        _LIB.qmpc_ref_solve_batch_convex_sched.argtypes = [C.POINTER(QmpcConfig), vp, vp, C.c_int, vp, C.c_int]
        _LIB.qmpc_ref_solve_batch_warm.argtypes = [C.POINTER(QmpcConfig), vp, vp, vp, C.c_int, vp, C.c_int]
        _LIB.qmpc_ref_goal_update.argtypes = [vp, vp, C.c_int, vp]
        _LIB.qmpc_ref_raibert_targets.argtypes = [C.POINTER(QmpcRaibertParams), vp, C.c_int, vp, vp]
        _LIB.qmpc_ref_predict_schedule.argtypes = [C.POINTER(QmpcConfig), vp, C.c_int, vp]
        _LIB.qmpc_ref_leg_kinematics.argtypes = [C.POINTER(QmpcLegParams), vp, C.c_int, vp, vp]
        _LIB.qmpc_ref_joint_torques.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
        _LIB.qmpc_ref_leg_fsm_init.argtypes = [vp, vp, C.c_double, C.c_int]
        _LIB.qmpc_ref_foot_update.argtypes = [vp, vp, C.c_double, C.c_double, C.c_int, vp]
        _LIB.kat_double_integrator.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, dp, dp, C.POINTER(Stats)]
        _LIB.kat_pendulum.argtypes = [C.c_int, dp, dp, C.POINTER(Stats)]
        _LIB.kat_pendulum_ls.argtypes = [C.c_int, C.c_int, dp, dp, C.POINTER(Stats)]
        _LIB.kat_double_integrator_ls.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp, C.POINTER(Stats)]
        _LIB.kat_quat_golden.argtypes = [C.c_int, dp, dp, C.POINTER(Stats)]
        _LIB.kat_quat_rollout.argtypes = [C.c_int, dp, dp]
        _LIB.kat_pendulum_midpoint.argtypes = [dp, dp, C.c_float, dp, dp]
        _LIB.kat_di_dynamics.argtypes = [dp, dp, C.c_float, dp, dp]
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def solve_batch(cfg, problems, nthreads=1):
    """QuatMpc::grf_update on every element of `problems` (PROBLEM_DTYPE array) -> RESULT_DTYPE array."""
    problems = np.ascontiguousarray(problems, dtype=PROBLEM_DTYPE)
    out = np.zeros(problems.shape[0], dtype=RESULT_DTYPE)
    rc = lib().qmpc_ref_solve_batch(C.byref(cfg), problems.ctypes.data, problems.shape[0], out.ctypes.data, nthreads)
    if rc:
        raise RuntimeError(f"oracle failed rc={rc}")
    return out


def solve_batch_convex(cfg, problems, nthreads=1):
    problems = np.ascontiguousarray(problems, dtype=CONVEX_PROBLEM_DTYPE)
    out = np.zeros(problems.shape[0], dtype=RESULT_DTYPE)
    rc = lib().qmpc_ref_solve_batch_convex(C.byref(cfg), problems.ctypes.data, problems.shape[0], out.ctypes.data, nthreads)
    if rc:
        raise RuntimeError(f"oracle failed rc={rc}")
    return out


def solve_batch_diag(cfg, problems, warm=None, schedule=None, nthreads=1):
    """solve_batch / solve_batch_warm plus a conditioning diagnostic per solve: the largest ratio of the largest to
    the smallest Cholesky pivot of any Quu factored during the solve (an estimate of cond(Quu)).  Used by the parity
    policy: a solve whose ratio exceeds 1e12 amplifies 1-ulp differences between two fp64 implementations past the
    1e-4 N tolerance - it is numerically undetermined, and is reported as such instead of being hidden."""
    problems = np.ascontiguousarray(problems, dtype=PROBLEM_DTYPE)
    sched = None if schedule is None else _sched_bytes(schedule, problems.shape[0])
    out = np.zeros(problems.shape[0], dtype=RESULT_DTYPE)
    ratio = np.zeros(problems.shape[0])
    f = lib().qmpc_ref_solve_batch_diag
    f.argtypes = [C.POINTER(QmpcConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    rc = f(C.byref(cfg), problems.ctypes.data, sched.ctypes.data if sched is not None else None,
           warm.ctypes.data if warm is not None else None, problems.shape[0], out.ctypes.data, ratio.ctypes.data, nthreads)
    if rc:
        raise RuntimeError(f"oracle failed rc={rc}")
    return out, ratio


def _sched_bytes(schedule, batch):
    from quaternion_mpc_b200.abi import QMPC_MAX_HORIZON
    schedule = np.ascontiguousarray(schedule, dtype=np.uint8)
    assert schedule.shape == (batch, QMPC_MAX_HORIZON), schedule.shape
    return schedule


def solve_batch_sched(cfg, problems, schedule, nthreads=1):
    """Per-step contact-schedule extension: schedule[b, k] = contact bit mask of knot k."""
    problems = np.ascontiguousarray(problems, dtype=PROBLEM_DTYPE)
    schedule = _sched_bytes(schedule, problems.shape[0])
    out = np.zeros(problems.shape[0], dtype=RESULT_DTYPE)
    rc = lib().qmpc_ref_solve_batch_sched(C.byref(cfg), problems.ctypes.data, schedule.ctypes.data,
                                          problems.shape[0], out.ctypes.data, nthreads)
    if rc:
        raise RuntimeError(f"oracle failed rc={rc}")
    return out


def solve_batch_warm(cfg, problems, warm, schedule=None, nthreads=1):
    """Warm-started solve (row N4): `warm` (WARM_DTYPE array) is read and updated IN PLACE."""
    problems = np.ascontiguousarray(problems, dtype=PROBLEM_DTYPE)
    assert warm.dtype == WARM_DTYPE and warm.flags.c_contiguous and warm.shape[0] == problems.shape[0]
    sched = None if schedule is None else _sched_bytes(schedule, problems.shape[0])
    out = np.zeros(problems.shape[0], dtype=RESULT_DTYPE)
    rc = lib().qmpc_ref_solve_batch_warm(C.byref(cfg), problems.ctypes.data, sched.ctypes.data if sched is not None else None,
                                         warm.ctypes.data, problems.shape[0], out.ctypes.data, nthreads)
    if rc:
        raise RuntimeError(f"oracle failed rc={rc}")
    return out


def solve_batch_convex_sched(cfg, problems, schedule, nthreads=1):
    problems = np.ascontiguousarray(problems, dtype=CONVEX_PROBLEM_DTYPE)
    schedule = _sched_bytes(schedule, problems.shape[0])
    out = np.zeros(problems.shape[0], dtype=RESULT_DTYPE)
    rc = lib().qmpc_ref_solve_batch_convex_sched(C.byref(cfg), problems.ctypes.data, schedule.ctypes.data,
                                                 problems.shape[0], out.ctypes.data, nthreads)
    if rc:
        raise RuntimeError(f"oracle failed rc={rc}")
    return out


def new_goal_state(batch):
    """Zero-filled per-robot state of QuatMpc::goal_update (filters + desired position), opaque bytes."""
    return np.zeros((batch, lib().qmpc_ref_goal_state_bytes()), dtype=np.uint8)


def goal_update(state, goal_inputs, problems):
    """One QuatMpc::goal_update tick; `state` (from new_goal_state) and `problems` are updated IN PLACE."""
    g = np.ascontiguousarray(goal_inputs, dtype=GOAL_INPUT_DTYPE)
    assert problems.dtype == PROBLEM_DTYPE and problems.flags.c_contiguous and state.flags.c_contiguous
    rc = lib().qmpc_ref_goal_update(state.ctypes.data, g.ctypes.data, g.shape[0], problems.ctypes.data)
    assert rc == 0
    return problems


def raibert_targets(rp, goal_inputs):
    g = np.ascontiguousarray(goal_inputs, dtype=GOAL_INPUT_DTYPE)
    tw, tr = np.zeros((g.shape[0], 12)), np.zeros((g.shape[0], 12))
    rc = lib().qmpc_ref_raibert_targets(C.byref(rp), g.ctypes.data, g.shape[0], tw.ctypes.data, tr.ctypes.data)
    assert rc == 0
    return tw, tr


def new_leg_fsm(batch, gait=None, gait_freq=2.2):
    """Per-robot state of the four LeggedContactFSM objects after reset_params + reset (opaque bytes)."""
    st = np.zeros((batch, lib().qmpc_ref_leg_fsm_state_bytes()), dtype=np.uint8)
    g = None if gait is None else np.ascontiguousarray(gait, dtype=np.int32)
    assert lib().qmpc_ref_leg_fsm_init(st.ctypes.data, g.ctypes.data if g is not None else None, gait_freq, batch) == 0
    return st


def foot_update(state, inputs, dt=5.0 / 1000.0, gait_freq=2.2):
    """One QuatMpc::foot_update tick; `state` (from new_leg_fsm) is updated IN PLACE."""
    i = np.ascontiguousarray(inputs, dtype=FOOT_UPDATE_INPUT_DTYPE)
    out = np.zeros(i.shape[0], dtype=FOOT_UPDATE_OUTPUT_DTYPE)
    assert lib().qmpc_ref_foot_update(state.ctypes.data, i.ctypes.data, dt, gait_freq, i.shape[0], out.ctypes.data) == 0
    return out


def predict_schedule(cfg, gait_states):
    """LeggedContactFSM::predict_contact_state at t + k*dt, k < horizon -> (batch, QMPC_MAX_HORIZON) uint8."""
    g = np.ascontiguousarray(gait_states, dtype=GAIT_STATE_DTYPE)
    out = np.zeros((g.shape[0], QMPC_MAX_HORIZON), dtype=np.uint8)
    rc = lib().qmpc_ref_predict_schedule(C.byref(cfg), g.ctypes.data, g.shape[0], out.ctypes.data)
    assert rc == 0
    return out


def leg_kinematics(leg_params, joint_pos):
    """a1_kin.fk / a1_kin.jac for all legs -> (foot_pos_body (batch,12), jac_foot (batch,36))."""
    q = np.ascontiguousarray(joint_pos, dtype=np.float64)
    foot, jac = np.zeros((q.shape[0], 12)), np.zeros((q.shape[0], 36))
    rc = lib().qmpc_ref_leg_kinematics(C.byref(leg_params), q.ctypes.data, q.shape[0], foot.ctypes.data, jac.ctypes.data)
    assert rc == 0
    return foot, jac


def joint_torques(results, jac_foot, plan_contacts, movement_mode):
    results = np.ascontiguousarray(results, dtype=RESULT_DTYPE)
    jac = np.ascontiguousarray(jac_foot, dtype=np.float64)
    pc = None if plan_contacts is None else np.ascontiguousarray(plan_contacts, dtype=np.int32)
    tau = np.zeros((results.shape[0], 12))
    rc = lib().qmpc_ref_joint_torques(results.ctypes.data, jac.ctypes.data, pc.ctypes.data if pc is not None else None,
                                      int(movement_mode), results.shape[0], tau.ctypes.data)
    assert rc == 0
    return tau


def kat_double_integrator(variant, penalty_initial=0.0, penalty_scaling=0.0, iterations_max=0, cubic=False):
    """variant 0..2: TestDoubleIntegrator.cpp:69-375; 3: the second-order-cone control bound (:377-491).
    cubic: ALTRO's default (strong-Wolfe, cubic interpolation) line search instead of back-tracking."""
    X, U, st = np.zeros((11, 4)), np.zeros((10, 2)), Stats()
    lib().kat_double_integrator_ls(variant, penalty_initial, penalty_scaling, iterations_max, int(cubic), _dp(X), _dp(U),
                                   C.byref(st))
    return X, U, st


def soc_project(z):
    """Projection of z = (v, s) onto the second-order cone |v| <= s and its Jacobian (oracle/altro_ref.c)."""
    z = np.ascontiguousarray(z, dtype=np.float64)
    p = z.shape[0]
    out, J = np.zeros(p), np.zeros((p, p))
    fn = lib().altro_ref_soc_project
    fn.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    fn.restype = None
    fn(_dp(z), p, _dp(out), _dp(J))
    return out, J


def kat_pendulum(variant, cubic=False):
    N = 50 if variant == 0 else 20
    X, U, st = np.zeros((N + 1, 2)), np.zeros((N, 1)), Stats()
    lib().kat_pendulum_ls(variant, int(cubic), _dp(X), _dp(U), C.byref(st))
    return X, U, st


def kat_quat_golden(which):
    m = 12 if which == 0 else 6
    X, U, st = np.zeros((21, 13)), np.zeros((20, m)), Stats()
    lib().kat_quat_golden(which, _dp(X), _dp(U), C.byref(st))
    return X, U, st


def kat_quat_rollout(which, U):
    U = np.ascontiguousarray(U, dtype=np.float64)
    X = np.zeros((21, 13))
    lib().kat_quat_rollout(which, _dp(U), _dp(X))
    return X


def kat_pendulum_midpoint(x, u, h):
    x, u = np.array(x, float), np.array(u, float)
    xn, J = np.zeros(2), np.zeros(6)
    lib().kat_pendulum_midpoint(_dp(x), _dp(u), h, _dp(xn), _dp(J))
    return xn, J.reshape(3, 2).T  # column-major 2x3


def kat_di_dynamics(x, u, h):
    x, u = np.array(x, float), np.array(u, float)
    xn, J = np.zeros(4), np.zeros(24)
    lib().kat_di_dynamics(_dp(x), _dp(u), h, _dp(xn), _dp(J))
    return xn, J.reshape(6, 4).T
