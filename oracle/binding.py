"""ctypes binding of oracle/libqmpc_oracle.so (TEST INFRASTRUCTURE ONLY).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm —
never by quaternion_mpc_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from quaternion_mpc_b200.abi import (CONVEX_PROBLEM_DTYPE, PROBLEM_DTYPE, RESULT_DTYPE, QmpcConfig)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("status", C.c_int), ("ls_trials", C.c_int),
                ("cost", C.c_double), ("max_violation", C.c_double), ("stationarity", C.c_double),
                ("penalty", C.c_double)]


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
        _LIB.kat_double_integrator.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, dp, dp, C.POINTER(Stats)]
        _LIB.kat_pendulum.argtypes = [C.c_int, dp, dp, C.POINTER(Stats)]
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


def kat_double_integrator(variant, penalty_initial=0.0, penalty_scaling=0.0, iterations_max=0):
    X, U, st = np.zeros((11, 4)), np.zeros((10, 2)), Stats()
    lib().kat_double_integrator(variant, penalty_initial, penalty_scaling, iterations_max, _dp(X), _dp(U), C.byref(st))
    return X, U, st


def kat_pendulum(variant):
    N = 50 if variant == 0 else 20
    X, U, st = np.zeros((N + 1, 2)), np.zeros((N, 1)), Stats()
    lib().kat_pendulum(variant, _dp(X), _dp(U), C.byref(st))
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
