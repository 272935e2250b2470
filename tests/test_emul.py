"""CPU-side check of the CUDA kernels' per-problem bodies (tests/emul): the very same QMPC_HD source
that the kernels run is compiled with g++ and compared with the oracle.  This is NOT the product
path (the product is libqmpc_b200.so on a GPU; the -m gpu tests go through its C-ABI) — it keeps the
kernel arithmetic covered in a container without a GPU.  Tolerance: the north_star's 1e-4 N on the
GRFs (achieved ~1e-8); integer outputs (contact schedules) must be bit-exact."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from quaternion_mpc_b200 import abi
from quaternion_mpc_b200.config import default_config
from quaternion_mpc_b200.workloads import (predict_schedule_numpy, random_batch, random_convex_batch,
                                           random_gait_states)

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-4


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libqmpc_emul.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DQMPC_EMUL_SRB", "-ffp-contract=off",
                           "-o", so, os.path.join(HERE, "emul", "emul.cpp")])
    lib = C.CDLL(so)
    for name in ("emul_solve_dense", "emul_solve_srb", "emul_solve_coop", "emul_solve_phased"):
        getattr(lib, name).argtypes = [C.POINTER(abi.QmpcConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p]
    lib.emul_predict_schedule.argtypes = [C.POINTER(abi.QmpcConfig), C.c_void_p, C.c_int, C.c_void_p]
    lib.emul_leg_kinematics.argtypes = [C.POINTER(abi.QmpcLegParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def _run(lib, fn, cfg, probs, sched=None, warm=None):
    out = np.zeros(len(probs), dtype=abi.RESULT_DTYPE)
    rc = getattr(lib, fn)(C.byref(cfg), probs.ctypes.data, sched.ctypes.data if sched is not None else None,
                          warm.ctypes.data if warm is not None else None, len(probs), out.ctypes.data)
    assert rc == 0
    return out


def _agree(res, ref, max_flagged=0.1):
    flagged = (res["status"] >= 2) | (ref["status"] >= 2)
    ok = ~flagged
    assert flagged.sum() <= max(1, int(len(res) * max_flagged))
    assert (res["status"][ok] == ref["status"][ok]).all()
    assert (res["iterations"][ok] == ref["iterations"][ok]).all()
    err = np.abs(res["grf_body"][ok] - ref["grf_body"][ok]).max()
    assert err < TOL, err
    return err


@pytest.mark.parametrize("kernel", ["emul_solve_coop", "emul_solve_srb", "emul_solve_dense"])
def test_kernel_bodies_match_oracle(emul, oracle, kernel):
    cfg = default_config(abi.QMPC_MODEL_QUAT_4FOOT, 10)
    p = random_batch(24, seed=0, gait="trot")
    _agree(_run(emul, kernel, cfg, p), oracle.solve_batch(cfg, p, nthreads=4))


@pytest.mark.parametrize("kernel", ["emul_solve_coop", "emul_solve_srb", "emul_solve_dense"])
def test_kernel_bodies_with_contact_schedule(emul, oracle, kernel):
    """Row N1: per-knot masks from the gait tables (trot / trot-with-stand / crawl)."""
    cfg = default_config(abi.QMPC_MODEL_QUAT_4FOOT, 10)
    p = random_batch(24, seed=1, gait="trot")
    sched = predict_schedule_numpy(random_gait_states(24, seed=1), 10, cfg.dt)
    assert len(np.unique(sched[:, :10])) > 3
    _agree(_run(emul, kernel, cfg, p, sched), oracle.solve_batch_sched(cfg, p, sched, nthreads=4))


def test_constant_schedule_is_bit_identical_to_plain(emul, oracle):
    cfg = default_config(abi.QMPC_MODEL_QUAT_4FOOT, 10)
    p = random_batch(16, seed=2, gait="mixed")
    m = (p["plan_contacts"] * np.array([1, 2, 4, 8])).sum(1).astype(np.uint8)
    sched = np.repeat(m[:, None], abi.QMPC_MAX_HORIZON, 1)
    for a, b in ((_run(emul, "emul_solve_coop", cfg, p, sched), _run(emul, "emul_solve_coop", cfg, p)),
                 (oracle.solve_batch_sched(cfg, p, sched), oracle.solve_batch(cfg, p))):
        assert np.array_equal(a["grf_body"], b["grf_body"]) and np.array_equal(a["iterations"], b["iterations"])


def test_convex_body_with_schedule(emul, oracle):
    cfg = default_config(abi.QMPC_MODEL_EULER_CONVEX, 10)
    p = random_convex_batch(12, seed=3)
    sched = predict_schedule_numpy(random_gait_states(12, seed=4), 10, cfg.dt)
    _agree(_run(emul, "emul_solve_dense", cfg, p, sched), oracle.solve_batch_convex_sched(cfg, p, sched, nthreads=4))
    _agree(_run(emul, "emul_solve_dense", cfg, p), oracle.solve_batch_convex(cfg, p, nthreads=4))


@pytest.mark.parametrize("kernel", ["emul_solve_coop", "emul_solve_srb", "emul_solve_dense"])
def test_warm_start_bodies(emul, oracle, kernel):
    """Row N4: two consecutive ticks with the trajectory-shift warm start; the buffer written by the
    kernel body must match the oracle's, and an invalid buffer must give the cold result bit for bit."""
    cfg = default_config(abi.QMPC_MODEL_QUAT_4FOOT, 10)
    p = random_batch(16, seed=8, gait="trot")
    w, wr = np.zeros(16, dtype=abi.WARM_DTYPE), np.zeros(16, dtype=abi.WARM_DTYPE)
    a0, r0 = _run(emul, kernel, cfg, p, None, w), oracle.solve_batch_warm(cfg, p, wr, nthreads=4)
    assert np.array_equal(a0["grf_body"], _run(emul, kernel, cfg, p)["grf_body"])
    _agree(a0, r0)
    assert (w["valid"] == 1).all() and np.abs(w["u"] - wr["u"]).max() < TOL
    assert np.abs(w["u"][:, 0, :] - a0["grf_body"]).max() == 0.0      # knot 0 of the buffer is the returned GRF
    a1, r1 = _run(emul, kernel, cfg, p, None, w), oracle.solve_batch_warm(cfg, p, wr, nthreads=4)
    _agree(a1, r1, max_flagged=0.3)   # warm-started iterates sit closer to the line-search floor
    assert np.abs(a1["grf_body"] - a0["grf_body"]).max() > 1e-6        # the warm start changed the iterate


@pytest.mark.parametrize("N,dt", [(10, 0.005), (20, 0.005), (30, 0.008)])
def test_convex_cooperative_body(emul, oracle, N, dt):
    """Row A8 on the cooperative design: ConvexMpc's Euler SRB through the same phase functions (state blocks
    swapped pairwise, fourth knot block Dw) against the oracle and the generic dense body, at the shipped
    horizons / steps (config/gazebo_go1_convex_mpc.yaml:36-37, hardware_go1_convex_mpc.yaml:36-37)."""
    cfg = default_config(abi.QMPC_MODEL_EULER_CONVEX, N)
    cfg.dt = dt
    p = random_convex_batch(16, seed=30 + N)
    ref = oracle.solve_batch_convex(cfg, p, nthreads=4)
    a = _run(emul, "emul_solve_coop", cfg, p)
    e = _agree(a, ref)
    b = _run(emul, "emul_solve_phased", cfg, p)
    assert a.tobytes() == b.tobytes()
    sched = predict_schedule_numpy(random_gait_states(16, seed=31), N, cfg.dt)
    _agree(_run(emul, "emul_solve_coop", cfg, p, sched), oracle.solve_batch_convex_sched(cfg, p, sched, nthreads=4))
    print(f"convex coop N={N}: max|dGRF| = {e:.2e}")


@pytest.mark.parametrize("model,N,gait", [(0, 10, "trot"), (0, 16, "mixed"), (1, 20, "stand"), (0, 1, "trot"), (0, 32, "trot")])
def test_phased_path_is_bit_identical_to_fused(emul, oracle, model, N, gait):
    """The split launches (set-up / backward / forward, solver state stored to and reloaded from the per-problem
    block between them) run the very phase functions of the fused kernel: identical results, bit for bit -
    including schedules, warm starts and a problem that stops early."""
    cfg = default_config(model, N)
    p = random_batch(24, seed=50 + N, gait=gait, **({"nfeet": 2, "max_angle": 0.2} if model == 1 else {}))
    p["torso_lin_vel_world"][3, 0] = np.nan       # non-finite: finishes in the set-up launch
    a, b = _run(emul, "emul_solve_coop", cfg, p), _run(emul, "emul_solve_phased", cfg, p)
    assert a.tobytes() == b.tobytes()
    sched = predict_schedule_numpy(random_gait_states(24, seed=51), N, cfg.dt)
    wa, wb = np.zeros(24, dtype=abi.WARM_DTYPE), np.zeros(24, dtype=abi.WARM_DTYPE)
    for tick in range(2):
        a, b = _run(emul, "emul_solve_coop", cfg, p, sched, wa), _run(emul, "emul_solve_phased", cfg, p, sched, wb)
        assert a.tobytes() == b.tobytes() and wa.tobytes() == wb.tobytes()
    cfg.iterations_max = 0
    assert _run(emul, "emul_solve_coop", cfg, p).tobytes() == _run(emul, "emul_solve_phased", cfg, p).tobytes()


def test_two_foot_model_body(emul, oracle):
    cfg = default_config(abi.QMPC_MODEL_QUAT_2FOOT, 12)
    p = random_batch(12, seed=5, nfeet=2, max_angle=0.2)
    _agree(_run(emul, "emul_solve_coop", cfg, p), oracle.solve_batch(cfg, p, nthreads=4))


def test_schedule_predictor_body_bit_exact(emul, oracle):
    cfg = default_config(abi.QMPC_MODEL_QUAT_4FOOT, 20)
    g = random_gait_states(4096, seed=6, gaits=(0, 1, 2, 3))
    out = np.zeros((len(g), abi.QMPC_MAX_HORIZON), np.uint8)
    assert emul.emul_predict_schedule(C.byref(cfg), g.ctypes.data, len(g), out.ctypes.data) == 0
    assert np.array_equal(out, oracle.predict_schedule(cfg, g))


def test_leg_kinematics_body(emul, oracle):
    lp = abi.QmpcLegParams()
    abi.load_library().qmpc_default_leg_params(C.byref(lp))
    rng = np.random.default_rng(7)
    q = np.stack([rng.uniform(-0.8, 0.8, (512, 4)), rng.uniform(-1.05, 4.19, (512, 4)),
                  rng.uniform(-2.69, -0.92, (512, 4))], axis=2).reshape(512, 12)
    foot, jac = np.zeros((512, 12)), np.zeros((512, 36))
    assert emul.emul_leg_kinematics(C.byref(lp), q.ctypes.data, 512, foot.ctypes.data, jac.ctypes.data) == 0
    rf, rj = oracle.leg_kinematics(lp, q)
    assert np.abs(foot - rf).max() < 1e-12 and np.abs(jac - rj).max() < 1e-12   # fp64, different term order
