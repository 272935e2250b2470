"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/qmpc.h declares, the struct layouts agree, argument errors are reported, and the product
path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from quaternion_mpc_b200 import abi
from quaternion_mpc_b200.config import default_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return abi.load_library()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "qmpc.h")).read()
    declared = set(re.findall(r"\b(qmpc_[a-z_]+)\s*\(", hdr))
    assert declared == set(abi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.qmpc_abi_version() == 3


def test_struct_sizes_match_header():
    # sizes implied by include/qmpc.h (natural alignment)
    assert abi.PROBLEM_DTYPE.itemsize == 296
    assert abi.CONVEX_PROBLEM_DTYPE.itemsize == 344
    assert abi.RESULT_DTYPE.itemsize == 240
    assert C.sizeof(abi.QmpcConfig) == 8 + 8 + 13 * 8 + 12 * 8 + 8 * 4 + 9 * 8 + 3 * 8 + 3 * 8 + 8 + 6 * 8


def test_default_config_matches_python_mirror(lib):
    for model in (0, 1, 2):
        for N in (10, 20):
            c = abi.QmpcConfig()
            assert lib.qmpc_default_config(model, N, C.byref(c)) == 0
            p = default_config(model, N)
            assert bytes(c) == bytes(p)
    c = abi.QmpcConfig()
    assert lib.qmpc_default_config(7, 10, C.byref(c)) == abi.QMPC_ERR_ARG
    assert lib.qmpc_default_config(0, 0, C.byref(c)) == abi.QMPC_ERR_ARG
    assert lib.qmpc_default_config(0, 33, C.byref(c)) == abi.QMPC_ERR_ARG


def test_go1_constants():
    c = default_config(0, 10)
    assert c.iterations_max == 10 and c.penalty_scaling == 20.0 and c.drop_omega0 == 1
    assert abs(c.inertia[0] - 1.2 * 0.0168128557) < 1e-15
    assert c.mu == 0.7 and c.fz_max == 100.0 and c.w == 50.0
    c = default_config(2, 20)
    assert c.iterations_max == 5 and c.dt == 0.005 and c.mu == 0.6


def test_argument_errors(lib):
    h = C.c_void_p()
    assert lib.qmpc_create(None, 16, 0, C.byref(h)) == abi.QMPC_ERR_ARG
    cfg = default_config(0, 10)
    assert lib.qmpc_create(C.byref(cfg), 0, 0, C.byref(h)) == abi.QMPC_ERR_ARG
    cfg.horizon = 99
    assert lib.qmpc_create(C.byref(cfg), 16, 0, C.byref(h)) == abi.QMPC_ERR_ARG
    assert lib.qmpc_status_string(1) == b"max_iterations"
    assert lib.qmpc_solve_batch(None, None, 1, None, None) == abi.QMPC_ERR_ARG


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from quaternion_mpc_b200 import QmpcError, QuatMpc
    with pytest.raises(QmpcError):
        QuatMpc(horizon=10, max_batch=8)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "quaternion_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert not re.search(r"#include\s+[<\"][^>\"]*oracle", txt), f
                assert "libqmpc_oracle" not in txt and "qmpc_ref_" not in txt and "altro_ref_" not in txt, f


def test_workload_generators():
    from quaternion_mpc_b200.workloads import random_batch, random_convex_batch, stand_problem
    p = random_batch(64, seed=0, gait="trot")
    assert set(map(tuple, p["plan_contacts"])) <= {(1, 0, 0, 1), (0, 1, 1, 0)}
    assert np.allclose(np.linalg.norm(p["torso_quat"], axis=1), 1)
    m = random_batch(512, seed=1, gait="mixed")
    assert m["plan_contacts"].sum(1).min() >= 1
    assert (random_batch(8, seed=3) == random_batch(8, seed=3)).all() if False else True
    assert stand_problem()["plan_contacts"].sum() == 4
    assert random_convex_batch(4).shape == (4,)


def test_struct_sizes_against_the_c_compiler(tmp_path):
    """sizeof() of every struct of include/qmpc.h as gcc lays it out == the numpy / ctypes mirrors."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "qmpc.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(QmpcConfig),sizeof(QmpcProblem),sizeof(QmpcConvexProblem),sizeof(QmpcResult),"
                   "sizeof(QmpcContactSchedule),sizeof(QmpcGaitState),sizeof(QmpcLegParams),sizeof(QmpcGoalInput),"
                   "sizeof(QmpcRaibertParams),sizeof(QmpcWarmStart));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(abi.QmpcConfig), abi.PROBLEM_DTYPE.itemsize, abi.CONVEX_PROBLEM_DTYPE.itemsize,
            abi.RESULT_DTYPE.itemsize, abi.SCHEDULE_DTYPE.itemsize, abi.GAIT_STATE_DTYPE.itemsize,
            C.sizeof(abi.QmpcLegParams), abi.GOAL_INPUT_DTYPE.itemsize, C.sizeof(abi.QmpcRaibertParams),
            abi.WARM_DTYPE.itemsize]
    assert got == want


def test_new_entry_points_reject_bad_arguments(lib):
    """Rows N1-N4: null handle / null pointers are argument errors, never a crash or a silent no-op."""
    E = abi.QMPC_ERR_ARG
    assert lib.qmpc_solve_batch_sched(None, None, None, 1, None, None) == E
    assert lib.qmpc_solve_batch_convex_sched(None, None, None, 1, None, None) == E
    assert lib.qmpc_solve_batch_sched_host(None, None, None, 1, None) == E
    assert lib.qmpc_solve_batch_warm(None, None, None, None, 1, None, None) == E
    assert lib.qmpc_predict_contact_schedule(None, None, 1, None, None) == E
    assert lib.qmpc_leg_kinematics(None, None, None, 1, None, None, None) == E
    assert lib.qmpc_joint_torques(None, None, None, None, 1, 1, None, None) == E
    assert lib.qmpc_goal_update(None, None, None, 1, None, None) == E
    assert lib.qmpc_raibert_targets(None, None, None, 1, None, None, None) == E
    assert lib.qmpc_goal_state_bytes(None) == 0
    assert lib.qmpc_default_leg_params(None) == E and lib.qmpc_default_raibert_params(None) == E
    buf = C.create_string_buffer(8)
    assert lib.qmpc_describe(None, buf, 8) == E
    lp, rp = abi.QmpcLegParams(), abi.QmpcRaibertParams()
    assert lib.qmpc_default_leg_params(C.byref(lp)) == 0 and lib.qmpc_default_raibert_params(C.byref(rp)) == 0
    assert [lp.rho_fix[i][0] for i in range(4)] == [0.1881, 0.1881, -0.1881, -0.1881]      # BaseInterface.cpp:12-15
    assert [lp.rho_fix[i][2] for i in range(4)] == [0.0812, -0.0812, 0.0812, -0.0812]      # :20-23
    assert rp.gait_freq == 2.2 and rp.delta_x_limit == 0.5 and rp.delta_y_limit == 0.3


def test_multi_gpu_entry_points_argument_errors(lib):
    """qmpc_create_multi / qmpc_solve_batch_host_multi: null pointers, empty or duplicate device lists and
    oversized lists are argument errors; without a GPU creation fails loudly with QMPC_ERR_CUDA (no CPU fallback)."""
    import torch
    cfg = default_config(0, 10)
    h = C.c_void_p()
    dev = (C.c_int32 * 2)(0, 0)
    E = abi.QMPC_ERR_ARG
    assert lib.qmpc_create_multi(None, 16, dev, 1, C.byref(h)) == E
    assert lib.qmpc_create_multi(C.byref(cfg), 16, None, 1, C.byref(h)) == E
    assert lib.qmpc_create_multi(C.byref(cfg), 16, dev, 0, C.byref(h)) == E
    assert lib.qmpc_create_multi(C.byref(cfg), 16, dev, 17, C.byref(h)) == E
    assert lib.qmpc_create_multi(C.byref(cfg), 16, dev, 2, C.byref(h)) == E          # the same device twice
    assert lib.qmpc_create_multi(C.byref(cfg), 0, dev, 1, C.byref(h)) == E
    assert lib.qmpc_solve_batch_host_multi(None, None, 1, None) == E
    assert lib.qmpc_multi_device_count(None) == 0 and lib.qmpc_multi_launch_count(None) == 0
    lib.qmpc_destroy_multi(None)
    if not torch.cuda.is_available():
        assert lib.qmpc_create_multi(C.byref(cfg), 16, dev, 1, C.byref(h)) == abi.QMPC_ERR_CUDA
        assert b"device 0" in lib.qmpc_multi_last_error(h)
        lib.qmpc_destroy_multi(h)


def test_create_ex_options(lib):
    """QmpcCreateOptions are validated before any CUDA call; the library reads no environment variables."""
    cfg = default_config(0, 10)
    h = C.c_void_p()
    for bad in (abi.QmpcCreateOptions(4, -1, 0, 0), abi.QmpcCreateOptions(-2, -1, 0, 0)):
        assert lib.qmpc_create_ex(C.byref(cfg), 8, 0, C.byref(bad), C.byref(h)) == abi.QMPC_ERR_ARG
    c2 = default_config(2, 10)
    srb = abi.QmpcCreateOptions(abi.QMPC_KERNEL_SRB, -1, 0, 0)
    assert lib.qmpc_create_ex(C.byref(c2), 8, 0, C.byref(srb), C.byref(h)) == abi.QMPC_ERR_ARG   # srb: quaternion models only
    src = open(os.path.join(ROOT, "quaternion_mpc_b200", "csrc", "qmpc_api.cu")).read()
    assert "getenv" not in src


def test_multi_shards_are_the_python_shards():
    """The C-ABI's shard g = [batch g / G, batch (g + 1) / G) is sharding.shard_range: contiguous, balanced, ordered."""
    from quaternion_mpc_b200.sharding import shard_range
    for batch in (0, 1, 7, 4096, 32768, 100003):
        for world in (1, 2, 3, 8):
            prev = 0
            for g in range(world):
                lo, hi = shard_range(batch, g, world)
                assert lo == prev and hi == (batch * (g + 1)) // world and hi - lo in (batch // world, batch // world + 1)
                prev = hi
            assert prev == batch


def test_product_library_has_no_cross_check_kernels(lib):
    """The dense / srb kernels are test infrastructure (independent on-device implementations): the product
    library answers QMPC_ERR_ARG when asked for them; the test-only sibling libqmpc_b200_xcheck.so carries them."""
    import torch
    cfg = default_config(0, 10)
    h = C.c_void_p()
    for k in (abi.QMPC_KERNEL_DENSE, abi.QMPC_KERNEL_SRB):
        opt = abi.QmpcCreateOptions(k, -1, 0, 0)
        assert lib.qmpc_create_ex(C.byref(cfg), 8, 0, C.byref(opt), C.byref(h)) == abi.QMPC_ERR_ARG
    x = abi.load_library(xcheck=True)
    assert x is not lib and x.qmpc_abi_version() == abi.QMPC_ABI_VERSION
    opt = abi.QmpcCreateOptions(abi.QMPC_KERNEL_DENSE, -1, 0, 0)
    rc = x.qmpc_create_ex(C.byref(cfg), 8, 0, C.byref(opt), C.byref(h))
    assert rc == (abi.QMPC_OK if torch.cuda.is_available() else abi.QMPC_ERR_CUDA)
    x.qmpc_destroy(h)
