"""Rows N1 / N2 of the scope table: gait-table contact prediction, leg kinematics,
joint-torque mapping.  CPU tests pin the oracle restatement (oracle/periph_ref.c) against
independent statements of the same maths; the -m gpu tests compare the CUDA kernels (through the
C-ABI) with that oracle: bit-exact for the integer schedules, 1e-12 for the fp64 kinematics/torques
(device sin/cos and FMA contraction differ from libm in the last ulp)."""
import ctypes as C

import numpy as np
import pytest

from quaternion_mpc_b200 import abi
from quaternion_mpc_b200.config import GO1_NOMINAL_FEET, default_config
from quaternion_mpc_b200.workloads import (GAIT_CRAWL, GAIT_STAND, GAIT_TROT, GAIT_TROT_WITH_STAND,
                                           predict_schedule_numpy, random_batch, random_gait_states)


def _leg_params():
    lp = abi.QmpcLegParams()
    assert abi.load_library().qmpc_default_leg_params(C.byref(lp)) == 0
    return lp


def _random_joints(n, seed):
    rng = np.random.default_rng(seed)   # joint ranges quoted in TestInvKin.cpp:38-40
    return np.stack([rng.uniform(-0.8, 0.8, (n, 4)), rng.uniform(-1.05, 4.19, (n, 4)),
                     rng.uniform(-2.69, -0.92, (n, 4))], axis=2).reshape(n, 12)


# ------------------------------------------------------------------------------------------ oracle (CPU)
def test_oracle_predictor_matches_numpy_tables(oracle):
    for N in (1, 10, 20, 32):
        cfg = default_config(abi.QMPC_MODEL_QUAT_4FOOT, N)
        g = random_gait_states(2000, seed=N, gaits=(GAIT_TROT, GAIT_TROT_WITH_STAND, GAIT_CRAWL, GAIT_STAND))
        s = oracle.predict_schedule(cfg, g)
        assert np.array_equal(s, predict_schedule_numpy(g, N, cfg.dt))
        assert (s[:, N:] == 0).all()


def test_oracle_predictor_known_answers(oracle):
    """Hand-derived from the tables (LeggedContactFSM.cpp:87-108): trot at phase 0 -> FL,RR stance
    (mask 1001b = 9) until the phase passes 0.5, then FR,RL (0110b = 6); phase wraps at 1."""
    cfg = default_config(abi.QMPC_MODEL_QUAT_4FOOT, 32)   # dt = 0.01 s
    g = np.zeros(3, dtype=abi.GAIT_STATE_DTYPE)
    g["gait"] = (GAIT_TROT, GAIT_STAND, GAIT_CRAWL)
    g["gait_freq"] = 2.5                                    # 0.025 cycle per knot
    g["gait_phase"] = np.array([0.0, 0.3, 0.9])[:, None]
    s = oracle.predict_schedule(cfg, g)
    assert (s[0, :21] == 9).all() and (s[0, 21:32] == 6).all()          # 0.5 reached at k = 20 (<= is stance)
    assert (s[1, :32] == 15).all()
    # crawl from phase 0.9: RR swings (phase > 0.75) until the wrap; then FL swings while phase <= 0.25
    assert s[2, 0] == 0b0111 and s[2, 4] == 0b0111
    assert s[2, 5] == 0b1110 and s[2, 14] == 0b1110 and s[2, 15] == 0b1101


def test_oracle_kinematics_against_numeric_derivative_and_geometry(oracle):
    lp = _leg_params()
    q = _random_joints(256, 0)
    foot, jac = oracle.leg_kinematics(lp, q)
    eps = 1e-6
    for j in range(3):
        qp, qm = q.reshape(-1, 4, 3).copy(), q.reshape(-1, 4, 3).copy()
        qp[:, :, j] += eps
        qm[:, :, j] -= eps
        d = (oracle.leg_kinematics(lp, qp.reshape(-1, 12))[0] - oracle.leg_kinematics(lp, qm.reshape(-1, 12))[0]) / (2 * eps)
        assert np.abs(d.reshape(-1, 4, 3) - jac.reshape(-1, 4, 3, 3)[:, :, j, :]).max() < 1e-8
    # leg length: |foot - hip joint| depends on the knee angle only (law of cosines, lt = lc = 0.213)
    hip = np.array([[lp.rho_fix[i][0], lp.rho_fix[i][1], 0.0] for i in range(4)])
    rel = foot.reshape(-1, 4, 3) - hip[None]
    d_off = np.array([lp.rho_fix[i][2] for i in range(4)])
    knee = q.reshape(-1, 4, 3)[:, :, 2]
    r2 = 0.213 ** 2 * (2 + 2 * np.cos(knee)) + d_off[None] ** 2
    assert np.abs((rel ** 2).sum(-1) - r2).max() < 1e-12
    # the nominal stance of the shipped config (gazebo_go1_quat_mpc.yaml:16-30) is reachable: x, z signs
    stand, _ = oracle.leg_kinematics(lp, np.array([[0.0, 0.8, -1.6] * 4]))
    assert np.sign(stand.reshape(4, 3)[:, :2]).tolist() == np.sign(np.array(GO1_NOMINAL_FEET)[:, :2]).tolist()


def test_oracle_torques_are_minus_jt_f(oracle):
    rng = np.random.default_rng(1)
    res = np.zeros(64, dtype=abi.RESULT_DTYPE)
    res["grf_body"] = rng.normal(0, 30, (64, 12))
    jac = rng.normal(0, 0.2, (64, 36))
    pc = rng.integers(0, 2, (64, 4)).astype(np.int32)
    J = jac.reshape(64, 4, 3, 3).transpose(0, 1, 3, 2)       # [b, leg, row, col]
    want = -np.einsum("blrc,blr->blc", J, res["grf_body"].reshape(64, 4, 3))
    assert np.abs(oracle.joint_torques(res, jac, pc, 0).reshape(64, 4, 3) - want).max() < 1e-12
    got = oracle.joint_torques(res, jac, pc, 1).reshape(64, 4, 3)
    assert np.abs(got - want * pc[:, :, None]).max() < 1e-12 and (got[pc == 0] == 0).all()


# ------------------------------------------------------------------------------------------ CUDA (GPU)
@pytest.mark.gpu
def test_gpu_schedule_predictor_bit_exact(oracle):
    import torch
    from quaternion_mpc_b200 import QuatMpc
    for N, B in ((10, 100_000), (20, 4097), (32, 1)):
        mpc = QuatMpc(horizon=N, max_batch=B)
        g = random_gait_states(B, seed=N, gaits=(0, 1, 2, 3))
        g["gait_freq"] = np.random.default_rng(N).uniform(0.5, 4.0, B)
        d_g = torch.from_numpy(g.view(np.uint8).reshape(B, -1)).cuda()
        s = mpc.predict_contact_schedule(d_g).cpu().numpy()
        assert np.array_equal(s, oracle.predict_schedule(mpc.cfg, g))
        mpc.close()


@pytest.mark.gpu
def test_gpu_leg_kinematics_and_torques(oracle):
    import torch
    from quaternion_mpc_b200 import QuatMpc
    B = 50_001
    mpc = QuatMpc(horizon=10, max_batch=B)
    lp = _leg_params()
    q = _random_joints(B, 3)
    foot, jac = mpc.leg_kinematics(torch.from_numpy(q).cuda())
    rf, rj = oracle.leg_kinematics(lp, q)
    assert np.abs(foot.cpu().numpy() - rf).max() < 1e-12
    assert np.abs(jac.cpu().numpy() - rj).max() < 1e-12
    rng = np.random.default_rng(4)
    res = np.zeros(B, dtype=abi.RESULT_DTYPE)
    res["grf_body"] = rng.normal(0, 40, (B, 12))
    pc = rng.integers(0, 2, (B, 4)).astype(np.int32)
    d_res = torch.from_numpy(res.view(np.uint8).reshape(B, -1)).cuda()
    for mode in (0, 1):
        tau = mpc.joint_torques(d_res, jac, torch.from_numpy(pc).cuda(), movement_mode=mode).cpu().numpy()
        assert np.abs(tau - oracle.joint_torques(res, rj, pc, mode)).max() < 1e-11
    tau = mpc.joint_torques(d_res, jac, None, movement_mode=1).cpu().numpy()
    assert np.abs(tau - oracle.joint_torques(res, rj, None, 1)).max() < 1e-11


@pytest.mark.gpu
def test_gpu_pipeline_joints_to_torques(oracle):
    """joint angles -> foot positions -> scheduled QuatMpc solve -> joint torques, all on the device,
    against the same chain through the oracle."""
    import torch
    from quaternion_mpc_b200 import QuatMpc
    B = 1024
    mpc = QuatMpc(horizon=10, max_batch=B)
    lp = _leg_params()
    rng = np.random.default_rng(5)
    q = np.array([0.0, 0.8, -1.6] * 4)[None] + rng.uniform(-0.15, 0.15, (B, 12))
    p = random_batch(B, seed=5, gait="trot")
    g = random_gait_states(B, seed=5)
    d_g = torch.from_numpy(g.view(np.uint8).reshape(B, -1)).cuda()
    d_sched = mpc.predict_contact_schedule(d_g)
    foot, jac = mpc.leg_kinematics(torch.from_numpy(q).cuda())
    d_p = mpc.to_device(p)
    off = abi.PROBLEM_DTYPE.fields["foot_pos_body"][1]
    d_p[:, off:off + 96] = foot.view(torch.uint8).reshape(B, 96)          # feed the solve from the FK output
    sched = d_sched.cpu().numpy()
    pc = torch.from_numpy(np.stack([(sched[:, 0] >> i) & 1 for i in range(4)], 1).astype(np.int32)).cuda()
    d_res = mpc.grf_update_sched_device(d_p, d_sched)
    tau = mpc.joint_torques(d_res, jac, pc, movement_mode=1).cpu().numpy()
    res = mpc.results_to_numpy(d_res)
    # oracle chain
    rf, rj = oracle.leg_kinematics(lp, q)
    p["foot_pos_body"] = rf
    rs = oracle.predict_schedule(mpc.cfg, g)
    assert np.array_equal(rs, sched)
    ref = oracle.solve_batch_sched(mpc.cfg, p, rs, nthreads=8)
    ok = (res["status"] < 2) & (ref["status"] < 2)
    assert ok.mean() > 0.9
    assert np.abs(res["grf_body"][ok] - ref["grf_body"][ok]).max() < 1e-4
    rtau = oracle.joint_torques(ref, rj, pc.cpu().numpy(), 1)
    assert np.abs(tau[ok] - rtau[ok]).max() < 1e-4
