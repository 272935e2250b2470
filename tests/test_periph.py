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


# ------------------------------------------------------------------------------------------ row N3
def _goal_inputs(n, seed, tick=0):
    rng = np.random.default_rng(seed * 1000 + tick)
    g = np.zeros(n, dtype=abi.GOAL_INPUT_DTYPE)
    g["joy_vel"] = np.stack([rng.uniform(-0.5, 0.5, n), rng.uniform(-0.1, 0.1, n)], 1)   # joystick scales, yaml:100-101
    g["joy_ang_rate"] = rng.uniform(-0.3, 0.3, (n, 3))
    g["joy_body_height"] = rng.uniform(0.25, 0.32, n)
    g["torso_pos_world"] = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(0.24, 0.33, n)], 1)
    q = rng.normal(size=(n, 4)) * np.array([1, 0.1, 0.1, 0.5]) + np.array([2.0, 0, 0, 0])
    g["torso_quat"] = q / np.linalg.norm(q, axis=1, keepdims=True)
    g["torso_lin_vel_world"] = rng.normal(0, 0.3, (n, 3))
    return g


def test_oracle_goal_update_filter_and_frames(oracle):
    """The restated MovingWindowFilter(100) against a plain numpy moving average over >2 windows, and
    the reference frames of goal_update against a direct evaluation (QuatMpc.cpp:80-106)."""
    n, ticks = 8, 230
    st = oracle.new_goal_state(n)
    probs = np.zeros(n, dtype=abi.PROBLEM_DTYPE)
    hist_v, hist_p = [], []
    pos_d = None
    for t in range(ticks):
        g = _goal_inputs(n, 3, t)
        if t:
            g["torso_pos_world"] = prev_pos + 0.002          # slowly moving robot
        prev_pos = g["torso_pos_world"].copy()
        oracle.goal_update(st, g, probs)
        w, x, y, z = g["torso_quat"].T
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]]).transpose(2, 0, 1)
        yaw = np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
        vw = np.stack([np.cos(yaw) * g["joy_vel"][:, 0] - np.sin(yaw) * g["joy_vel"][:, 1],
                       np.sin(yaw) * g["joy_vel"][:, 0] + np.cos(yaw) * g["joy_vel"][:, 1], np.zeros(n)], 1)
        if pos_d is None:
            pos_d = g["torso_pos_world"].copy()
        pos_d[:, :2] += vw[:, :2] * 5.0 / 1000.0
        pos_d[:, 2] = g["joy_body_height"]
        hist_v.append(np.einsum("bji,bj->bi", R, vw))
        hist_p.append(np.einsum("bji,bj->bi", R, pos_d - g["torso_pos_world"]))
        want_v = np.sum(hist_v[-100:], axis=0) / 100.0       # divides by the window size even while filling
        want_p = np.sum(hist_p[-100:], axis=0) / 100.0
        assert np.abs(probs["torso_lin_vel_d_body"] - want_v).max() < 1e-12
        assert np.abs(probs["torso_pos_d_body"] - want_p).max() < 1e-12
        assert np.array_equal(probs["torso_ang_vel_d_body"], g["joy_ang_rate"])


def test_emulated_goal_update_and_raibert_bodies(oracle, tmp_path):
    """The kernel bodies (tests/emul) against the oracle over 250 ticks: the Neumaier filter state is
    carried in the element-major device layout."""
    import os
    import subprocess
    so = str(tmp_path / "libqmpc_emul.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                           os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul", "emul.cpp")])
    em = C.CDLL(so)
    n, cap = 33, 40
    state = np.zeros(em.emul_goal_state_doubles() * cap)
    st = oracle.new_goal_state(n)
    a, b = np.zeros(n, dtype=abi.PROBLEM_DTYPE), np.zeros(n, dtype=abi.PROBLEM_DTYPE)
    em.emul_goal_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    for t in range(250):
        g = _goal_inputs(n, 5, t)
        assert em.emul_goal_update(state.ctypes.data, cap, g.ctypes.data, n, a.ctypes.data) == 0
        oracle.goal_update(st, g, b)
        for f in ("torso_lin_vel_d_body", "torso_pos_d_body", "torso_ang_vel_d_body", "torso_quat", "torso_lin_vel_world"):
            assert np.abs(a[f] - b[f]).max() < 1e-12, (t, f)
    rp = abi.QmpcRaibertParams()
    abi.load_library().qmpc_default_raibert_params(C.byref(rp))
    g = _goal_inputs(n, 6)
    tw, tr = np.zeros((n, 12)), np.zeros((n, 12))
    em.emul_raibert.argtypes = [C.POINTER(abi.QmpcRaibertParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    assert em.emul_raibert(C.byref(rp), g.ctypes.data, n, tw.ctypes.data, tr.ctypes.data) == 0
    rw, rr = oracle.raibert_targets(rp, g)
    assert np.abs(tw - rw).max() < 1e-12 and np.abs(tr - rr).max() < 1e-12
    # standing still, no command: targets are the default stance rotated by the yaw, under the torso
    g["torso_lin_vel_world"] = 0
    g["joy_vel"] = 0
    rw, rr = oracle.raibert_targets(rp, g)
    assert np.abs(rw.reshape(n, 4, 3)[:, :, 2] - (g["torso_pos_world"][:, 2:3] - 0.30)).max() < 1e-12


@pytest.mark.gpu
def test_gpu_goal_update_and_raibert(oracle):
    import torch
    from quaternion_mpc_b200 import QuatMpc
    B = 5000
    mpc = QuatMpc(horizon=10, max_batch=B + 7)       # stride of the state layout != batch
    d_state = mpc.alloc_goal_state()
    st = oracle.new_goal_state(B)
    ref = np.zeros(B, dtype=abi.PROBLEM_DTYPE)
    d_probs = mpc.to_device(np.zeros(B, dtype=abi.PROBLEM_DTYPE))
    for t in range(130):                               # past one full filter window
        g = _goal_inputs(B, 9, t)
        d_g = torch.from_numpy(g.view(np.uint8).reshape(B, -1)).cuda()
        mpc.goal_update(d_state, d_g, d_probs)
        oracle.goal_update(st, g, ref)
        if t in (0, 1, 99, 100, 129):
            got = d_probs.cpu().numpy().reshape(-1).view(abi.PROBLEM_DTYPE)
            for f in ("torso_lin_vel_d_body", "torso_pos_d_body", "torso_ang_vel_d_body", "torso_quat",
                      "torso_lin_vel_world"):
                assert np.abs(got[f] - ref[f]).max() < 1e-11, (t, f)   # fp64; device atan2/sin/cos differ in the last ulp
            assert (got["foot_pos_body"] == 0).all()    # fields goal_update does not own are untouched
    tw, tr = mpc.raibert_targets(d_g)
    rp = abi.QmpcRaibertParams()
    mpc.lib.qmpc_default_raibert_params(C.byref(rp))
    rw, rr = oracle.raibert_targets(rp, g)
    assert np.abs(tw.cpu().numpy() - rw).max() < 1e-11 and np.abs(tr.cpu().numpy() - rr).max() < 1e-11


# ------------------------------------------------------------------------------------------ row N3, gait-FSM half
def _foot_inputs(n, seed, tick, mode=1, flag_prob=0.3):
    """Feet under a slowly walking torso, Raibert-like targets ahead of them, random foot-force flags."""
    rng = np.random.default_rng(seed * 7919 + tick)
    i = np.zeros(n, dtype=abi.FOOT_UPDATE_INPUT_DTYPE)
    base = np.array([0.20, 0.14, 0.0, 0.20, -0.14, 0.0, -0.20, 0.14, 0.0, -0.20, -0.14, 0.0])[None]
    drift = 0.001 * tick
    i["foot_pos_world"] = base + drift + rng.uniform(-0.02, 0.02, (n, 12))
    i["foot_pos_target_world"] = base + drift + 0.05 + rng.uniform(-0.03, 0.03, (n, 12))
    i["foot_contact_flag"] = rng.uniform(size=(n, 4)) < flag_prob
    i["movement_mode"] = mode
    return i


def _quintic_numpy(t, T, p0, pT):
    """Independent statement of QuinticCurve::get_foot_swing_target (Utils.cpp:236-293): the unique quintic through
    the six conditions per axis, solved with numpy (float T and t as in the reference's signature)."""
    t, T = float(np.float32(t)), float(np.float32(T))
    rows = lambda s: (np.array([1, s, s**2, s**3, s**4, s**5]), np.array([0, 1, 2 * s, 3 * s**2, 4 * s**3, 5 * s**4]))
    Cm = np.stack([rows(0)[0], rows(T)[0], rows(0)[1], rows(T)[1], rows(T / 2)[0], rows(T / 2)[1]])
    d = pT[:2] - p0[:2]
    vmid = 1.26 / T * d            # k |d| (cos, sin)(theta) with the signs restored = k d
    out = np.zeros(9)
    for ax in range(3):
        con = [p0[ax], pT[ax], 0, 0, (p0[ax] + pT[ax]) / 2, vmid[ax]] if ax < 2 else [p0[2], pT[2], 0.1, -0.1, 0.1, 0.0]
        a = np.linalg.solve(Cm, np.array(con, float))
        out[ax] = np.polyval(a[::-1], t)
        out[3 + ax] = np.polyval(np.polyder(a[::-1]), t)
        out[6 + ax] = np.polyval(np.polyder(a[::-1], 2), t)
    return out


def test_oracle_leg_fsm_follows_the_gait_tables_and_the_quintic(oracle):
    """LeggedContactFSM restated (oracle) against independent statements: contact states of an undisturbed trot
    follow the pattern table tick by tick; swing targets equal the numpy quintic; stance feet hold the touch-down
    position; early contact (> 90 % of the swing + foot-force flag) switches to stance; movement_mode 0 resets."""
    n, freq, dt = 4, 2.2, 5.0 / 1000.0
    st = oracle.new_leg_fsm(n, gait_freq=freq)
    phase = np.zeros(4)
    swing_start = {}
    for tick in range(260):                                   # > 2.8 gait cycles
        inp = _foot_inputs(n, 1, tick, flag_prob=0.0)
        out = oracle.foot_update(st, inp, dt, freq)
        phase = phase + freq * dt                             # every leg of a robot runs the same clock in a trot
        cyc = phase[0] % 1.0 if phase[0] % 1.0 != 0 else 1.0
        # trot table (LeggedContactFSM.cpp:87-108): FL/RR stance for phase < 0.5, FR/RL swing first
        if abs(cyc - 0.5) > 2 * freq * dt and min(cyc, 1 - cyc) > 2 * freq * dt:
            want = np.array([1, 0, 0, 1]) if cyc < 0.5 else np.array([0, 1, 1, 0])
            assert (out["plan_contacts"] == want).all(), (tick, cyc, out["plan_contacts"][0])
        for leg in range(4):
            sw = out["plan_contacts"][0, leg] == 0
            if sw and (leg not in swing_start):
                # a leg that begins in swing at reset starts its state at phase 0 (reset(), :16); one entered through
                # a transition starts it at the phase of the entering tick (common_enter, :217)
                swing_start[leg] = (inp["foot_pos_world"][0, 3 * leg:3 * leg + 3].copy(),
                                    0.0 if tick == 0 else out["gait_counter"][0, leg])
            if not sw:
                swing_start.pop(leg, None)
        # swing feet: the oracle's target equals the independent quintic evaluated at the same state percent
        pat_end = {0: 1.0, 3: 1.0, 1: 0.5, 2: 0.5}
        for leg, (p0, ph0) in list(swing_start.items()):
            if out["plan_contacts"][0, leg] != 0:
                continue
            start = ph0
            pct = min(max((out["gait_counter"][0, leg] - start) / (pat_end[leg] - start), 0.0), 1.0)
            want = _quintic_numpy(0.5 * pct / freq, 0.5 / freq, p0, inp["foot_pos_target_world"][0, 3 * leg:3 * leg + 3])
            got = np.concatenate([out[f][0, 3 * leg:3 * leg + 3] for f in ("foot_pos_target", "foot_vel_target", "foot_acc_target")])
            # the reference builds C from FLOAT powers of T (Utils.cpp:236-244): ~1e-7 relative off the exact quintic
            assert np.abs(got[:6] - want[:6]).max() < 1e-5 and np.abs(got[6:] - want[6:]).max() < 1e-3, (tick, leg)
    # swing-curve end points: t = 0 -> start with z velocity 0.1; t = T -> target; apex 0.1 at T / 2
    T = 0.5 / freq
    p0, pT = np.array([0.1, -0.05, 0.0]), np.array([0.25, 0.02, 0.01])
    a, m, b = _quintic_numpy(0, T, p0, pT), _quintic_numpy(T / 2, T, p0, pT), _quintic_numpy(T, T, p0, pT)
    assert np.abs(a[:3] - p0).max() < 1e-9 and abs(a[5] - 0.1) < 1e-9
    assert np.abs(b[:3] - pT).max() < 1e-6 and abs(m[2] - 0.1) < 1e-9
    # early contact: a flagged swing foot past 90 % of its swing lands at once and holds its position
    st = oracle.new_leg_fsm(1, gait_freq=freq)
    landed = None
    for tick in range(200):
        inp = _foot_inputs(1, 2, tick, flag_prob=0.0)
        pct_fr = (tick * freq * dt % 1.0) / 0.5                # FR swings over phase 0 .. 0.5
        if 0.92 < pct_fr < 1.0 and landed is None:
            inp["foot_contact_flag"][0, 1] = 1
        out = oracle.foot_update(st, inp, dt, freq)
        if inp["foot_contact_flag"][0, 1] and landed is None:
            assert out["plan_contacts"][0, 1] == 1              # before the table's switch time
            landed = inp["foot_pos_world"][0, 3:6].copy()
            assert np.array_equal(out["foot_pos_target"][0, 3:6], landed) and (out["foot_vel_target"][0, 3:6] == 0).all()
    assert landed is not None
    # movement_mode 0: reset, every foot planned in contact, phase 0
    out = oracle.foot_update(st, _foot_inputs(1, 2, 500, mode=0), dt, freq)
    assert (out["plan_contacts"] == 1).all() and (out["gait_counter"] == 0).all()


def _emul_lib(tmp_path):
    import os
    import subprocess
    so = str(tmp_path / "libqmpc_emul.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                           os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul", "emul.cpp")])
    return C.CDLL(so)


def test_emulated_leg_fsm_body_matches_oracle(oracle, tmp_path):
    """The kernel body (tests/emul) against the oracle over 400 ticks, all four gait patterns, random foot-force
    flags, mode switches: contact states and gait counters bit-exact, swing targets to 1e-12."""
    em = _emul_lib(tmp_path)
    n, cap, freq, dt = 64, 70, 2.2, 5.0 / 1000.0
    gait = (np.arange(n) % 4).astype(np.int32)
    state = np.zeros(em.emul_fsm_state_doubles() * 4 * cap)
    em.emul_leg_fsm_init.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    em.emul_foot_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]
    assert em.emul_leg_fsm_init(state.ctypes.data, cap, gait.ctypes.data, n) == 0
    st = oracle.new_leg_fsm(n, gait=gait, gait_freq=freq)
    seen = set()
    for tick in range(400):
        inp = _foot_inputs(n, 3, tick, mode=0 if tick in (0, 1, 250) else 1)
        got = np.zeros(n, dtype=abi.FOOT_UPDATE_OUTPUT_DTYPE)
        assert em.emul_foot_update(state.ctypes.data, cap, inp.ctypes.data, dt, freq, n, got.ctypes.data) == 0
        ref = oracle.foot_update(st, inp, dt, freq)
        assert np.array_equal(got["plan_contacts"], ref["plan_contacts"]), tick
        assert np.array_equal(got["gait_counter"], ref["gait_counter"]), tick
        for f in ("foot_pos_target", "foot_vel_target", "foot_acc_target"):
            assert np.abs(got[f] - ref[f]).max() < 1e-12, (tick, f)
        seen |= set(map(tuple, ref["plan_contacts"]))
    assert len(seen) >= 6          # trot, trot-with-stand, crawl and stand masks all occurred


@pytest.mark.gpu
def test_gpu_foot_update_matches_oracle_and_closes_the_loop(oracle):
    """Batched QuatMpc::foot_update on the device against the oracle (contact states / gait counters bit-exact,
    swing targets 1e-12), then one whole controller tick without leaving the device:
    goal_update -> foot_update (writes plan_contacts + gait states) -> predict schedule -> scheduled solve -> torques."""
    import torch
    from quaternion_mpc_b200 import QuatMpc
    B, freq, dt = 3000, 2.2, 5.0 / 1000.0
    mpc = QuatMpc(horizon=10, max_batch=B)
    gait = (np.arange(B) % 4).astype(np.int32)
    d_fsm = mpc.alloc_leg_fsm(torch.from_numpy(gait).cuda())
    st = oracle.new_leg_fsm(B, gait=gait, gait_freq=freq)
    d_probs = mpc.to_device(random_batch(B, seed=8, gait="trot"))
    d_gait = torch.zeros((B, abi.GAIT_STATE_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    for tick in range(150):
        inp = _foot_inputs(B, 4, tick, mode=0 if tick == 0 else 1)
        d_out = mpc.foot_update(d_fsm, torch.from_numpy(inp.view(np.uint8).reshape(B, -1)).cuda(), freq, dt, d_probs, d_gait)
        ref = oracle.foot_update(st, inp, dt, freq)
        if tick % 10 == 0 or tick > 140:
            got = d_out.cpu().numpy().reshape(-1).view(abi.FOOT_UPDATE_OUTPUT_DTYPE)
            assert np.array_equal(got["plan_contacts"], ref["plan_contacts"]) and np.array_equal(got["gait_counter"], ref["gait_counter"])
            for f in ("foot_pos_target", "foot_vel_target", "foot_acc_target"):
                assert np.abs(got[f] - ref[f]).max() < 1e-12, (tick, f)
    probs = d_probs.cpu().numpy().reshape(-1).view(abi.PROBLEM_DTYPE)
    assert np.array_equal(probs["plan_contacts"], ref["plan_contacts"])
    g = d_gait.cpu().numpy().reshape(-1).view(abi.GAIT_STATE_DTYPE)
    assert np.array_equal(g["gait_phase"], ref["gait_counter"]) and (g["gait"] == gait).all() and (g["gait_freq"] == freq).all()
    # the rest of the tick on the device: schedule from the FSM's own gait states, solve, torques
    d_sched = mpc.predict_contact_schedule(d_gait)
    d_res = mpc.grf_update_sched_device(d_probs, d_sched)
    q = np.array([0.0, 0.8, -1.6] * 4)[None] + np.random.default_rng(1).uniform(-0.1, 0.1, (B, 12))
    _, jac = mpc.leg_kinematics(torch.from_numpy(q).cuda())
    pc = torch.from_numpy(ref["plan_contacts"].astype(np.int32)).cuda()
    tau = mpc.joint_torques(d_res, jac, pc, movement_mode=1).cpu().numpy()
    sched = oracle.predict_schedule(mpc.cfg, g)
    assert np.array_equal(d_sched.cpu().numpy(), sched)
    rr, ratio = oracle.solve_batch_diag(mpc.cfg, probs, schedule=sched, nthreads=8)
    res = mpc.results_to_numpy(d_res)
    # same parity policy as tests/test_gpu_parity.py::_check: solves neither side flags and whose Quu stayed below the
    # fp64 conditioning limit must agree in status, iteration count and GRFs (1e-4 N); the others are counted
    err = np.abs(res["grf_body"] - rr["grf_body"]).max(axis=1)
    exempt = (res["status"] >= 2) | (rr["status"] >= 2) | (ratio > 1e12)
    agree = (res["status"] == rr["status"]) & (res["iterations"] == rr["iterations"]) & (err < 1e-4)
    print(f"[pipeline parity: solves={B} exempt={int(exempt.sum())} exempt_differ={int((exempt & ~agree).sum())} "
          f"disagree={int((~exempt & ~agree).sum())} max_err={float(err[agree].max()):.2e}]", end=" ")
    assert not (~exempt & ~agree).any(), (int((~exempt & ~agree).sum()), float(err[~exempt & ~agree].max()))
    assert (exempt & ~agree).sum() <= B // 100
    ok = agree
    rtau = oracle.joint_torques(rr, oracle.leg_kinematics(_leg_params(), q)[1], ref["plan_contacts"].astype(np.int32), 1)
    assert np.abs(tau[ok] - rtau[ok]).max() < 1e-4


@pytest.mark.gpu
def test_gpu_schedule_predictor_survives_non_finite_and_huge_phases():
    """One bad robot record (Inf / NaN / 1e12 gait phase or frequency) must not hang the launch: the reference's
    `while (ph > 1.0) ph -= 1.0` is evaluated in closed form and absurd phases fall through to STANCE."""
    import torch
    from quaternion_mpc_b200 import QuatMpc
    mpc = QuatMpc(horizon=16, max_batch=64)
    g = random_gait_states(64, seed=3)
    g["gait_phase"][1] = np.inf
    g["gait_phase"][2] = np.nan
    g["gait_phase"][3] = 1e12 + 0.25
    g["gait_freq"][4] = np.inf
    g["gait_phase"][5] = 7.0          # a whole number wraps to 1.0, not 0.0
    out = mpc.predict_contact_schedule(torch.from_numpy(g.view(np.uint8).reshape(64, -1)).cuda()).cpu().numpy()
    want = predict_schedule_numpy(g[6:], 16, mpc.cfg.dt)
    assert np.array_equal(out[6:], want)
    assert (out[1, :16] == 15).all() and (out[2, 0] == 15) and (out[4, 1:16] == 15).all()
    g5 = g[5:6].copy(); g5["gait_phase"] = 1.0
    assert np.array_equal(out[5], predict_schedule_numpy(g5, 16, mpc.cfg.dt)[0])
