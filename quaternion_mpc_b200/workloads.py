"""Seeded synthetic Go1 problem batches — the distributions of BASELINE.md section 5 / SURVEY.md 8d.

Pure numpy host code (input generation only; no solver arithmetic).
"""
import numpy as np

from .abi import CONVEX_PROBLEM_DTYPE, PROBLEM_DTYPE
from .config import GO1_NOMINAL_FEET


def _rot(q):
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (y * z + w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def stand_problem():
    """Config 1: single solve, stand, identity attitude, zero velocities, nominal feet."""
    p = np.zeros(1, dtype=PROBLEM_DTYPE)
    p["torso_quat"][0] = (1, 0, 0, 0)
    p["torso_quat_d"][0] = (1, 0, 0, 0)
    p["foot_pos_body"][0] = np.array(GO1_NOMINAL_FEET).reshape(-1)
    p["plan_contacts"][0] = (1, 1, 1, 1)
    return p


def random_batch(batch, seed=0, gait="trot", max_angle=0.5, nfeet=4):
    """Configs 2/3/5.  gait: 'trot' -> masks {1001, 0110}; 'mixed' -> the 15 non-empty masks;
    'stand' -> 1111.  Attitude: axis uniform on S2, angle U(0,max_angle); v_body N(0,0.3^2);
    omega N(0,0.5^2) (ignored by the reference, drop_omega0); feet nominal + U(-0.05,0.05);
    p_ref (U(+-.02), U(+-.02), U(+-.03)); v_ref (U(+-.5), U(+-.1), 0); q_ref yaw-only U(+-0.2)."""
    rng = np.random.default_rng(seed)
    p = np.zeros(batch, dtype=PROBLEM_DTYPE)
    ax = rng.normal(size=(batch, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = rng.uniform(0, max_angle, batch)
    q = np.concatenate([np.cos(ang / 2)[:, None], np.sin(ang / 2)[:, None] * ax], axis=1)
    p["torso_quat"] = q
    v_body = rng.normal(0, 0.3, (batch, 3))
    p["torso_lin_vel_world"] = np.einsum("bij,bj->bi", _rot(q), v_body)
    p["torso_ang_vel_body"] = rng.normal(0, 0.5, (batch, 3))
    feet = np.array(GO1_NOMINAL_FEET).reshape(-1)[None, :] + rng.uniform(-0.05, 0.05, (batch, 12))
    p["foot_pos_body"] = feet
    p["torso_pos_d_body"] = np.stack([rng.uniform(-.02, .02, batch), rng.uniform(-.02, .02, batch),
                                      rng.uniform(-.03, .03, batch)], axis=1)
    p["torso_lin_vel_d_body"] = np.stack([rng.uniform(-.5, .5, batch), rng.uniform(-.1, .1, batch),
                                          np.zeros(batch)], axis=1)
    yaw = rng.uniform(-0.2, 0.2, batch)
    p["torso_quat_d"] = np.stack([np.cos(yaw / 2), np.zeros(batch), np.zeros(batch), np.sin(yaw / 2)], axis=1)
    p["torso_ang_vel_d_body"] = 0.0
    if gait == "trot":
        masks = np.array([[1, 0, 0, 1], [0, 1, 1, 0]])
        p["plan_contacts"] = masks[rng.integers(0, 2, batch)]
    elif gait == "mixed":
        m = rng.integers(1, 16, batch)
        p["plan_contacts"] = np.stack([(m >> 3) & 1, (m >> 2) & 1, (m >> 1) & 1, m & 1], axis=1)
    elif gait == "stand":
        p["plan_contacts"] = 1
    else:
        raise ValueError(gait)
    if nfeet == 2:  # config 4: two-contact model, both feet planted, diagonal pair geometry
        p["plan_contacts"] = (1, 1, 0, 0)
        f2 = np.array([0.17, 0.13, -0.3, -0.17, -0.13, -0.3])[None, :] + rng.uniform(-0.03, 0.03, (batch, 6))
        p["foot_pos_body"][:, :6] = f2
        p["foot_pos_body"][:, 6:] = 0.0
    return p


def random_convex_batch(batch, seed=0, gait="trot"):
    """ConvexMpc inputs with the same spirit: small roll/pitch, world-frame quantities."""
    rng = np.random.default_rng(seed)
    p = np.zeros(batch, dtype=CONVEX_PROBLEM_DTYPE)
    eul = np.stack([rng.uniform(-0.15, 0.15, batch), rng.uniform(-0.15, 0.15, batch),
                    rng.uniform(-1.0, 1.0, batch)], axis=1)
    p["torso_euler"] = eul
    p["torso_pos_world"] = np.stack([rng.uniform(-1, 1, batch), rng.uniform(-1, 1, batch),
                                     rng.uniform(0.25, 0.33, batch)], axis=1)
    p["torso_ang_vel_world"] = rng.normal(0, 0.3, (batch, 3))
    p["torso_lin_vel_world"] = rng.normal(0, 0.3, (batch, 3))
    cr, sr = np.cos(eul[:, 0]), np.sin(eul[:, 0])
    cp, sp = np.cos(eul[:, 1]), np.sin(eul[:, 1])
    cy, sy = np.cos(eul[:, 2]), np.sin(eul[:, 2])
    R = np.empty((batch, 3, 3))
    R[:, 0, 0] = cy * cp; R[:, 0, 1] = cy * sp * sr - sy * cr; R[:, 0, 2] = cy * sp * cr + sy * sr
    R[:, 1, 0] = sy * cp; R[:, 1, 1] = sy * sp * sr + cy * cr; R[:, 1, 2] = sy * sp * cr - cy * sr
    R[:, 2, 0] = -sp; R[:, 2, 1] = cp * sr; R[:, 2, 2] = cp * cr
    p["torso_rot_mat"] = R.reshape(batch, 9)
    feet_b = np.array(GO1_NOMINAL_FEET)[None] + rng.uniform(-0.05, 0.05, (batch, 4, 3))
    p["foot_pos_abs_com"] = np.einsum("bij,bfj->bfi", R, feet_b).reshape(batch, 12)
    p["torso_pos_d_world"] = p["torso_pos_world"] + np.stack(
        [rng.uniform(-.02, .02, batch), rng.uniform(-.02, .02, batch), rng.uniform(-.03, .03, batch)], axis=1)
    p["torso_lin_vel_d_world"] = np.stack([rng.uniform(-.5, .5, batch), rng.uniform(-.1, .1, batch),
                                           np.zeros(batch)], axis=1)
    p["yaw_rate_d"] = rng.uniform(-0.5, 0.5, batch)
    if gait == "trot":
        masks = np.array([[1, 0, 0, 1], [0, 1, 1, 0]])
        p["plan_contacts"] = masks[rng.integers(0, 2, batch)]
    else:
        p["plan_contacts"] = 1
    return p


# Gait phase tables of LeggedContactFSM (legged_ctrl/src/utils/LeggedContactFSM.cpp:87-206):
# per leg (FL, FR, RL, RR) a list of (switch_time, in_stance) segments over one gait cycle.
GAIT_TROT, GAIT_TROT_WITH_STAND, GAIT_CRAWL, GAIT_STAND = 0, 1, 2, 3
GAIT_TABLES = {
    GAIT_TROT: [[(0.5, 1), (1.0, 0)], [(0.5, 0), (1.0, 1)], [(0.5, 0), (1.0, 1)], [(0.5, 1), (1.0, 0)]],
    GAIT_TROT_WITH_STAND: [[(0.6, 1), (1.0, 0)], [(0.1, 1), (0.5, 0), (1.0, 1)],
                           [(0.1, 1), (0.5, 0), (1.0, 1)], [(0.6, 1), (1.0, 0)]],
    GAIT_CRAWL: [[(0.25, 0), (1.0, 1)], [(0.25, 1), (0.5, 0), (1.0, 1)],
                 [(0.5, 1), (0.75, 0), (1.0, 1)], [(0.75, 1), (1.0, 0)]],
    GAIT_STAND: [[(1.0, 1)]] * 4,
}

from .abi import GAIT_STATE_DTYPE  # noqa: E402


def random_gait_states(batch, seed=0, gaits=(GAIT_TROT, GAIT_TROT_WITH_STAND, GAIT_CRAWL), gait_freq=2.2):
    """Per-robot gait clocks: one phase per leg FSM (all four legs of a robot share the clock in the
    reference's main loop, QuatMpc.cpp:288-300, so the four phases are equal), uniform in [0,1)."""
    rng = np.random.default_rng(seed)
    g = np.zeros(batch, dtype=GAIT_STATE_DTYPE)
    g["gait_phase"] = rng.uniform(0, 1, batch)[:, None]
    g["gait_freq"] = gait_freq
    g["gait"] = np.asarray(gaits)[rng.integers(0, len(gaits), batch)]
    return g


def predict_schedule_numpy(gait_states, horizon, dt):
    """Host restatement of LeggedContactFSM::predict_contact_state (LeggedContactFSM.cpp:272-286) at
    t + k*dt for k = 0..horizon-1 -> (batch, QMPC_MAX_HORIZON) uint8 contact masks (input generator
    for tests; the product's batched predictor is qmpc_predict_contact_schedule)."""
    from .abi import QMPC_MAX_HORIZON
    out = np.zeros((len(gait_states), QMPC_MAX_HORIZON), dtype=np.uint8)
    for b, g in enumerate(gait_states):
        table = GAIT_TABLES[int(g["gait"])]
        for k in range(horizon):
            m = 0
            for leg in range(4):
                ph = g["gait_phase"][leg] + g["gait_freq"] * (k * dt)
                while ph > 1.0:
                    ph -= 1.0
                stance = 1
                for sw, st in table[leg]:
                    if ph <= sw:
                        stance = st
                        break
                m |= stance << leg
            out[b, k] = m
    return out


def mirror_problems(p):
    """Reflection y -> -y of the whole scene (a symmetry of the single-rigid-body problem when cfg.com_offset[1]
    is negated too): vectors flip y, pseudo-vectors flip x and z, quaternions (w, x, y, z) flip x and z, left and
    right feet swap (FL<->FR, RL<->RR).  Used by the size-independent parity-by-property tests."""
    m = p.copy()
    vec = np.array([1.0, -1.0, 1.0])
    axial = np.array([-1.0, 1.0, -1.0])
    quat = np.array([1.0, -1.0, 1.0, -1.0])
    for f in ("torso_lin_vel_world", "torso_pos_d_body", "torso_lin_vel_d_body"):
        m[f] = p[f] * vec
    for f in ("torso_ang_vel_body", "torso_ang_vel_d_body"):
        m[f] = p[f] * axial
    for f in ("torso_quat", "torso_quat_d"):
        m[f] = p[f] * quat
    swap = [1, 0, 3, 2]
    m["foot_pos_body"] = (p["foot_pos_body"].reshape(-1, 4, 3)[:, swap, :] * vec).reshape(-1, 12)
    m["plan_contacts"] = p["plan_contacts"][:, swap]
    return m


def mirror_grf(g):
    """The reflection of mirror_problems() applied to a [B, 12] array of per-foot forces."""
    return (np.asarray(g).reshape(-1, 4, 3)[:, [1, 0, 3, 2], :] * np.array([1.0, -1.0, 1.0])).reshape(-1, 12)
