"""Shipped Go1 solver/robot constants (host-side mirror of qmpc_default_config in csrc/qmpc_api.cu).

Sources (relative to the reference tree):
  legged_ctrl/config/gazebo_go1_quat_mpc.yaml:35-75,115-122   Q, R, w, mu, fz_max, mass, inertia
  legged_ctrl/config/gazebo_go1_convex_mpc.yaml:36-72,129     ConvexMpc weights
  legged_ctrl/src/mpc/QuatMpc.cpp:21-26,182                   AltroOptions, 1.2 x inertia
  legged_ctrl/src/mpc/ConvexMpc.cpp:36-38                     iterations_max = 5
  legged_ctrl/src/utils/AltroUtils.cpp:373-374                COM offset / trunk mass
"""
from .abi import (QMPC_MODEL_EULER_CONVEX, QMPC_MODEL_QUAT_2FOOT, QMPC_MODEL_QUAT_4FOOT,
                  QMPC_MAX_HORIZON, QmpcConfig)

TRUNK_INERTIA = (0.0168128557, 0.063009565, 0.0716547275)
GO1_MASS = 12.84
GO1_NOMINAL_FEET = ((0.20, 0.14, -0.30), (0.20, -0.14, -0.30), (-0.20, 0.14, -0.30), (-0.20, -0.14, -0.30))


def default_config(model=QMPC_MODEL_QUAT_4FOOT, horizon=10):
    if model not in (QMPC_MODEL_QUAT_4FOOT, QMPC_MODEL_QUAT_2FOOT, QMPC_MODEL_EULER_CONVEX):
        raise ValueError("unknown model")
    if not 1 <= horizon <= QMPC_MAX_HORIZON:
        raise ValueError("horizon out of range")
    c = QmpcConfig()
    c.model, c.horizon = model, horizon
    c.robot_mass, c.gravity, c.quat_d_dt = GO1_MASS, 9.81, 5.0 / 1000.0
    c.com_offset[:] = (0.0223, 0.002, -0.0005)
    c.com_mass = 5.204
    c.r_weights[:] = [1e-6] * 12
    c.penalty_initial, c.penalty_max = 1.0, 1e8
    c.tol_cost_intermediate = c.tol_primal_feasibility = c.tol_stationarity = 1e-4
    c.drop_omega0 = 1
    if model == QMPC_MODEL_EULER_CONVEX:
        c.dt = 5.0 / 1000.0
        c.q_weights[:] = [3.0, 3.0, 3.0, 1.0, 1.0, 20.0, 0.0, 0.0, 3.0, 2.0, 3.0, 2.0, 0.0]
        c.w, c.mu, c.fz_max = 0.0, 0.6, 200.0
        scale = 1.0
        c.iterations_max, c.penalty_scaling = 5, 10.0
    else:
        c.dt = 10.0 / 1000.0
        c.q_weights[:] = [2.5, 2.5, 10.0, 0, 0, 0, 0, 0.1, 0.1, 0.1, 0.15, 0.15, 0.15]
        c.w, c.mu, c.fz_max = 50.0, 0.7, 100.0
        scale = 1.2
        c.iterations_max, c.penalty_scaling = 10, 20.0
    for i in range(9):
        c.inertia[i] = 0.0
    for i in range(3):
        c.inertia[4 * i] = scale * TRUNK_INERTIA[i]
    return c
