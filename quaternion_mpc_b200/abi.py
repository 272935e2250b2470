"""ctypes / numpy mirror of include/qmpc.h (the C-ABI of libqmpc_b200.so).

Only layout definitions and the library loader live here — no arithmetic.  The numpy structured
dtypes are byte-identical to the C structs so a batch is one contiguous array that can be handed
to the C-ABI (host entry points) or copied to the device as raw bytes (device entry points).
"""
import ctypes as C
import os

import numpy as np

QMPC_MODEL_QUAT_4FOOT = 0
QMPC_MODEL_QUAT_2FOOT = 1
QMPC_MODEL_EULER_CONVEX = 2
QMPC_MAX_HORIZON = 32

QMPC_OK = 0
QMPC_ERR_ARG = -1
QMPC_ERR_CUDA = -2
QMPC_ERR_CAPACITY = -3

QMPC_KERNEL_AUTO, QMPC_KERNEL_DENSE, QMPC_KERNEL_SRB, QMPC_KERNEL_COOP, QMPC_KERNEL_PHASED = -1, 0, 1, 2, 3
KERNEL_NAMES = {"auto": -1, "dense": 0, "srb": 1, "coop": 2, "phased": 3}
QMPC_ABI_VERSION = 3

STATUS_NAMES = {0: "success", 1: "max_iterations", 2: "linesearch_failed", 3: "backward_failed",
                4: "nonfinite"}


class QmpcConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("horizon", C.c_int32), ("dt", C.c_double),
        ("q_weights", C.c_double * 13), ("r_weights", C.c_double * 12), ("w", C.c_double),
        ("mu", C.c_double), ("fz_max", C.c_double), ("robot_mass", C.c_double),
        ("inertia", C.c_double * 9), ("com_offset", C.c_double * 3), ("com_mass", C.c_double),
        ("gravity", C.c_double), ("quat_d_dt", C.c_double),
        ("iterations_max", C.c_int32), ("drop_omega0", C.c_int32),
        ("penalty_initial", C.c_double), ("penalty_scaling", C.c_double), ("penalty_max", C.c_double),
        ("tol_cost_intermediate", C.c_double), ("tol_primal_feasibility", C.c_double),
        ("tol_stationarity", C.c_double),
    ]


class QmpcCreateOptions(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("smem_residents", C.c_int32), ("packed_launch", C.c_int32),
                ("host_chunks", C.c_int32)]


PROBLEM_DTYPE = np.dtype([
    ("torso_quat", "f8", 4), ("torso_lin_vel_world", "f8", 3), ("torso_ang_vel_body", "f8", 3),
    ("foot_pos_body", "f8", 12), ("torso_pos_d_body", "f8", 3), ("torso_lin_vel_d_body", "f8", 3),
    ("torso_quat_d", "f8", 4), ("torso_ang_vel_d_body", "f8", 3), ("plan_contacts", "i4", 4),
], align=True)

CONVEX_PROBLEM_DTYPE = np.dtype([
    ("torso_euler", "f8", 3), ("torso_pos_world", "f8", 3), ("torso_ang_vel_world", "f8", 3),
    ("torso_lin_vel_world", "f8", 3), ("foot_pos_abs_com", "f8", 12), ("torso_rot_mat", "f8", 9),
    ("torso_pos_d_world", "f8", 3), ("torso_lin_vel_d_world", "f8", 3), ("yaw_rate_d", "f8"),
    ("plan_contacts", "i4", 4), ("pad_", "i4", 2),
], align=True)

RESULT_DTYPE = np.dtype([
    ("grf_body", "f8", 12), ("grf_world", "f8", 12), ("torso_quat_d", "f8", 4),
    ("max_violation", "f8"), ("iterations", "i4"), ("status", "i4"),
], align=True)

SCHEDULE_DTYPE = np.dtype([("mask", "u1", QMPC_MAX_HORIZON)])
WARM_DTYPE = np.dtype([("u", "f8", (QMPC_MAX_HORIZON, 12)), ("valid", "i4"), ("pad_", "i4")], align=True)
GAIT_STATE_DTYPE = np.dtype([("gait_phase", "f8", 4), ("gait_freq", "f8"), ("gait", "i4"), ("pad_", "i4")],
                            align=True)
GOAL_INPUT_DTYPE = np.dtype([("joy_vel", "f8", 2), ("joy_ang_rate", "f8", 3), ("joy_body_height", "f8"),
                             ("torso_pos_world", "f8", 3), ("torso_quat", "f8", 4), ("torso_lin_vel_world", "f8", 3)],
                            align=True)
FOOT_UPDATE_INPUT_DTYPE = np.dtype([("foot_pos_world", "f8", 12), ("foot_pos_target_world", "f8", 12),
                                    ("foot_contact_flag", "i4", 4), ("movement_mode", "i4"), ("pad_", "i4", 3)], align=True)
FOOT_UPDATE_OUTPUT_DTYPE = np.dtype([("foot_pos_target", "f8", 12), ("foot_vel_target", "f8", 12), ("foot_acc_target", "f8", 12),
                                     ("gait_counter", "f8", 4), ("plan_contacts", "i4", 4)], align=True)
assert FOOT_UPDATE_INPUT_DTYPE.itemsize == 224 and FOOT_UPDATE_OUTPUT_DTYPE.itemsize == 336
QMPC_GAIT_TROT, QMPC_GAIT_TROT_WITH_STAND, QMPC_GAIT_CRAWL, QMPC_GAIT_STAND = 0, 1, 2, 3


class QmpcRaibertParams(C.Structure):
    _fields_ = [("gait_freq", C.c_double), ("default_foot_pos_rel", C.c_double * 12),
                ("delta_x_limit", C.c_double), ("delta_y_limit", C.c_double)]


class QmpcLegParams(C.Structure):
    _fields_ = [("rho_fix", (C.c_double * 5) * 4), ("rho_opt", (C.c_double * 3) * 4)]


assert PROBLEM_DTYPE.itemsize == 35 * 8 + 16
assert SCHEDULE_DTYPE.itemsize == 32 and GAIT_STATE_DTYPE.itemsize == 48
assert GOAL_INPUT_DTYPE.itemsize == 16 * 8
assert WARM_DTYPE.itemsize == QMPC_MAX_HORIZON * 12 * 8 + 8
assert CONVEX_PROBLEM_DTYPE.itemsize == 40 * 8 + 24
assert RESULT_DTYPE.itemsize == 29 * 8 + 8

EXPORTED_SYMBOLS = [
    "qmpc_default_config", "qmpc_create", "qmpc_solve_batch", "qmpc_solve_batch_convex",
    "qmpc_solve_batch_host", "qmpc_solve_batch_convex_host", "qmpc_destroy", "qmpc_launch_count",
    "qmpc_last_error", "qmpc_status_string", "qmpc_abi_version", "qmpc_measure_fma_peak",
    "qmpc_predict_contact_schedule", "qmpc_solve_batch_sched", "qmpc_solve_batch_convex_sched",
    "qmpc_solve_batch_sched_host", "qmpc_solve_batch_convex_sched_host", "qmpc_default_leg_params", "qmpc_leg_kinematics", "qmpc_joint_torques", "qmpc_describe", "qmpc_solve_batch_warm", "qmpc_goal_state_bytes", "qmpc_goal_update",
    "qmpc_default_raibert_params", "qmpc_raibert_targets",
    "qmpc_create_ex", "qmpc_create_multi", "qmpc_solve_batch_host_multi", "qmpc_destroy_multi",
    "qmpc_multi_device_count", "qmpc_multi_launch_count", "qmpc_multi_last_error",
    "qmpc_leg_fsm_state_bytes", "qmpc_leg_fsm_init", "qmpc_foot_update",
]

_LIB = None
_XLIB = None


def lib_path(xcheck=False):
    # QMPC_LIB lets an experiment point at an alternative build of the same library
    if xcheck:
        return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libqmpc_b200_xcheck.so")
    return os.environ.get("QMPC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libqmpc_b200.so")


def load_library(xcheck=False):
    """Load libqmpc_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent:
    there is no Python / CPU fallback for the solve.  xcheck=True loads the TEST-ONLY sibling
    libqmpc_b200_xcheck.so, the same library plus the dense / srb cross-check kernels (the product
    library does not contain them)."""
    global _LIB, _XLIB
    if xcheck and _XLIB is not None:
        return _XLIB
    if not xcheck and _LIB is not None:
        return _LIB
    path = lib_path(xcheck)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: the CUDA extension is not built. Run `python -c 'import "
            "__graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.qmpc_default_config.argtypes = [i32, i32, C.POINTER(QmpcConfig)]
    lib.qmpc_default_config.restype = C.c_int
    lib.qmpc_create.argtypes = [C.POINTER(QmpcConfig), i32, i32, C.POINTER(vp)]
    lib.qmpc_create.restype = C.c_int
    lib.qmpc_create_ex.argtypes = [C.POINTER(QmpcConfig), i32, i32, C.POINTER(QmpcCreateOptions), C.POINTER(vp)]
    lib.qmpc_create_ex.restype = C.c_int
    lib.qmpc_create_multi.argtypes = [C.POINTER(QmpcConfig), i32, C.POINTER(i32), i32, C.POINTER(vp)]
    lib.qmpc_create_multi.restype = C.c_int
    lib.qmpc_solve_batch_host_multi.argtypes = [vp, vp, i32, vp]
    lib.qmpc_solve_batch_host_multi.restype = C.c_int
    lib.qmpc_destroy_multi.argtypes = [vp]
    lib.qmpc_destroy_multi.restype = None
    lib.qmpc_multi_device_count.argtypes = [vp]
    lib.qmpc_multi_device_count.restype = i32
    lib.qmpc_multi_launch_count.argtypes = [vp]
    lib.qmpc_multi_launch_count.restype = i64
    lib.qmpc_multi_last_error.argtypes = [vp]
    lib.qmpc_multi_last_error.restype = C.c_char_p
    for name in ("qmpc_solve_batch", "qmpc_solve_batch_convex"):
        f = getattr(lib, name)
        f.argtypes = [vp, vp, i32, vp, vp]
        f.restype = C.c_int
    for name in ("qmpc_solve_batch_host", "qmpc_solve_batch_convex_host"):
        f = getattr(lib, name)
        f.argtypes = [vp, vp, i32, vp]
        f.restype = C.c_int
    for name in ("qmpc_solve_batch_sched", "qmpc_solve_batch_convex_sched"):
        f = getattr(lib, name)
        f.argtypes = [vp, vp, vp, i32, vp, vp]
        f.restype = C.c_int
    lib.qmpc_solve_batch_warm.argtypes = [vp, vp, vp, vp, i32, vp, vp]
    lib.qmpc_solve_batch_warm.restype = C.c_int
    for name in ("qmpc_solve_batch_sched_host", "qmpc_solve_batch_convex_sched_host"):
        f = getattr(lib, name)
        f.argtypes = [vp, vp, vp, i32, vp]
        f.restype = C.c_int
    lib.qmpc_predict_contact_schedule.argtypes = [vp, vp, i32, vp, vp]
    lib.qmpc_predict_contact_schedule.restype = C.c_int
    lib.qmpc_default_leg_params.argtypes = [C.POINTER(QmpcLegParams)]
    lib.qmpc_default_leg_params.restype = C.c_int
    lib.qmpc_leg_kinematics.argtypes = [vp, C.POINTER(QmpcLegParams), vp, i32, vp, vp, vp]
    lib.qmpc_leg_kinematics.restype = C.c_int
    lib.qmpc_joint_torques.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.qmpc_joint_torques.restype = C.c_int
    lib.qmpc_goal_state_bytes.argtypes = [vp]
    lib.qmpc_goal_state_bytes.restype = i64
    lib.qmpc_goal_update.argtypes = [vp, vp, vp, i32, vp, vp]
    lib.qmpc_goal_update.restype = C.c_int
    lib.qmpc_leg_fsm_state_bytes.argtypes = [vp]
    lib.qmpc_leg_fsm_state_bytes.restype = i64
    lib.qmpc_leg_fsm_init.argtypes = [vp, vp, vp, i32, vp]
    lib.qmpc_leg_fsm_init.restype = C.c_int
    lib.qmpc_foot_update.argtypes = [vp, vp, vp, C.c_double, C.c_double, i32, vp, vp, vp, vp]
    lib.qmpc_foot_update.restype = C.c_int
    lib.qmpc_default_raibert_params.argtypes = [C.POINTER(QmpcRaibertParams)]
    lib.qmpc_default_raibert_params.restype = C.c_int
    lib.qmpc_raibert_targets.argtypes = [vp, C.POINTER(QmpcRaibertParams), vp, i32, vp, vp, vp]
    lib.qmpc_raibert_targets.restype = C.c_int
    lib.qmpc_describe.argtypes = [vp, C.c_char_p, i32]
    lib.qmpc_describe.restype = C.c_int
    lib.qmpc_destroy.argtypes = [vp]
    lib.qmpc_destroy.restype = None
    lib.qmpc_launch_count.argtypes = [vp]
    lib.qmpc_launch_count.restype = i64
    lib.qmpc_last_error.argtypes = [vp]
    lib.qmpc_last_error.restype = C.c_char_p
    lib.qmpc_status_string.argtypes = [i32]
    lib.qmpc_status_string.restype = C.c_char_p
    lib.qmpc_measure_fma_peak.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.qmpc_measure_fma_peak.restype = C.c_int
    lib.qmpc_abi_version.argtypes = []
    lib.qmpc_abi_version.restype = i32
    if xcheck:
        _XLIB = lib
    else:
        _LIB = lib
    return lib
