"""Host-side Python mirror of the reference's MPC objects, on top of the C-ABI (libqmpc_b200.so).

`QuatMpc` / `ConvexMpc` mirror legged::QuatMpc / legged::ConvexMpc
(legged_ctrl/include/mpc/QuatMpc.h, ConvexMpc.h): constructed once from the parameters, then
`grf_update(problems)` is the batched equivalent of `grf_update(LeggedState&)`
(QuatMpc.cpp:109-276, ConvexMpc.cpp:81-198).  PyTorch is used only for device memory and
streams; all arithmetic happens inside the CUDA library.  If the library or a CUDA device is
missing the constructor raises — there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import abi
from .config import default_config


class QmpcError(RuntimeError):
    pass


class _BatchedMpc:
    MODEL = None
    PROBLEM_DTYPE = None
    _solve_dev = None
    _solve_host = None
    _solve_sched_dev = None
    _solve_sched_host = None

    def __init__(self, horizon=10, max_batch=4096, device=0, cfg=None, kernel="auto", smem_residents=-1,
                 packed_launch=False, host_chunks=0):
        """`kernel` / `smem_residents` / `packed_launch` / `host_chunks` are the QmpcCreateOptions of include/qmpc.h
        (explicit per-handle choices; the library reads no environment variables).  kernel "dense" / "srb"
        select the on-device cross-check kernels used by the tests."""
        # the dense / srb cross-check kernels live in the test-only sibling library, not in the product
        self.lib = abi.load_library(xcheck=kernel in ("dense", "srb"))
        if self.lib.qmpc_abi_version() != abi.QMPC_ABI_VERSION:
            raise QmpcError("libqmpc_b200.so ABI mismatch")
        self.cfg = cfg if cfg is not None else default_config(self.MODEL, horizon)
        if self.cfg.model != self.MODEL and not (
                self.MODEL == abi.QMPC_MODEL_QUAT_4FOOT and self.cfg.model == abi.QMPC_MODEL_QUAT_2FOOT):
            raise ValueError("config model does not match this class")
        self.device = int(device)
        self.max_batch = int(max_batch)
        self._h = C.c_void_p()
        opt = abi.QmpcCreateOptions(abi.KERNEL_NAMES[kernel], int(smem_residents), int(bool(packed_launch)), int(host_chunks))
        rc = self.lib.qmpc_create_ex(C.byref(self.cfg), self.max_batch, self.device, C.byref(opt), C.byref(self._h))
        if rc != abi.QMPC_OK:
            msg = self.lib.qmpc_last_error(self._h).decode() if self._h else ""
            self.close()
            raise QmpcError(f"qmpc_create failed rc={rc} {msg} (no CPU fallback exists)")

    # -- device path: torch uint8 tensors holding PROBLEM_DTYPE / RESULT_DTYPE bytes -------------
    def to_device(self, problems):
        import torch
        problems = np.ascontiguousarray(problems, dtype=self.PROBLEM_DTYPE)
        t = torch.from_numpy(problems.view(np.uint8).reshape(problems.shape[0], -1))
        return t.to(f"cuda:{self.device}", non_blocking=False)

    def alloc_results(self, batch):
        import torch
        return torch.empty((batch, abi.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=f"cuda:{self.device}")

    def grf_update_device(self, d_problems, d_results=None, stream=None):
        """Enqueue one batched solve on the current torch stream; returns the result tensor."""
        import torch
        batch = d_problems.shape[0]
        assert d_problems.is_cuda and d_problems.dtype == torch.uint8 and d_problems.is_contiguous()
        assert d_problems.shape[1] == self.PROBLEM_DTYPE.itemsize
        if d_results is None:
            d_results = self.alloc_results(batch)
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        rc = getattr(self.lib, self._solve_dev)(self._h, d_problems.data_ptr(), batch, d_results.data_ptr(), s)
        self._check(rc)
        return d_results

    # -- per-step contact schedules ------------------------------------------------
    def schedule_to_device(self, schedule):
        """(batch, QMPC_MAX_HORIZON) uint8 contact masks -> device tensor of QmpcContactSchedule."""
        import torch
        schedule = np.ascontiguousarray(schedule, dtype=np.uint8)
        assert schedule.ndim == 2 and schedule.shape[1] == abi.QMPC_MAX_HORIZON
        return torch.from_numpy(schedule).to(f"cuda:{self.device}")

    def predict_contact_schedule(self, d_gait, d_sched=None, stream=None):
        """Batched LeggedContactFSM::predict_contact_state at t + k*dt (LeggedContactFSM.cpp:272-286).
        d_gait: uint8 device tensor of GAIT_STATE_DTYPE records."""
        import torch
        batch = d_gait.shape[0]
        assert d_gait.is_cuda and d_gait.dtype == torch.uint8 and d_gait.shape[1] == abi.GAIT_STATE_DTYPE.itemsize
        if d_sched is None:
            d_sched = torch.empty((batch, abi.QMPC_MAX_HORIZON), dtype=torch.uint8, device=d_gait.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.qmpc_predict_contact_schedule(self._h, d_gait.data_ptr(), batch, d_sched.data_ptr(), s))
        return d_sched

    def grf_update_sched_device(self, d_problems, d_sched, d_results=None, stream=None):
        """grf_update_device with one contact mask per knot (d_sched from schedule_to_device /
        predict_contact_schedule; None = the reference's constant mask)."""
        import torch
        batch = d_problems.shape[0]
        assert d_problems.is_cuda and d_problems.dtype == torch.uint8 and d_problems.is_contiguous()
        assert d_problems.shape[1] == self.PROBLEM_DTYPE.itemsize
        if d_sched is not None:
            assert d_sched.is_cuda and d_sched.dtype == torch.uint8 and d_sched.is_contiguous()
            assert tuple(d_sched.shape) == (batch, abi.QMPC_MAX_HORIZON)
        if d_results is None:
            d_results = self.alloc_results(batch)
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        rc = getattr(self.lib, self._solve_sched_dev)(self._h, d_problems.data_ptr(),
                                                      d_sched.data_ptr() if d_sched is not None else None,
                                                      batch, d_results.data_ptr(), s)
        self._check(rc)
        return d_results

    # -- reference generation ------------------------------------------------------
    def alloc_goal_state(self):
        """Device state of QuatMpc::goal_update for max_batch robots (zero = freshly constructed)."""
        import torch
        n = int(self.lib.qmpc_goal_state_bytes(self._h))
        return torch.zeros(n, dtype=torch.uint8, device=f"cuda:{self.device}")

    def goal_update(self, d_goal_state, d_goal_inputs, d_problems, stream=None):
        """One batched QuatMpc::goal_update tick (QuatMpc.cpp:68-107): fills the reference fields of
        d_problems in place from joystick + torso feedback (GOAL_INPUT_DTYPE records, uint8 tensor)."""
        import torch
        batch = d_goal_inputs.shape[0]
        assert d_goal_inputs.is_cuda and d_goal_inputs.dtype == torch.uint8
        assert d_goal_inputs.shape[1] == abi.GOAL_INPUT_DTYPE.itemsize and d_problems.shape[0] == batch
        assert d_problems.shape[1] == self.PROBLEM_DTYPE.itemsize and d_problems.is_contiguous()
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.qmpc_goal_update(self._h, d_goal_state.data_ptr(), d_goal_inputs.data_ptr(), batch,
                                              d_problems.data_ptr(), s))
        return d_problems

    def raibert_targets(self, d_goal_inputs, params=None, stream=None):
        """Raibert foot-hold targets (BaseInterface.cpp:265-288) -> (world (batch,12), rel (batch,12))."""
        import torch
        batch = d_goal_inputs.shape[0]
        if params is None:
            params = abi.QmpcRaibertParams()
            self.lib.qmpc_default_raibert_params(C.byref(params))
        tw = torch.empty((batch, 12), dtype=torch.float64, device=d_goal_inputs.device)
        tr = torch.empty((batch, 12), dtype=torch.float64, device=d_goal_inputs.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.qmpc_raibert_targets(self._h, C.byref(params), d_goal_inputs.data_ptr(), batch,
                                                  tw.data_ptr(), tr.data_ptr(), s))
        return tw, tr

    # -- gait FSM (second half): QuatMpc::foot_update --------------------------------
    def alloc_leg_fsm(self, d_gait=None):
        """Device state of the four LeggedContactFSM objects of max_batch robots, initialised like
        reset_params + reset (LeggedContactFSM.cpp:4-31); d_gait: int32 (max_batch,) QMPC_GAIT_* or None = trot."""
        import torch
        n = int(self.lib.qmpc_leg_fsm_state_bytes(self._h))
        st = torch.zeros(n, dtype=torch.uint8, device=f"cuda:{self.device}")
        if d_gait is not None:
            assert d_gait.dtype == torch.int32 and d_gait.is_cuda and d_gait.shape[0] == self.max_batch
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.qmpc_leg_fsm_init(self._h, st.data_ptr(), d_gait.data_ptr() if d_gait is not None else None,
                                               self.max_batch, s))
        return st

    def foot_update(self, d_fsm_state, d_inputs, gait_freq=2.2, dt=5.0 / 1000.0, d_problems=None, d_gait_out=None, stream=None):
        """One batched QuatMpc::foot_update tick (QuatMpc.cpp:278-305): FOOT_UPDATE_INPUT_DTYPE records in,
        FOOT_UPDATE_OUTPUT_DTYPE records out; optionally writes plan_contacts into d_problems and the gait
        states predict_contact_schedule reads into d_gait_out."""
        import torch
        batch = d_inputs.shape[0]
        assert d_inputs.is_cuda and d_inputs.dtype == torch.uint8 and d_inputs.shape[1] == abi.FOOT_UPDATE_INPUT_DTYPE.itemsize
        out = torch.empty((batch, abi.FOOT_UPDATE_OUTPUT_DTYPE.itemsize), dtype=torch.uint8, device=d_inputs.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.qmpc_foot_update(self._h, d_fsm_state.data_ptr(), d_inputs.data_ptr(), float(dt), float(gait_freq),
                                              batch, out.data_ptr(),
                                              d_problems.data_ptr() if d_problems is not None else None,
                                              d_gait_out.data_ptr() if d_gait_out is not None else None, s))
        return out

    # -- warm start / trajectory shift ---------------------------------------------
    def alloc_warm(self, batch):
        """Device buffer of `batch` QmpcWarmStart records, all invalid (first solve starts cold)."""
        import torch
        return torch.zeros((batch, abi.WARM_DTYPE.itemsize), dtype=torch.uint8, device=f"cuda:{self.device}")

    def grf_update_warm_device(self, d_problems, d_warm, d_sched=None, d_results=None, stream=None):
        """Solve starting from the previous solution shifted by one knot; d_warm is updated in place."""
        import torch
        batch = d_problems.shape[0]
        assert d_problems.is_cuda and d_problems.dtype == torch.uint8 and d_problems.is_contiguous()
        assert d_warm.is_cuda and d_warm.dtype == torch.uint8 and d_warm.is_contiguous()
        assert tuple(d_warm.shape) == (batch, abi.WARM_DTYPE.itemsize)
        if self.MODEL == abi.QMPC_MODEL_EULER_CONVEX:
            raise QmpcError("warm start is implemented for the quaternion models")
        if d_results is None:
            d_results = self.alloc_results(batch)
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.qmpc_solve_batch_warm(self._h, d_problems.data_ptr(),
                                            d_sched.data_ptr() if d_sched is not None else None,
                                            d_warm.data_ptr(), batch, d_results.data_ptr(), s)
        self._check(rc)
        return d_results

    # -- leg kinematics in / joint torques out -------------------------------------
    def leg_kinematics(self, d_joint_pos, leg_params=None, want_foot=True, want_jac=True, stream=None):
        """a1_kin.fk / a1_kin.jac for the four legs (BaseInterface.cpp:204-212).
        d_joint_pos: (batch, 12) float64 device tensor -> (foot_pos_body (batch,12), jac_foot (batch,36))."""
        import torch
        assert d_joint_pos.is_cuda and d_joint_pos.dtype == torch.float64 and d_joint_pos.is_contiguous()
        batch = d_joint_pos.shape[0]
        if leg_params is None:
            leg_params = abi.QmpcLegParams()
            self.lib.qmpc_default_leg_params(C.byref(leg_params))
        foot = torch.empty((batch, 12), dtype=torch.float64, device=d_joint_pos.device) if want_foot else None
        jac = torch.empty((batch, 36), dtype=torch.float64, device=d_joint_pos.device) if want_jac else None
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.qmpc_leg_kinematics(self._h, C.byref(leg_params), d_joint_pos.data_ptr(), batch,
                                                 foot.data_ptr() if want_foot else None,
                                                 jac.data_ptr() if want_jac else None, s))
        return foot, jac

    def joint_torques(self, d_results, d_jac_foot, d_plan_contacts=None, movement_mode=1, stream=None):
        """ctrl.joint_tau_tgt = -jac^T * optimized_input per leg (BaseInterface.cpp:343-405)."""
        import torch
        batch = d_results.shape[0]
        assert d_jac_foot.dtype == torch.float64 and tuple(d_jac_foot.shape) == (batch, 36)
        if d_plan_contacts is not None:
            assert d_plan_contacts.dtype == torch.int32 and tuple(d_plan_contacts.shape) == (batch, 4)
            assert d_plan_contacts.is_contiguous()
        tau = torch.empty((batch, 12), dtype=torch.float64, device=d_results.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.qmpc_joint_torques(self._h, d_results.data_ptr(), d_jac_foot.data_ptr(),
                                                d_plan_contacts.data_ptr() if d_plan_contacts is not None else None,
                                                int(movement_mode), batch, tau.data_ptr(), s))
        return tau

    @staticmethod
    def results_to_numpy(d_results):
        return d_results.cpu().numpy().reshape(-1).view(abi.RESULT_DTYPE)

    # -- host path: numpy in / numpy out through the host entry point ---------------------------
    def grf_update(self, problems, out=None):
        """Batched grf_update with HOST buffers (H2D, solve, D2H inside the call)."""
        problems = np.ascontiguousarray(problems, dtype=self.PROBLEM_DTYPE)
        if out is None:
            out = np.empty(problems.shape[0], dtype=abi.RESULT_DTYPE)
        rc = getattr(self.lib, self._solve_host)(self._h, problems.ctypes.data, problems.shape[0], out.ctypes.data)
        self._check(rc)
        return out

    def grf_update_sched(self, problems, schedule, out=None):
        """Host-buffer solve with per-knot contact masks ((batch, QMPC_MAX_HORIZON) uint8)."""
        problems = np.ascontiguousarray(problems, dtype=self.PROBLEM_DTYPE)
        schedule = np.ascontiguousarray(schedule, dtype=np.uint8)
        assert schedule.shape == (problems.shape[0], abi.QMPC_MAX_HORIZON)
        if self._solve_sched_host is None:
            raise QmpcError("no host schedule entry point for this model; use grf_update_sched_device")
        if out is None:
            out = np.empty(problems.shape[0], dtype=abi.RESULT_DTYPE)
        rc = getattr(self.lib, self._solve_sched_host)(self._h, problems.ctypes.data, schedule.ctypes.data,
                                                       problems.shape[0], out.ctypes.data)
        self._check(rc)
        return out

    def grf_update_host_ptr(self, in_ptr, batch, out_ptr):
        """Same with raw (e.g. pinned) host pointers — used by bench.py's e2e leg."""
        self._check(getattr(self.lib, self._solve_host)(self._h, in_ptr, batch, out_ptr))

    def describe(self):
        buf = C.create_string_buffer(512)
        self._check(self.lib.qmpc_describe(self._h, buf, 512))
        return buf.value.decode()

    @property
    def launch_count(self):
        return int(self.lib.qmpc_launch_count(self._h))

    def _check(self, rc):
        if rc != abi.QMPC_OK:
            raise QmpcError(f"qmpc solve failed rc={rc}: {self.lib.qmpc_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.qmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class QuatMpc(_BatchedMpc):
    """legged::QuatMpc (13-state quaternion SRB, AL-iLQR, <=10 iterations)."""
    MODEL = abi.QMPC_MODEL_QUAT_4FOOT
    PROBLEM_DTYPE = abi.PROBLEM_DTYPE
    _solve_dev = "qmpc_solve_batch"
    _solve_host = "qmpc_solve_batch_host"
    _solve_sched_dev = "qmpc_solve_batch_sched"
    _solve_sched_host = "qmpc_solve_batch_sched_host"


class ConvexMpc(_BatchedMpc):
    """legged::ConvexMpc (12-state Euler SRB, LQR cost, <=5 iterations)."""
    MODEL = abi.QMPC_MODEL_EULER_CONVEX
    PROBLEM_DTYPE = abi.CONVEX_PROBLEM_DTYPE
    _solve_dev = "qmpc_solve_batch_convex"
    _solve_host = "qmpc_solve_batch_convex_host"
    _solve_sched_dev = "qmpc_solve_batch_convex_sched"
    _solve_sched_host = "qmpc_solve_batch_convex_sched_host"


class MultiGpuMpc:
    """One host batch over several GPUs of the box through ONE C-ABI call (qmpc_create_multi /
    qmpc_solve_batch_host_multi): contiguous balanced shards, a stream and pinned staging per device,
    results in one host array in batch order.  No collective is involved."""

    def __init__(self, cfg, max_batch, devices):
        self.lib = abi.load_library()
        self.cfg = cfg
        self.max_batch = int(max_batch)
        self.devices = list(devices)
        self.PROBLEM_DTYPE = abi.CONVEX_PROBLEM_DTYPE if cfg.model == abi.QMPC_MODEL_EULER_CONVEX else abi.PROBLEM_DTYPE
        self._h = C.c_void_p()
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        rc = self.lib.qmpc_create_multi(C.byref(self.cfg), self.max_batch, arr, len(self.devices), C.byref(self._h))
        if rc != abi.QMPC_OK:
            msg = self.lib.qmpc_multi_last_error(self._h).decode() if self._h else ""
            self.close()
            raise QmpcError(f"qmpc_create_multi failed rc={rc} {msg} (no CPU fallback exists)")

    def grf_update(self, problems, out=None):
        problems = np.ascontiguousarray(problems, dtype=self.PROBLEM_DTYPE)
        if out is None:
            out = np.empty(problems.shape[0], dtype=abi.RESULT_DTYPE)
        self.grf_update_host_ptr(problems.ctypes.data, problems.shape[0], out.ctypes.data)
        return out

    def grf_update_host_ptr(self, in_ptr, batch, out_ptr):
        rc = self.lib.qmpc_solve_batch_host_multi(self._h, in_ptr, batch, out_ptr)
        if rc != abi.QMPC_OK:
            raise QmpcError(f"qmpc_solve_batch_host_multi failed rc={rc}: {self.lib.qmpc_multi_last_error(self._h).decode()}")

    @property
    def launch_count(self):
        return int(self.lib.qmpc_multi_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.qmpc_destroy_multi(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
