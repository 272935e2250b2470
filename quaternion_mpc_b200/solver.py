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

    def __init__(self, horizon=10, max_batch=4096, device=0, cfg=None):
        self.lib = abi.load_library()
        if self.lib.qmpc_abi_version() != 1:
            raise QmpcError("libqmpc_b200.so ABI mismatch")
        self.cfg = cfg if cfg is not None else default_config(self.MODEL, horizon)
        if self.cfg.model != self.MODEL and not (
                self.MODEL == abi.QMPC_MODEL_QUAT_4FOOT and self.cfg.model == abi.QMPC_MODEL_QUAT_2FOOT):
            raise ValueError("config model does not match this class")
        self.device = int(device)
        self.max_batch = int(max_batch)
        self._h = C.c_void_p()
        rc = self.lib.qmpc_create(C.byref(self.cfg), self.max_batch, self.device, C.byref(self._h))
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

    def grf_update_host_ptr(self, in_ptr, batch, out_ptr):
        """Same with raw (e.g. pinned) host pointers — used by bench.py's e2e leg."""
        self._check(getattr(self.lib, self._solve_host)(self._h, in_ptr, batch, out_ptr))

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


class ConvexMpc(_BatchedMpc):
    """legged::ConvexMpc (12-state Euler SRB, LQR cost, <=5 iterations)."""
    MODEL = abi.QMPC_MODEL_EULER_CONVEX
    PROBLEM_DTYPE = abi.CONVEX_PROBLEM_DTYPE
    _solve_dev = "qmpc_solve_batch_convex"
    _solve_host = "qmpc_solve_batch_convex_host"
