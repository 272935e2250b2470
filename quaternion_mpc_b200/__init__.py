"""quaternion_mpc_b200 — batched B200 (sm_100a) solver for legged_ctrl's QuatMpc / ConvexMpc GRF solve.

Product = quaternion_mpc_b200/libqmpc_b200.so (hand-written CUDA behind the C-ABI of
include/qmpc.h).  This package is the thin host-side mirror of the reference's MPC classes.
"""
from . import abi
from .config import default_config
from .solver import ConvexMpc, MultiGpuMpc, QmpcError, QuatMpc

__all__ = ["abi", "default_config", "QuatMpc", "ConvexMpc", "MultiGpuMpc", "QmpcError"]
