"""Batch-axis sharding of independent MPC problems over the GPUs of one node.

Every solve is independent (the reference builds a fresh solver per call, QuatMpc.cpp:218), so the
only multi-GPU strategy is: rank r owns the contiguous range [r*B/W, (r+1)*B/W) of the global
batch, solves it on its own GPU with no data-path collective, and the GRFs are gathered on rank 0
with ONE collective (torch.distributed.gather over NCCL/NVLink; gloo in the CPU tests).
"""
import numpy as np

from . import abi


def shard_range(batch, rank, world):
    """Contiguous, balanced (sizes differ by at most 1), order-preserving partition."""
    if not (0 <= rank < world) or batch < 0:
        raise ValueError("bad shard arguments")
    lo = (batch * rank) // world
    hi = (batch * (rank + 1)) // world
    return lo, hi


def shard_sizes(batch, world):
    return [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]


def gather_results(local_results, batch, device=None, dst=0):
    """Gather per-rank RESULT_DTYPE arrays (rank order = batch order) on `dst`.

    One collective: ranks pad their shard to the largest shard size, `dist.gather` moves the raw
    result bytes, rank `dst` trims the padding and returns the global array (others return None).
    """
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(batch, world)
    assert len(local_results) == sizes[rank]
    width = abi.RESULT_DTYPE.itemsize
    pad = max(sizes)
    buf = np.zeros((pad, width), dtype=np.uint8)
    buf[:sizes[rank]] = np.ascontiguousarray(local_results).view(np.uint8).reshape(-1, width)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, outs, dst=dst)
    if rank != dst:
        return None
    parts = [o.cpu().numpy()[:sizes[r]].reshape(-1).view(abi.RESULT_DTYPE) for r, o in enumerate(outs)]
    return np.concatenate(parts)
