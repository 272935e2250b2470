"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo processes shard a batch, each "solves" its
shard (with the CPU oracle standing in for the GPU, this is a test), rank 0 gathers with the single
collective and must obtain exactly the single-process result in batch order."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch():
    from quaternion_mpc_b200.sharding import shard_range, shard_sizes
    for batch in (0, 1, 7, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            s = shard_sizes(batch, world)
            assert max(s) - min(s) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, batch, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import binding as oracle
    from quaternion_mpc_b200.config import default_config
    from quaternion_mpc_b200.sharding import gather_results, shard_range
    from quaternion_mpc_b200.workloads import random_batch
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = default_config(0, 6)
    cfg.iterations_max = 3
    probs = random_batch(batch, seed=9, gait="mixed")
    lo, hi = shard_range(batch, rank, world)
    local = oracle.solve_batch(cfg, probs[lo:hi])
    full = gather_results(local, batch)
    if rank == 0:
        ref = oracle.solve_batch(cfg, probs)
        q.put(bool(full.tobytes() == ref.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 37), (3, 10)])
def test_gloo_shard_solve_gather(world, batch, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
