#!/usr/bin/env python
"""bench.py — Go1 quaternion-MPC solves/s on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # CPU arm: the fp64 oracle port on all host cores
  torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, batch sharded, no collective
                                                           # on the data path (one NCCL gather of the GRFs)

A "step" is one batched solve of the workload (default: BASELINE configs[1] = batch 4096 per GPU,
Go1, horizon 10, trot masks {1001,0110}, random states, seed 0).  `value` = solves/s with the
problem batch already resident in HBM (CUDA events, L2 flushed between steps); `e2e` = the same
through the host-buffer C-ABI call (pinned host memory, H2D + solve + D2H inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_KNOT_ITER = 65.8e3   # dense reference-equivalent AL-iLQR, ne=12, m=12, p=24
IN_BYTES, OUT_BYTES = 296, 240  # sizeof(QmpcProblem), sizeof(QmpcResult)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # written by tools/ncu_traffic.py from an ncu capture
ORACLE_PIN = ("oracle pinned on the reference's inactive-cone goldens (quat_mpc_test.json u0 to 5e-7 N, trot_quat_mpc_test.json) "
              "+ ALTRO's 3/5-iteration toy KATs + an independent SLSQP fixed point; the capped active-cone iterate of the real "
              "ALTRO fork is unpinned (no reference vector exists)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU per step")
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--gait", default="trot")
    ap.add_argument("--model", default="quat", choices=["quat", "quat2", "convex"],
                    help="quat = QuatMpc (BASELINE metric); quat2 = 2-contact model (config 4); convex = ConvexMpc")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the streaming kernels either side of the solve")
    ap.add_argument("--no-config1", action="store_true", help="skip the BASELINE config-1 latency block")
    ap.add_argument("--kernel", default="auto", choices=["auto", "coop", "phased", "srb", "dense"],
                    help="QmpcCreateOptions.kernel (auto = the product default)")
    ap.add_argument("--host-chunks", type=int, default=0,
                    help="QmpcCreateOptions.host_chunks (2..4 = chunked copy / solve pipeline of the host entry points; default off)")
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def mark(self):
        """Start of the window whose samples are reported (the sampler itself is started earlier: nvidia-smi
        needs ~100 ms before its first line, longer than a whole default run's timed region)."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = getattr(self, "t_mark", 0.0)
        rows = [r[1:] for r in self.rows if r[0] >= t0] or [r[1:] for r in self.rows[-3:]]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_arm(cfg, probs, threads):
    """Times the fp64 oracle port (oracle/, test infrastructure) on the host cores."""
    from oracle import binding as oracle
    from quaternion_mpc_b200 import abi
    solve = oracle.solve_batch_convex if cfg.model == abi.QMPC_MODEL_EULER_CONVEX else oracle.solve_batch
    solve(cfg, probs[:min(64, len(probs))], nthreads=threads)  # warm the caches / pages
    t0 = time.perf_counter()
    out = solve(cfg, probs, nthreads=threads)
    dt = time.perf_counter() - t0
    return len(probs) / dt, dt, out


def aux_kernels(device, cfg, hbm_peak, robots=1 << 20, reps=20):
    """The HBM-bound streaming kernels either side of the solve, timed with CUDA events
    on 2^20 robots (inputs resident, > L2): achieved GB/s of the ALGORITHMIC bytes against the HBM peak."""
    import torch
    from quaternion_mpc_b200 import QuatMpc, abi
    from quaternion_mpc_b200.workloads import random_gait_states
    mpc = QuatMpc(max_batch=robots, device=device, cfg=cfg)
    dev = f"cuda:{device}"
    g = torch.from_numpy(random_gait_states(robots, seed=0).view(np.uint8).reshape(robots, -1)).to(dev)
    q = (torch.rand((robots, 12), dtype=torch.float64, device=dev) - 0.5)
    gin = np.zeros(robots, dtype=abi.GOAL_INPUT_DTYPE)
    gin["torso_quat"][:, 0] = 1.0
    gin["joy_vel"][:, 0] = 0.3
    gin["torso_pos_world"][:, 2] = 0.3
    d_gin = torch.from_numpy(gin.view(np.uint8).reshape(robots, -1)).to(dev)
    d_probs = torch.zeros((robots, abi.PROBLEM_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_res = torch.zeros((robots, abi.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_state = mpc.alloc_goal_state()
    foot, jac = mpc.leg_kinematics(q)
    pc = torch.ones((robots, 4), dtype=torch.int32, device=dev)
    d_fsm = mpc.alloc_leg_fsm()
    fin = np.zeros(robots, dtype=abi.FOOT_UPDATE_INPUT_DTYPE)
    fin["movement_mode"] = 1
    fin["foot_pos_target_world"] = 0.05
    d_fin = torch.from_numpy(fin.view(np.uint8).reshape(robots, -1)).to(dev)
    cases = {
        # bytes per robot: what the kernel must read + write (struct sizes from include/qmpc.h)
        "predict_contact_schedule": (48 + 32, lambda: mpc.predict_contact_schedule(g)),
        "leg_kinematics": (96 + 96 + 288, lambda: mpc.leg_kinematics(q)),
        "joint_torques": (96 + 288 + 16 + 96, lambda: mpc.joint_torques(d_res, jac, pc)),
        "goal_update": (128 + 5 * 16 + 6 * (2 * 16 + 16) + 128, lambda: mpc.goal_update(d_state, d_gin, d_probs)),
        "raibert_targets": (128 + 192, lambda: mpc.raibert_targets(d_gin)),
        # QuatMpc::foot_update: input + output records, the 27-field FSM state of 4 legs read and written
        "foot_update": (224 + 336 + 2 * 27 * 4 * 8, lambda: mpc.foot_update(d_fsm, d_fin)),
    }
    out = {"robots": robots}
    for name, (bpr, fn) in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        gbs = robots * bpr / (us * 1e-6) / 1e9
        out[name] = {"us_per_launch": us, "bytes_per_robot": bpr, "achieved_gbs": gbs, "frac_hbm": gbs / hbm_peak}
    mpc.close()
    return out


def workload_of(a, world):
    """Model config, problem generator, workload name and the `config` dict - IDENTICAL in both arms."""
    from quaternion_mpc_b200 import abi, workloads
    from quaternion_mpc_b200.config import default_config
    flops, in_bytes = FLOPS_PER_KNOT_ITER, 296
    if a.model == "quat":
        cfg = default_config(abi.QMPC_MODEL_QUAT_4FOOT, a.horizon)
        gen = workloads.random_batch
        name = f"go1_quat_mpc_N{a.horizon}_{a.gait}_batch{a.batch}_per_gpu_seed0"
    elif a.model == "quat2":   # BASELINE config 4: 2-contact model (ct_srb_trot_quat_*), m = 6, 12 cone rows
        cfg = default_config(abi.QMPC_MODEL_QUAT_2FOOT, a.horizon)
        gen = lambda n, seed=0, gait=None: workloads.random_batch(n, seed=seed, gait="stand", max_angle=0.2, nfeet=2)
        name = f"two_contact_quat_mpc_N{a.horizon}_batch{a.batch}_per_gpu_seed0"
        flops = 38e3           # SURVEY.md 8d, m = 6, p = 12
    else:                      # ConvexMpc (row A8)
        cfg = default_config(abi.QMPC_MODEL_EULER_CONVEX, a.horizon)
        gen = lambda n, seed=0, gait="trot": workloads.random_convex_batch(n, seed=seed, gait=gait)
        name = f"go1_convex_mpc_N{a.horizon}_{a.gait}_batch{a.batch}_per_gpu_seed0"
        in_bytes = 344
    config = {"workload": name, "model": a.model, "horizon": a.horizon, "gait": a.gait, "batch_per_gpu": a.batch,
              "global_batch": a.batch * world, "iterations_max": cfg.iterations_max,
              "problems": "rank r solves random_batch(batch_per_gpu, seed=r); the CPU arm solves the same global batch",
              "l2": "GPU arm: 256 MiB flush between timed steps; CPU arm: not applicable",
              "parallelism": f"batch-sharded x{world}"}
    return cfg, gen, config, flops, in_bytes


def parity_block(res, ref, tol=1e-4):
    """Scope of the parity claim of this very run (GPU results vs the oracle on the same problems): how many solves
    converged / stopped at the cap / were flagged, how many were compared, how many disagree.  Nothing is dropped:
    flagged solves (either side reports linesearch_failed / backward_failed) are compared like the others and
    counted separately."""
    err = np.maximum(np.abs(res["grf_body"] - ref["grf_body"]).max(axis=1), np.abs(res["grf_world"] - ref["grf_world"]).max(axis=1))
    flagged = (res["status"] >= 2) | (ref["status"] >= 2)
    agree = (res["status"] == ref["status"]) & (res["iterations"] == ref["iterations"]) & (err < tol)
    return {"compared": int(len(res)), "converged": int((ref["status"] == 0).sum()), "capped": int((ref["status"] == 1).sum()),
            "flagged": int(flagged.sum()), "flagged_agree": int((flagged & agree).sum()),
            "flagged_differ": int((flagged & ~agree).sum()), "disagree": int((~flagged & ~agree).sum()),
            "max_err": float(err[agree].max()) if agree.any() else None, "max_err_all": float(np.nanmax(err)),
            "tolerance_N": tol, "status_equal": int((res["status"] == ref["status"]).sum()),
            "iterations_equal": int((res["iterations"] == ref["iterations"]).sum()), "oracle_vs_altro": ORACLE_PIN}


def config1_block(mpc_cls, device):
    """BASELINE config 1: one Go1 QuatMpc solve, N=10, stand.  Single-thread CPU latency of the
    oracle port (median / p99 over 1000 solves), the B200 batch-1 latency through the host entry point the ROS shim
    calls, and the reference's implied budget: the mpc_thread period of 5 ms = 200 Hz (Main.cpp:115)."""
    from oracle import binding as oracle
    from quaternion_mpc_b200 import abi
    from quaternion_mpc_b200.config import default_config
    from quaternion_mpc_b200.workloads import stand_problem
    cfg = default_config(abi.QMPC_MODEL_QUAT_4FOOT, 10)
    p = stand_problem()
    for _ in range(20):
        oracle.solve_batch(cfg, p)
    t = []
    for _ in range(1000):
        t0 = time.perf_counter()
        oracle.solve_batch(cfg, p)
        t.append(time.perf_counter() - t0)
    t = np.array(t) * 1e6
    mpc = mpc_cls(max_batch=1, device=device, cfg=cfg)
    out = np.empty(1, dtype=abi.RESULT_DTYPE)
    for _ in range(20):
        mpc.grf_update(p, out)
    g = []
    for _ in range(500):
        t0 = time.perf_counter()
        mpc.grf_update(p, out)
        g.append(time.perf_counter() - t0)
    g = np.array(g) * 1e6
    mpc.close()
    return {"workload": "single Go1 QuatMpc solve, N=10, stand (BASELINE configs[0])",
            "cpu_single_thread_us": {"median": float(np.median(t)), "p99": float(np.percentile(t, 99)), "solves": 1000,
                                     "kind": "port"},
            "gpu_batch1_us": {"median": float(np.median(g)), "p99": float(np.percentile(g, 99)), "solves": 500,
                              "path": "qmpc_solve_batch_host(batch=1), pageable buffers, pinned staging inside"},
            "budget_us": 5000.0, "budget_source": "mpc_thread period 5 ms = 200 Hz (legged_ctrl/src/Main.cpp:115)"}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    from quaternion_mpc_b200 import abi
    cfg, random_batch, config, flops_per_knot_iter, IN_BYTES = workload_of(a, world)
    threads = os.cpu_count() or 1
    B = a.batch

    # ------------------------------------------------------------------ reference (CPU) arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        probe, _, _ = cpu_arm(cfg, random_batch(threads * 16, seed=0, gait=a.gait), threads)
        # one step = the GPU arm's global batch (the same problems, rank by rank) when that is <= ~10 s of host
        # work; else a bounded sample of it
        full = B * world
        n = a.cpu_sample or (full if full <= probe * 10.0 else int(max(probe * 4.0, 512)))
        n = min(n, full)
        parts, left, r = [], n, 0
        while left > 0:
            m = min(B, left)
            parts.append(random_batch(B, seed=r, gait=a.gait)[:m])
            left -= m
            r += 1
        probs = np.concatenate(parts)
        vals = []
        for i in range(a.warmup + a.steps):
            v, dt, _ = cpu_arm(cfg, probs, threads)
            if i >= a.warmup:
                vals.append((v, dt))
        v = float(np.mean([x[0] for x in vals]))
        sample = (f"the whole global batch ({n} problems) per step" if n == full else
                  f"the first {n} of the {full} problems of the global batch per step") + f", {threads} pthreads"
        line = {
            "impl": "reference", "metric": "go1_quat_mpc_solves_per_sec", "value": v, "unit": "solves/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": float(np.mean([x[1] for x in vals]) * 1e3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "note": "reference CPU solver = fp64 oracle port (ALTRO/Eigen/ROS absent: reference unbuildable here)",
            "cpu_baseline": {"value": v, "unit": "solves/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the product path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        # host-side barrier for the one-call multi-GPU e2e leg: the other ranks must wait WITHOUT a kernel on their
        # GPU (an NCCL barrier spins on the device and takes SMs from the solve rank 0 launches there)
        cpu_group = dist.new_group(backend="gloo")
    from quaternion_mpc_b200 import ConvexMpc, MultiGpuMpc, QuatMpc
    Mpc = ConvexMpc if a.model == "convex" else QuatMpc

    probs = random_batch(B, seed=0 + rank, gait=a.gait)   # each rank owns its shard of the global batch
    mpc = Mpc(max_batch=B, device=local, cfg=cfg, kernel=a.kernel, host_chunks=a.host_chunks)
    d_in = mpc.to_device(probs)
    d_out = mpc.alloc_results(B)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()      # before the warm-up: nvidia-smi takes ~100 ms to produce its first sample
    for _ in range(a.warmup):
        flush.fill_(1)
        mpc.grf_update_device(d_in, d_out)
    barrier()
    launches0 = mpc.launch_count
    sampler.mark()           # samples from here (timed region + the e2e loop, GPU busy throughout) are reported
    evs = []
    for _ in range(a.steps):
        flush.fill_(1)                        # L2 flush between timed iterations (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mpc.grf_update_device(d_in, d_out)
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    kernel_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = float(sum(kernel_ms))
    launches = mpc.launch_count - launches0

    # ---- the one collective of the path, for device-resident callers: every rank's results gathered
    # on rank 0 over NCCL, device to device, timed on the device (the solve above is timed without it)
    gather_ms = None
    if world > 1:
        outs = [torch.empty_like(d_out) for _ in range(world)] if rank == 0 else None
        for i in range(3 + 5):
            if i == 3:
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
            dist.gather(d_out, outs, dst=0)
        g1.record(stream)
        barrier()
        gather_ms = g0.elapsed_time(g1) / 5
        if rank == 0:
            assert all(torch.equal(outs[0], d_out) if r == 0 else outs[r].shape == d_out.shape for r in range(world))

    # ---- e2e: HOST buffers through the C-ABI (H2D + solve + D2H inside the call, per step).
    #   1 GPU : qmpc_solve_batch_host on this rank's handle
    #   N GPUs: ONE qmpc_solve_batch_host_multi call on rank 0 - the whole global batch in one pinned host array,
    #           sharded over the N devices inside the C-ABI, every result back in one host array (the other ranks
    #           wait at the barrier; their GPUs are driven by rank 0's call)
    res = mpc.results_to_numpy(d_out)
    if world == 1:
        h_in = torch.from_numpy(probs.view(np.uint8).reshape(B, -1).copy()).pin_memory()
        h_out = torch.empty((B, OUT_BYTES), dtype=torch.uint8).pin_memory()
        for _ in range(2):
            mpc.grf_update_host_ptr(h_in.data_ptr(), B, h_out.data_ptr())
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            mpc.grf_update_host_ptr(h_in.data_ptr(), B, h_out.data_ptr())
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_path = "qmpc_solve_batch_host, pinned host buffers" + (f", up to {a.host_chunks} chunks" if a.host_chunks > 1 else "")
        assert h_out.numpy().tobytes() == res.tobytes()
    else:
        e2e_s = 0.0
        e2e_path = f"one qmpc_solve_batch_host_multi call on rank 0 over {world} devices, pinned host buffers"
        barrier()
        dist.barrier(group=cpu_group)
        if rank == 0:
            allp = np.concatenate([random_batch(B, seed=r, gait=a.gait) for r in range(world)])
            multi = MultiGpuMpc(cfg, B * world, list(range(world)))
            h_in = torch.from_numpy(allp.view(np.uint8).reshape(B * world, -1).copy()).pin_memory()
            h_out = torch.empty((B * world, OUT_BYTES), dtype=torch.uint8).pin_memory()
            for _ in range(2):
                multi.grf_update_host_ptr(h_in.data_ptr(), B * world, h_out.data_ptr())
            t0 = time.perf_counter()
            for _ in range(a.steps):
                multi.grf_update_host_ptr(h_in.data_ptr(), B * world, h_out.data_ptr())
            e2e_s = time.perf_counter() - t0
            allres = h_out.numpy().reshape(-1).view(abi.RESULT_DTYPE)
            assert allres[:B].tobytes() == res.tobytes() and np.isfinite(allres["grf_body"]).all()
            multi.close()
        dist.barrier(group=cpu_group)      # ranks != 0 wait here on the host, their GPUs idle for rank 0's call
        barrier()
    if rank == 0 and len([r for r in sampler.rows if r[0] >= sampler.t_mark]) < 3:
        t_end = time.perf_counter() + 0.25          # very short runs: keep the GPU busy with the same solve
        while time.perf_counter() < t_end:          # until a few samples under load exist (not timed)
            mpc.grf_update_device(d_in, d_out)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    mean_iters = float(res["iterations"].mean())

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_s * 1e3, mean_iters, gather_ms or 0.0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms, gather_ms = float(tmax[0]), float(tmax[1]), float(tmax[3])
        mean_iters = float(tsum[2]) / world
    else:
        e2e_ms = e2e_s * 1e3
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    solves = B * world * a.steps
    value = solves / (total_ms * 1e-3)
    e2e_value = solves / (e2e_ms * 1e-3)

    # ---- roofline: this path is bound by the vector-FMA pipe, not HBM/tensor
    f64, f32 = C.c_double(), C.c_double()
    mpc.lib.qmpc_measure_fma_peak(local, C.byref(f64), C.byref(f32))
    flops_per_solve = mean_iters * a.horizon * flops_per_knot_iter
    avg_step_s = (total_ms * 1e-3) / a.steps
    achieved_tflops = B * flops_per_solve / avg_step_s / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_achieved = B * (IN_BYTES + OUT_BYTES) / avg_step_s / 1e9
    # DRAM bytes actually moved: dram__bytes_read.sum + dram__bytes_write.sum from an `ncu` capture of this kernel at
    # this workload (profiles/ncu_traffic.json names the capture, its commit and its batch); null without a capture
    traffic, traffic_src = None, "no ncu capture for this workload / kernel in profiles/ncu_traffic.json"
    try:
        for rec in json.load(open(TRAFFIC_FILE)):
            if rec["model"] == a.model and rec["horizon"] == a.horizon and rec["kernel"] in mpc.describe():
                traffic = rec["dram_bytes_per_solve"] * B
                traffic_src = (f"ncu capture {rec['capture']} (commit {rec['commit']}, batch {rec['batch']}): "
                               f"{rec['dram_bytes_per_solve']:.0f} DRAM bytes per solve, scaled to this launch's batch")
    except Exception:
        pass
    roofline = {
        "bound": "fp64_fma", "achieved": achieved_tflops, "peak": f64.value, "unit": "TFLOP/s",
        "frac": achieved_tflops / f64.value if f64.value > 0 else None,
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": "measured live by qmpc_measure_fma_peak (FP64 vector FMA; FP32 = %.1f TFLOP/s)" % f32.value,
        "algorithmic_flops_per_solve": flops_per_solve, "mean_iterations": mean_iters,
        "launches_per_step": launches / a.steps,
        "hbm": {"achieved_gbs": hbm_achieved, "peak_gbs": hbm_peak, "frac": hbm_achieved / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_solve": IN_BYTES + OUT_BYTES},
    }

    aux = None
    if not a.no_aux and a.model == "quat":
        aux = aux_kernels(local, cfg, hbm_peak)

    cpu, parity = None, None
    if not a.no_cpu_baseline:
        # bounded sample: probe the host rate on 16 problems per core, then time ~8 s worth of the
        # same distribution (the first B problems are exactly rank 0's GPU batch -> parity block)
        probe, _, _ = cpu_arm(cfg, probs[:min(B, threads * 16)], threads)
        n = a.cpu_sample or int(min(max(probe * 8.0, 512), 1 << 17))
        cprobs = probs if n <= B else np.concatenate([probs, random_batch(n - B, seed=12345, gait=a.gait)])
        v, dt, ref = cpu_arm(cfg, cprobs[:n], threads)
        m = min(n, B)
        parity = parity_block(res[:m], ref[:m])
        cpu = {"value": v, "unit": "solves/s", "cores": threads, "kind": "port",
               "sample": f"{n} problems of the workload distribution (first {m} = rank 0's GPU batch), "
                         f"{threads} pthreads, {dt:.2f} s"}
    config1 = None
    if not a.no_config1 and a.model == "quat" and not a.no_cpu_baseline:
        config1 = config1_block(QuatMpc, local)

    line = {
        "metric": "go1_quat_mpc_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config, "kernel": mpc.describe(),
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": B * world * IN_BYTES,
                "d2h_bytes_per_step": B * world * OUT_BYTES, "path": e2e_path},
        "gather": None if gather_ms is None else {"ms": gather_ms, "bytes": B * world * OUT_BYTES,
                                                  "what": "torch.distributed.gather of the device result tensors on rank 0 (NCCL)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "config1": config1, "aux_kernels": aux,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
