#!/usr/bin/env python
"""On-device bit-identity check between builds of the library: each build solves the same workloads in its own
process (QMPC_LIB selects the .so) and dumps the raw result bytes; this script compares them with the first build.
usage: gpu_bitcheck.py base.so other.so ...   |   gpu_bitcheck.py --dump out.npz   (child mode)"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def dump(path):
    import torch
    from quaternion_mpc_b200 import QuatMpc, ConvexMpc, abi
    from quaternion_mpc_b200.config import default_config
    from quaternion_mpc_b200.workloads import random_batch, random_convex_batch, random_gait_states, predict_schedule_numpy
    out = {}
    for name, model, N, B, kw in [("quat_N10_trot", 0, 10, 4096, dict(gait="trot", seed=0)),
                                  ("quat_N16_mixed", 0, 16, 2048, dict(gait="mixed", seed=1)),
                                  ("quat2_N20", 1, 20, 1024, dict(gait="stand", seed=2, nfeet=2, max_angle=0.2)),
                                  ("convex_N10", 2, 10, 2048, dict(seed=3))]:
        cfg = default_config(model, N)
        cls = ConvexMpc if model == 2 else QuatMpc
        p = random_convex_batch(B, **kw) if model == 2 else random_batch(B, **kw)
        mpc = cls(horizon=N, max_batch=B, device=0, cfg=cfg)
        r = mpc.results_to_numpy(mpc.grf_update_device(mpc.to_device(p)))
        out[name] = r.view(np.uint8).copy()
        if model == 0 and N == 10:
            sched = predict_schedule_numpy(random_gait_states(B, seed=5), N, cfg.dt)
            warm = mpc.alloc_warm(B)
            d_s = mpc.schedule_to_device(sched)
            for tick in range(2):
                r = mpc.results_to_numpy(mpc.grf_update_warm_device(mpc.to_device(p), warm, d_s))
                out[f"{name}_sched_warm{tick}"] = r.view(np.uint8).copy()
        mpc.close()
    torch.cuda.synchronize()
    np.savez(path, **out)


def main():
    if sys.argv[1] == "--dump":
        return dump(sys.argv[2])
    tmp = tempfile.mkdtemp()
    res = []
    for i, lib in enumerate(sys.argv[1:]):
        f = os.path.join(tmp, f"r{i}.npz")
        env = dict(os.environ, QMPC_LIB=os.path.abspath(lib))
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--dump", f], env=env)
        res.append(dict(np.load(f)))
    bad = 0
    for i, lib in enumerate(sys.argv[2:], 1):
        for k in res[0]:
            same = np.array_equal(res[0][k], res[i][k])
            if not same:
                bad += 1
                from quaternion_mpc_b200 import abi
                a, b = res[0][k].view(abi.RESULT_DTYPE), res[i][k].view(abi.RESULT_DTYPE)
                d = np.abs(a["grf_body"] - b["grf_body"])
                print(f"{lib} {k}: DIFFERENT  max|dGRF|={np.nanmax(d):.3e} solves differing={(d.max(1) > 0).sum()} "
                      f"iters_equal={(a['iterations'] == b['iterations']).all()}")
        print(f"{lib}: compared {len(res[0])} workloads against {sys.argv[1]}")
    print("GPU bit-identical" if bad == 0 else f"{bad} workload results differ")


if __name__ == "__main__":
    main()
