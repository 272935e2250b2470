#!/usr/bin/env python
"""Host-side bit-identity check of a kernel refactoring: builds tests/emul from TWO source trees (a saved copy of
the previous sources and the working tree) and requires the cooperative body to return identical bytes on a set of
workloads (quaternion 4-foot N=10/16/32, 2-foot N=20, schedules, warm starts).  usage: emul_bitcheck.py /path/to/old_tree
(old_tree holds csrc/, include/, emul/ as laid out by `cp -r quaternion_mpc_b200/csrc include tests/emul`)."""
import ctypes as C, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quaternion_mpc_b200 import abi
from quaternion_mpc_b200.config import default_config
from quaternion_mpc_b200.workloads import predict_schedule_numpy, random_batch, random_gait_states, random_convex_batch

def build(src_emul, out, extra=()):
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DQMPC_EMUL_SRB", "-ffp-contract=off", *extra,
                           "-o", out, src_emul])
    lib = C.CDLL(out)
    for name in ("emul_solve_coop", "emul_solve_phased"):
        getattr(lib, name).argtypes = [C.POINTER(abi.QmpcConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return lib

def run(lib, fn, cfg, probs, sched=None, warm=None):
    out = np.zeros(len(probs), dtype=abi.RESULT_DTYPE)
    rc = getattr(lib, fn)(C.byref(cfg), probs.ctypes.data, sched.ctypes.data if sched is not None else None,
                          warm.ctypes.data if warm is not None else None, len(probs), out.ctypes.data)
    assert rc == 0
    return out

def main():
    old_tree = sys.argv[1]
    extra = sys.argv[2:]
    tmp = tempfile.mkdtemp()
    # the old emul.cpp includes ../../quaternion_mpc_b200/csrc/...: rebuild that layout around the saved tree
    lay = os.path.join(tmp, "old"); os.makedirs(os.path.join(lay, "tests")); os.makedirs(os.path.join(lay, "quaternion_mpc_b200"))
    import shutil
    shutil.copytree(os.path.join(old_tree, "emul"), os.path.join(lay, "tests", "emul"))
    shutil.copytree(os.path.join(old_tree, "csrc"), os.path.join(lay, "quaternion_mpc_b200", "csrc"))
    shutil.copytree(os.path.join(old_tree, "include"), os.path.join(lay, "include"))
    a = build(os.path.join(lay, "tests", "emul", "emul.cpp"), os.path.join(tmp, "a.so"))
    b = build(os.path.join(ROOT, "tests", "emul", "emul.cpp"), os.path.join(tmp, "b.so"), extra)
    n = 0
    for model, N, gait, kw in [(0, 10, "trot", {}), (0, 16, "mixed", {}), (0, 32, "trot", {}), (0, 1, "trot", {}),
                               (1, 20, "stand", {"nfeet": 2, "max_angle": 0.2}), (2, 10, None, {}), (2, 20, None, {})]:
        cfg = default_config(model, N)
        p = random_convex_batch(48, seed=7 + N) if model == 2 else random_batch(48, seed=7 + N, gait=gait, **kw)
        sched = predict_schedule_numpy(random_gait_states(48, seed=3), N, cfg.dt)
        for fn in ("emul_solve_coop", "emul_solve_phased"):
            for sc in (None, sched):
                wa, wb = np.zeros(48, dtype=abi.WARM_DTYPE), np.zeros(48, dtype=abi.WARM_DTYPE)
                for tick in range(2):
                    ra, rb = run(a, fn, cfg, p, sc, wa), run(b, fn, cfg, p, sc, wb)
                    same = ra.tobytes() == rb.tobytes() and wa.tobytes() == wb.tobytes()
                    if not same:
                        d = np.abs(ra["grf_body"] - rb["grf_body"]).max()
                        print(f"DIFF model={model} N={N} {fn} sched={sc is not None} tick={tick}: max|dGRF|={d:.3e} "
                              f"iters_equal={(ra['iterations'] == rb['iterations']).all()}")
                        n += 1
    print("bit-identical on every workload" if n == 0 else f"{n} workloads differ")
    return 1 if n else 0

if __name__ == "__main__":
    sys.exit(main())
