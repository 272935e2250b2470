#!/usr/bin/env python
"""Independent fixed-point check of the oracle: for trot problems of the headline workload, run the
oracle to convergence (200 iterations allowed), then examine its answer with machinery that shares nothing with the
AL-iLQR restatement except the model: the plain NLP cost by single shooting + adjoint gradient (qmpc_ref_nlp_eval),
non-negative least squares for the cone multipliers (KKT residual), and SciPy's SLSQP started from the oracle's point.
Writes a markdown table (stdout).  usage: fixed_point_check.py [n_problems]"""
import ctypes as C
import os
import sys
import time

import numpy as np
from scipy.optimize import minimize, nnls

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as oracle                                   # noqa: E402
from quaternion_mpc_b200 import abi                                    # noqa: E402
from quaternion_mpc_b200.config import default_config                  # noqa: E402
from quaternion_mpc_b200.workloads import random_batch                 # noqa: E402


def nlp(cfg, prob):
    lib = oracle.lib()
    lib.qmpc_ref_nlp_eval.argtypes = [C.POINTER(abi.QmpcConfig)] + [C.c_void_p] * 8
    N, m = cfg.horizon, 12
    A, b = np.zeros((24, 12)), np.zeros((N, 24))
    U0, c = np.zeros(N * m), C.c_double()
    assert lib.qmpc_ref_nlp_eval(C.byref(cfg), prob.ctypes.data, None, U0.ctypes.data, C.byref(c), None, None, A.ctypes.data, b.ctypes.data) == 0

    def f(U):
        U = np.ascontiguousarray(U, dtype=float)
        g, c = np.zeros(N * m), C.c_double()
        lib.qmpc_ref_nlp_eval(C.byref(cfg), prob.ctypes.data, None, U.ctypes.data, C.byref(c), g.ctypes.data, None, None, None)
        return c.value, g
    Abig = np.zeros((N * 24, N * m))
    for k in range(N):
        Abig[24 * k:24 * k + 24, 12 * k:12 * k + 12] = A
    return f, Abig, b.reshape(-1)


def check_one(cfg, prob, U):
    f, A, b = nlp(cfg, prob)
    J, g = f(U)
    c = A @ U + b
    act = c > -1e-6
    lam, _ = nnls(A[act].T, -g)
    kkt = float(np.abs(g + A[act].T @ lam).max())
    r = minimize(f, U, jac=True, method="SLSQP", constraints=[{"type": "ineq", "fun": lambda u: -(A @ u + b), "jac": lambda u: -A}],
                 options={"ftol": 1e-15, "maxiter": 300})
    return {"viol": float(c.max()), "kkt": kkt, "dcost": float(J - r.fun), "rel_dcost": float((J - r.fun) / max(abs(J), 1e-12)),
            "du0": float(np.abs(r.x[:12] - U[:12]).max()), "slsqp_viol": float((A @ r.x + b).max()), "cost": float(J)}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cfg10, cfg = default_config(0, 10), default_config(0, 10)
    cfg.iterations_max = 200
    probs = random_batch(n, seed=0, gait="trot")
    w = np.zeros(n, dtype=abi.WARM_DTYPE)
    t0 = time.time()
    ref = oracle.solve_batch_warm(cfg, probs, w, nthreads=os.cpu_count() or 1)
    r10 = oracle.solve_batch(cfg10, probs, nthreads=os.cpu_count() or 1)
    conv = np.flatnonzero(ref["status"] == 0)
    rows = [check_one(cfg, probs[i:i + 1], w["u"][i][:10].reshape(-1).copy()) for i in conv]
    q = lambda k, p: float(np.quantile([r[k] for r in rows], p))
    d10 = np.abs(r10["grf_body"][conv] - ref["grf_body"][conv]).max(axis=1)
    print(f"# Independent fixed-point check of the oracle ({n} trot problems of the headline workload, seed 0)\n")
    print(f"oracle with 200 iterations allowed: {len(conv)} converged, {(ref['status'] == 2).sum()} line-search failed, "
          f"{(ref['status'] == 1).sum()} at the cap ({time.time() - t0:.0f} s incl. checks)\n")
    print("| quantity over the converged solves | median | p90 | max |\n|---|---|---|---|")
    for k, name in (("viol", "max cone violation at the oracle's point [N]"), ("kkt", "KKT residual of the plain NLP (adjoint gradient + NNLS multipliers)"),
                    ("dcost", "cost decrease SLSQP finds from the oracle's point"), ("rel_dcost", "... relative to the cost"),
                    ("du0", "how far SLSQP moves u0 from the oracle's point [N]")):
        print(f"| {name} | {q(k, .5):.2e} | {q(k, .9):.2e} | {q(k, 1):.2e} |")
    print(f"| 10-iteration u0 (what the reference returns) vs converged u0 [N] | {np.median(d10):.2e} | {np.quantile(d10, .9):.2e} | {d10.max():.2e} |")


if __name__ == "__main__":
    main()
