"""Independent fixed-point check of the oracle's AL-iLQR restatement (no GPU).

The reference's solver (ALTRO) is absent, so besides the reference-held vectors (tests/test_oracle_kats.py) the
restatement is checked against machinery that shares nothing with it except the model: the plain NLP cost of an input
trajectory by single shooting with its adjoint gradient (oracle `qmpc_ref_nlp_eval`, itself checked here against
finite differences), non-negative least squares for the cone multipliers, and SciPy's SLSQP.

What can and cannot be asked (numbers: profiles/r02_fixed_point_check.md, tools/fixed_point_check.py): with
R = 1e-6 (gazebo_go1_quat_mpc.yaml:58-72) the trot problem has a nearly flat direction (the two stance feet squeezing
along the line that joins them), so at the reference's tolerances (1e-4) u0 is only determined to ~0.1 N - an
unrelated solver started AT the oracle's converged point lowers the cost by a relative 5e-7 and moves u0 by 0.04 N
(median).  The assertions are therefore: feasible to the solver's tolerance, KKT residual of the plain NLP at the
1e-3 level, no relevant improvement by SLSQP.  The 10-iteration iterate the reference actually returns is 0.2 N
(median) / 80 N (max) away from the converged point: a mid-flight AL iterate, unpinned."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_nlp_gradient_against_finite_differences(oracle):
    from fixed_point_check import nlp
    from quaternion_mpc_b200.config import default_config
    from quaternion_mpc_b200.workloads import random_batch
    cfg = default_config(0, 10)
    p = random_batch(3, seed=2, gait="mixed")
    rng = np.random.default_rng(0)
    for i in range(3):
        f, A, b = nlp(cfg, p[i:i + 1])
        U = rng.normal(0, 8, 120) + np.tile([0, 0, 30.0] * 4, 10)
        J, g = f(U)
        for j in rng.integers(0, 120, 12):
            e = np.zeros(120); e[j] = 1e-5
            fd = (f(U + e)[0] - f(U - e)[0]) / 2e-5
            assert abs(fd - g[j]) < 1e-6 * max(1.0, abs(g[j])), (i, j, fd, g[j])
        assert A.shape == (240, 120) and b.shape == (240,)
        # cone rows: zero force is feasible for every foot (fz_max >= 0 on stance feet, 0 <= fz <= 0 on swing feet)
        assert ((A @ np.zeros(120) + b) <= 0).all()


def test_converged_oracle_points_are_kkt_points_no_independent_solver_improves(oracle):
    from fixed_point_check import check_one
    from quaternion_mpc_b200 import abi
    from quaternion_mpc_b200.config import default_config
    from quaternion_mpc_b200.workloads import random_batch
    cfg = default_config(0, 10)
    cfg.iterations_max = 200
    n = 40
    probs = random_batch(n, seed=0, gait="trot")
    w = np.zeros(n, dtype=abi.WARM_DTYPE)
    ref = oracle.solve_batch_warm(cfg, probs, w, nthreads=os.cpu_count() or 1)
    conv = np.flatnonzero(ref["status"] == 0)[:12]
    assert len(conv) >= 8
    for i in conv:
        r = check_one(cfg, probs[i:i + 1], w["u"][i][:10].reshape(-1).copy())
        assert r["viol"] < 1e-4, (i, r)               # tol_primal_feasibility
        assert r["kkt"] < 2e-3, (i, r)                # stationarity of the PLAIN nlp (the solver tests its own measure at 1e-4)
        assert r["rel_dcost"] < 2e-6 and r["dcost"] > -1e-9, (i, r)   # SLSQP from this point: no relevant decrease
        assert r["du0"] < 0.5 and r["slsqp_viol"] < 1e-6, (i, r)       # ... and it stays in the flat neighbourhood
