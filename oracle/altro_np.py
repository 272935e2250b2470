"""NumPy restatement of the AL-iLQR ("ALTRO") solve used by legged_ctrl's QuatMpc / ConvexMpc.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported by the product path
(quaternion_mpc_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may use it, and only as the checker.

The solver arithmetic of the reference lives in a third-party library that is NOT in
/root/reference:  github.com/zixinz990/altro @ b47202ffb9e09d5a2013d4661260988810e2eaef
(fork of bjack205/altro, fetched by legged_ctrl/CMakeLists.txt:34-40).  This file restates the
published AL-iLQR algorithm (Howell/Jackson/Manchester, "ALTRO", IROS 2019; Jackson et al.,
"Planning with Attitude", RA-L 2021 for the quaternion error state) and anchors it on the
reference's own call sites, tests and golden vectors:

  * API semantics (index ranges, c<=0 inequality convention, error-coordinate constraint
    Jacobians, option names):  legged_ctrl/src/test/test_altro/TestDoubleIntegrator.cpp,
    TestPendulum.cpp, legged_ctrl/src/mpc/QuatMpc.cpp:194-229
  * pinned against: quat_mpc_test.json, trot_quat_mpc_test.json (golden trajectories),
    TestPendulum.cpp:31-42,110-113, TestDoubleIntegrator.cpp:51-66,255,367-374
    (see tests/test_oracle_kats.py)

Parity status: PINNED for the inactive-constraint regime (goldens) and for the toy KATs;
UNPINNED for QuatMpc solves with active cone constraints stopped at the iteration cap
(no reference vector exercises them).

Algorithm (one solve):
  rollout U -> X, merit phi
  for it in range(iterations_max):
      expansions at (X,U): error-state A_k,B_k, cost grad/Hessian, AL grad/GN-Hessian
      adjoint sweep -> stationarity = max_k |dL/du_k|_inf  (exact AL gradient)
      if it>0: convergence test / dual+penalty update (then AL terms are refreshed)
      Riccati backward pass (Cholesky of Quu, no regularisation)
      backtracking line search (alpha=1, x0.5, Armijo c1=1e-4, <=25 trials) on the AL merit
"""
import numpy as np

EQUALITY, INEQUALITY = 0, 1

DEFAULTS = dict(
    iterations_max=200, tol_cost=1e-4, tol_cost_intermediate=1e-4,
    tol_primal_feasibility=1e-4, tol_stationarity=1e-4,
    penalty_initial=1.0, penalty_scaling=10.0, penalty_max=1e8,
    ls_c1=1e-4, ls_decrease=0.5, ls_max=25, stat_mode='adjoint',
)


# ----------------------------------------------------------------------------- quaternion helpers
def skew(v):
    """Utils::skew  (legged_ctrl/src/utils/Utils.cpp:101-105)."""
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def quat_L(q):
    """QuaternionUtils::L (legged_ctrl/src/utils/QuaternionUtils.cpp:30-37); q=[w,x,y,z]."""
    L = np.zeros((4, 4))
    L[0, 0] = q[0]
    L[0, 1:] = -q[1:]
    L[1:, 0] = q[1:]
    L[1:, 1:] = q[0] * np.eye(3) + skew(q[1:])
    return L


def quat_G(q):
    """QuaternionUtils::G = L(q) H  (QuaternionUtils.cpp:48-52), 4x3."""
    return quat_L(q)[:, 1:]


def quat_to_rot(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


# ----------------------------------------------------------------------------- problem container
class Problem:
    """A generic ALTRO problem.  `qidx` = quat_start_index or None (plain vector state)."""

    def __init__(self, N, n, m, h, dyn, jac, x0, qidx=None):
        self.N, self.n, self.m = N, n, m
        self.h = float(np.float32(h))          # ALTRO passes the step as `float` (AltroUtils.cpp:10)
        self.dyn, self.jac = dyn, jac
        self.x0 = np.array(x0, float)
        self.qidx = qidx
        self.ne = n - 1 if qidx is not None else n
        self.Q = [np.zeros(n) for _ in range(N + 1)]
        self.R = [np.zeros(m) for _ in range(N + 1)]
        self.xref = [np.zeros(n) for _ in range(N + 1)]
        self.uref = [np.zeros(m) for _ in range(N + 1)]
        self.w = [0.0] * (N + 1)
        self.cons = [[] for _ in range(N + 1)]   # per knot: list of (fun, jac, dim, type)

    def set_lqr_cost(self, Q, R, xref, uref, k0=0, k1=None):
        for k in self._rng(k0, k1):
            self.Q[k], self.R[k] = np.array(Q, float), np.array(R, float)
            self.xref[k], self.uref[k] = np.array(xref, float), np.array(uref, float)
            self.w[k] = 0.0

    def set_quat_cost(self, Q, R, w, xref, uref, k0=0, k1=None):
        self.set_lqr_cost(Q, R, xref, uref, k0, k1)
        for k in self._rng(k0, k1):
            self.w[k] = float(w)

    def set_constraint(self, fun, jac, dim, ctype, k0=0, k1=None):
        for k in self._rng(k0, k1):
            self.cons[k].append((fun, jac, dim, ctype))

    def _rng(self, k0, k1):
        # ALTRO index convention: [k_start, k_stop); k_stop==0 -> just k_start; None -> all knots
        if k1 is None:
            return range(0, self.N + 1)
        if k1 == 0:
            return range(k0, k0 + 1)
        return range(k0, min(k1, self.N + 1))

    # ---- error-state machinery (Planning with Attitude; AltroUtils.cpp:153-168 shows the convention)
    def E(self, x):
        if self.qidx is None:
            return np.eye(self.n)
        i = self.qidx
        E = np.zeros((self.n, self.ne))
        E[:i, :i] = np.eye(i)
        E[i:i + 4, i:i + 3] = quat_G(x[i:i + 4])
        E[i + 4:, i + 3:] = np.eye(self.n - i - 4)
        return E

    def state_diff(self, x, xbar):
        """dx = x (-) xbar in error coordinates; attitude part = Cayley vector of conj(qbar)*q."""
        if self.qidx is None:
            return x - xbar
        i = self.qidx
        dq = quat_L(xbar[i:i + 4]).T @ x[i:i + 4]
        return np.concatenate([x[:i] - xbar[:i], dq[1:] / dq[0], x[i + 4:] - xbar[i + 4:]])


def solve(prob, U0, opts=None, trace=None):
    o = dict(DEFAULTS)
    if opts:
        o.update(opts)
    P = prob
    N, n, m, ne, h = P.N, P.n, P.m, P.ne, P.h
    hasu = [k < N for k in range(N + 1)]

    U = [np.array(U0[k], float).copy() for k in range(N)]
    mu = [[np.zeros(d) for (_, _, d, _) in P.cons[k]] for k in range(N + 1)]
    rho = [[o['penalty_initial'] for _ in P.cons[k]] for k in range(N + 1)]

    def rollout_open(U):
        X = [P.x0.copy()]
        for k in range(N):
            X.append(P.dyn(X[k], U[k], h))
        return X

    def cons_eval(X, U):
        out = []
        for k in range(N + 1):
            u = U[k] if k < N else np.zeros(m)
            out.append([f(X[k], u) for (f, _, _, _) in P.cons[k]])
        return out

    def stage_cost(k, x, u):
        dx = x - P.xref[k]
        J = 0.5 * dx @ (P.Q[k] * dx)
        if k < N:
            du = u - P.uref[k]
            J += 0.5 * du @ (P.R[k] * du)
        if P.w[k] != 0.0:
            i = P.qidx
            J += P.w[k] * (1.0 - abs(P.xref[k][i:i + 4] @ x[i:i + 4]))
        return J

    def merit(X, U):
        J = 0.0
        viol = 0.0
        C = cons_eval(X, U)
        for k in range(N + 1):
            J += stage_cost(k, X[k], U[k] if k < N else None)
            for j, (_, _, d, t) in enumerate(P.cons[k]):
                c, lam, r = C[k][j], mu[k][j], rho[k][j]
                if t == INEQUALITY:
                    lh = np.maximum(0.0, lam + r * c)
                    viol = max(viol, np.max(np.maximum(c, 0.0)))
                else:
                    lh = lam + r * c
                    viol = max(viol, np.max(np.abs(c)))
                J += (lh @ lh - lam @ lam) / (2 * r)
        return J, viol

    X = rollout_open(U)
    phi, viol = merit(X, U)
    status, iters = 'max_iterations', 0
    cost_decrease = np.inf

    for it in range(o['iterations_max']):
        # ---------------- dynamics + cost expansions (error state)
        A, B, lx, lu, lxx, luu = [], [], [], [], [], []
        for k in range(N + 1):
            Ek = P.E(X[k])
            dx = X[k] - P.xref[k]
            g = P.Q[k] * dx
            H = Ek.T @ (P.Q[k][:, None] * Ek)
            if P.qidx is not None:
                i = P.qidx
                q, qb = X[k][i:i + 4], P.xref[k][i:i + 4]
                if P.w[k] != 0.0:
                    s = 1.0 if qb @ q >= 0 else -1.0
                    g = g.copy()
                    g[i:i + 4] += -P.w[k] * s * qb
                # manifold Hessian correction: -I3 * (grad_q . q)
                H[i:i + 3, i:i + 3] -= np.eye(3) * (g[i:i + 4] @ q)
            lx.append(Ek.T @ g)
            lxx.append(H)
            if k < N:
                lu.append(P.R[k] * (U[k] - P.uref[k]))
                luu.append(np.diag(P.R[k]))
                Jd = P.jac(X[k], U[k], h)
                En = P.E(X[k + 1])
                A.append(En.T @ Jd[:, :n] @ Ek)
                B.append(En.T @ Jd[:, n:])

        def al_terms():
            """AL gradient / Gauss-Newton Hessian contributions at the current (X,U,mu,rho)."""
            gx = [np.zeros(ne) for _ in range(N + 1)]
            gu = [np.zeros(m) for _ in range(N + 1)]
            Hxx = [np.zeros((ne, ne)) for _ in range(N + 1)]
            Huu = [np.zeros((m, m)) for _ in range(N + 1)]
            Hux = [np.zeros((m, ne)) for _ in range(N + 1)]
            for k in range(N + 1):
                u = U[k] if k < N else np.zeros(m)
                for j, (f, jf, d, t) in enumerate(P.cons[k]):
                    c = f(X[k], u)
                    Jc = jf(X[k], u)           # d x (ne+m), error coordinates
                    lam, r = mu[k][j], rho[k][j]
                    if t == INEQUALITY:
                        est = lam + r * c
                        act = (est > 0).astype(float)
                        lh = np.maximum(0.0, est)
                    else:
                        act = np.ones(d)
                        lh = lam + r * c
                    Jx, Ju = Jc[:, :ne], Jc[:, ne:]
                    gx[k] += Jx.T @ lh
                    gu[k] += Ju.T @ lh
                    W = r * act
                    Hxx[k] += Jx.T @ (W[:, None] * Jx)
                    Huu[k] += Ju.T @ (W[:, None] * Ju)
                    Hux[k] += Ju.T @ (W[:, None] * Jx)
            return gx, gu, Hxx, Huu, Hux

        gx, gu, Hxx, Huu, Hux = al_terms()

        def stationarity():
            y = lx[N] + gx[N]
            s = 0.0
            for k in range(N - 1, -1, -1):
                s = max(s, np.max(np.abs(lu[k] + gu[k] + B[k].T @ y)))
                y = lx[k] + gx[k] + A[k].T @ y
            return s

        if o['stat_mode'] == 'adjoint' or it == 0:
            stat = stationarity()
        else:
            # ALTRO-style: residuals of the KKT stationarity conditions with the Riccati duals
            # y_k = P_k dx_k + p_k recorded during the accepted forward pass
            rx = np.max(np.abs(lx[N] + gx[N] - Y[N]))
            ru = 0.0
            for k in range(N):
                rx = max(rx, np.max(np.abs(lx[k] + gx[k] + A[k].T @ Y[k + 1] - Y[k])))
                ru = max(ru, np.max(np.abs(lu[k] + gu[k] + B[k].T @ Y[k + 1])))
            stat = max(rx, ru)
        if it > 0:
            if stat < o['tol_stationarity'] and viol < o['tol_primal_feasibility']:
                status = 'success'
                break
            if abs(cost_decrease) < o['tol_cost_intermediate'] or stat < o['tol_stationarity']:
                # dual + penalty update, then refresh the AL terms and the merit value
                C = cons_eval(X, U)
                for k in range(N + 1):
                    for j, (_, _, d, t) in enumerate(P.cons[k]):
                        est = mu[k][j] + rho[k][j] * C[k][j]
                        mu[k][j] = np.maximum(0.0, est) if t == INEQUALITY else est
                        rho[k][j] = min(rho[k][j] * o['penalty_scaling'], o['penalty_max'])
                gx, gu, Hxx, Huu, Hux = al_terms()
                phi, viol = merit(X, U)

        # ---------------- Riccati backward pass
        K, d = [None] * N, [None] * N
        Pm = lxx[N] + Hxx[N]
        p = lx[N] + gx[N]
        Ps, ps = [None] * (N + 1), [None] * (N + 1)
        Ps[N], ps[N] = Pm, p
        dphi0 = 0.0
        ok = True
        for k in range(N - 1, -1, -1):
            Qx = lx[k] + gx[k] + A[k].T @ p
            Qu = lu[k] + gu[k] + B[k].T @ p
            Qxx = lxx[k] + Hxx[k] + A[k].T @ Pm @ A[k]
            Quu = luu[k] + Huu[k] + B[k].T @ Pm @ B[k]
            Qux = Hux[k] + B[k].T @ Pm @ A[k]
            try:
                Lc = np.linalg.cholesky(Quu)
            except np.linalg.LinAlgError:
                ok = False
                break
            sol = np.linalg.solve(Lc.T, np.linalg.solve(Lc, np.column_stack([Qux, Qu])))
            K[k], d[k] = -sol[:, :ne], -sol[:, ne]
            Pm = Qxx + K[k].T @ Quu @ K[k] + K[k].T @ Qux + Qux.T @ K[k]
            Pm = 0.5 * (Pm + Pm.T)
            p = Qx + K[k].T @ Quu @ d[k] + K[k].T @ Qu + Qux.T @ d[k]
            dphi0 += Qu @ d[k]
            Ps[k], ps[k] = Pm, p
        if not ok:
            status = 'backward_pass_failed'
            break

        # ---------------- forward pass: backtracking line search on the AL merit
        alpha, accepted = 1.0, False
        for _ in range(o['ls_max']):
            Xn, Un = [P.x0.copy()], []
            for k in range(N):
                dxk = P.state_diff(Xn[k], X[k])
                Un.append(U[k] + alpha * d[k] + K[k] @ dxk)
                Xn.append(P.dyn(Xn[k], Un[k], h))
            phin, violn = merit(Xn, Un)
            if np.isfinite(phin) and phin <= phi + o['ls_c1'] * alpha * dphi0:
                accepted = True
                break
            alpha *= o['ls_decrease']
        iters = it + 1
        if trace is not None:
            trace.append(dict(it=it, phi0=phi, phi=phin, alpha=alpha, stat=stat, viol=violn,
                              dphi0=dphi0, rho=max([max(r) for r in rho if r] or [0]), accepted=accepted))
        if not accepted:
            status = 'linesearch_failed'
            break
        cost_decrease = phi - phin
        Y = [Ps[k] @ P.state_diff(Xn[k], X[k]) + ps[k] for k in range(N + 1)]
        X, U, phi, viol = Xn, Un, phin, violn

    return dict(X=X, U=U, status=status, iters=iters, viol=viol, mu=mu, rho=rho)


# ----------------------------------------------------------------------------- integrators
def midpoint_dynamics(f):
    """AltroUtils.cpp:9-22.  h is float32; h/2 is computed in float32 as in the reference."""
    def fd(x, u, h):
        h32 = np.float32(h)
        hh = float(h32 / np.float32(2))
        xm = x + hh * f(x, u)
        return x + float(h32) * f(xm, u)
    return fd


def midpoint_jacobian(f, df, n):
    """AltroUtils.cpp:78-110 (chain rule through the explicit midpoint step)."""
    def jd(x, u, h):
        h32 = np.float32(h)
        hh = float(h32 / np.float32(2))
        hf = float(h32)
        xm = x + hh * f(x, u)
        J = df(x, u)
        A, B = J[:, :n], J[:, n:]
        Jm = df(xm, u)
        Am, Bm = Jm[:, :n], Jm[:, n:]
        I = np.eye(n)
        Ad = I + hf * Am @ (I + hh * A)
        Bd = hf * (Am @ (hh * B) + Bm)
        return np.hstack([Ad, Bd])
    return jd


# ----------------------------------------------------------------------------- SRB models
R_COM = np.array([0.0223, 0.002, -0.0005])
TRUNK_MASS = 5.204


def srb_quat_model(foot_pos_body, inertia, mass, g_body):
    """ct_srb_quat_dynamics / ct_srb_quat_jacobian (AltroUtils.cpp:363-439) for nf feet.

    g_body = R0^T [0,0,-9.81] frozen at the measured attitude (AltroUtils.cpp:370-371);
    the 2-foot variant (AltroUtils.cpp:441-513) uses the world vector [0,0,-9.81].
    """
    r = np.asarray(foot_pos_body, float)      # 3 x nf
    nf = r.shape[1]
    Iinv = np.linalg.inv(inertia)
    tau_g = np.cross(R_COM, TRUNK_MASS * g_body)
    Bw = np.zeros((13, 3 * nf))
    for i in range(nf):
        Bw[7:10, 3 * i:3 * i + 3] = np.eye(3) / mass
        Bw[10:13, 3 * i:3 * i + 3] = Iinv @ skew(r[:, i])

    def f(x, u):
        xd = np.zeros(13)
        xd[0:3] = x[7:10]
        xd[3:7] = 0.5 * quat_G(x[3:7]) @ x[10:13]
        F = u.reshape(nf, 3)
        xd[7:10] = F.sum(0) / mass + g_body
        mom = tau_g.copy()
        for i in range(nf):
            mom += np.cross(r[:, i], F[i])
        xd[10:13] = Iinv @ mom
        return xd

    def df(x, u):
        J = np.zeros((13, 13 + 3 * nf))
        w = x[10:13]
        J[0:3, 7:10] = np.eye(3)
        J[3, 4:7] = -0.5 * w
        J[4:7, 3] = 0.5 * w
        J[4:7, 4:7] = -0.5 * skew(w)
        J[3:7, 10:13] = 0.5 * quat_G(x[3:7])
        J[:, 13:] = Bw
        return J
    return f, df
