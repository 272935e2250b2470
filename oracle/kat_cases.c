/*
 * kat_cases.c — C restatements of the reference's own solver test set-ups, run through altro_ref.c.
 * TEST INFRASTRUCTURE ONLY (see altro_ref.h).  Each function cites the reference test it rebuilds;
 * expected values are asserted in tests/test_oracle_kats.py.
 */
#include <math.h>
#include <string.h>

#include "altro_ref.h"
#include "qmpc_ref_internal.h"

/* ------------------------------------------------------------------------------------------------
 * Double integrator: legged_ctrl/src/test/test_altro/TestDoubleIntegrator.cpp:11-35 (dynamics),
 * :69-168 unconstrained, :170-256 goal equality (3 iterations), :258-375 control bounds
 * (u0 saturated at -1 to 1e-4, 5 iterations).
 */
static void di_dyn(void* ctx, double* xn, const double* x, const double* u, float h) {
  (void)ctx;
  double b = h * h / 2;
  for (int i = 0; i < 2; ++i) {
    xn[i] = x[i] + x[i + 2] * h + u[i] * b;
    xn[i + 2] = x[i + 2] + u[i] * h;
  }
}
static void di_jac(void* ctx, double* J, const double* x, const double* u, float h) {
  (void)ctx; (void)x; (void)u;
  double b = h * h / 2;
  memset(J, 0, sizeof(double) * 4 * 6);
  for (int i = 0; i < 2; ++i) {
    J[i * 4 + i] = 1.0;
    J[(i + 2) * 4 + i + 2] = 1.0;
    J[(i + 2) * 4 + i] = h;
    J[(4 + i) * 4 + i] = b;
    J[(4 + i) * 4 + i + 2] = h;
  }
}
void kat_di_dynamics(const double* x, const double* u, float h, double* xn, double* J) {
  di_dyn(0, xn, x, u, h);
  di_jac(0, J, x, u, h);
}
/* ctx: {N, soc}.  soc = 1: the control bound is the second-order cone (u, u_bnd), TestDoubleIntegrator.cpp:414-436 */
static void di_con(void* ctx, int k, double* c, const double* x, const double* u) {
  int N = ((int*)ctx)[0], soc = ((int*)ctx)[1];
  if (k == N) { /* goal constraint, xf = 0 */
    for (int i = 0; i < 4; ++i) c[i] = x[i];
  } else if (soc) { /* |u|_2 <= u_bnd = 1 as (u, u_bnd) in the cone */
    c[0] = u[0]; c[1] = u[1]; c[2] = 1.0;
  } else {      /* control bounds |u| <= 1 */
    for (int i = 0; i < 2; ++i) { c[i] = u[i] - 1.0; c[i + 2] = -1.0 - u[i]; }
  }
}
static void di_conjac(void* ctx, int k, double* J, const double* x, const double* u) {
  (void)x; (void)u;
  int N = ((int*)ctx)[0], soc = ((int*)ctx)[1];
  const int p = (k < N && soc) ? 3 : 4;
  if (k == N) {
    for (int i = 0; i < 4; ++i) J[i * p + i] = 1.0;
  } else if (soc) {
    for (int i = 0; i < 2; ++i) J[(4 + i) * p + i] = 1.0;   /* 3 x 6 column-major, d c_i / d u_i = 1 */
  } else {
    for (int i = 0; i < 2; ++i) { J[(4 + i) * p + i] = 1.0; J[(4 + i) * p + i + 2] = -1.0; }
  }
}
/* variant: 0 unconstrained, 1 goal equality, 2 goal + control bounds, 3 goal + second-order-cone control bound
 * (TestDoubleIntegrator.cpp:377-491) */
/* cubic != 0: ALTRO's default line search (the reference tests do not set use_backtracking_linesearch) */
int kat_double_integrator_ls(int variant, double penalty_initial, double penalty_scaling, int iterations_max, int cubic,
                             double* X, double* U, AltroRefStats* st);
int kat_double_integrator(int variant, double penalty_initial, double penalty_scaling, int iterations_max,
                          double* X, double* U, AltroRefStats* st) {
  return kat_double_integrator_ls(variant, penalty_initial, penalty_scaling, iterations_max, 0, X, U, st);
}
int kat_double_integrator_ls(int variant, double penalty_initial, double penalty_scaling, int iterations_max, int cubic,
                             double* X, double* U, AltroRefStats* st) {
  enum { N = 10, n = 4, m = 2 };
  float tf = 5.0f;
  const float h = tf / (double)N;
  double Q[(N + 1) * n], R[(N + 1) * m], xr[(N + 1) * n] = {0}, ur[(N + 1) * m] = {0}, w[N + 1] = {0};
  int p[N + 1] = {0}, ct[N + 1] = {0};
  for (int i = 0; i < (N + 1) * n; ++i) Q[i] = 1.0;
  for (int i = 0; i < (N + 1) * m; ++i) R[i] = 1e-2;
  double x0[4] = {variant >= 2 ? 2.0 : 1.0, 2.0, 0.0, 0.0};
  int NN[2] = {N, variant == 3};
  if (variant >= 1) { p[N] = 4; ct[N] = ALTRO_REF_EQUALITY; }
  if (variant == 2)
    for (int k = 0; k < N; ++k) { p[k] = 4; ct[k] = ALTRO_REF_INEQUALITY; }
  if (variant == 3)
    for (int k = 0; k < N; ++k) { p[k] = 3; ct[k] = ALTRO_REF_SOC; }
  AltroRefProblem P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.n = n; P.m = m; P.h = h; P.ctx = NN; P.dyn = di_dyn; P.jac = di_jac;
  P.Q = Q; P.R = R; P.xref = xr; P.uref = ur; P.w = w; P.p = p; P.ctype = ct;
  P.con = di_con; P.conjac = di_conjac; P.x0 = x0;
  AltroRefOptions o;
  altro_ref_default_options(&o);
  if (penalty_initial > 0) o.penalty_initial = penalty_initial;
  if (penalty_scaling > 0) o.penalty_scaling = penalty_scaling;
  if (iterations_max > 0) o.iterations_max = iterations_max;
  if (cubic) o.use_backtracking_linesearch = 0;
  memset(U, 0, sizeof(double) * N * m);
  return altro_ref_solve(&P, &o, X, U, st);
}

/* ------------------------------------------------------------------------------------------------
 * Pendulum: AltroTestUtils.cpp:43-82 (model), TestPendulum.cpp:15-43 (midpoint KATs),
 * :45-115 unconstrained swing-up, :117-203 goal-constrained.
 */
static void pend_f(void* ctx, double* xd, const double* x, const double* u) {
  (void)ctx;
  const double l = 0.5, g = 9.81, b = 0.1, m = 1.0 * l * l;
  xd[0] = x[1];
  xd[1] = u[0] / m - g * sin(x[0]) / l - b * x[1] / m;
}
static void pend_df(void* ctx, double* J, const double* x, const double* u) {
  (void)ctx; (void)u;
  const double l = 0.5, g = 9.81, b = 0.1, m = 1.0 * l * l;
  J[0] = 0.0; J[1] = -g * cos(x[0]) / l; J[2] = 1.0; J[3] = -b / m; J[4] = 0.0; J[5] = 1 / m;
}
static Model pend_model(void) {
  Model M;
  memset(&M, 0, sizeof(M));
  M.n = 2; M.m = 1; M.f = pend_f; M.df = pend_df;
  return M;
}
void kat_pendulum_midpoint(const double* x, const double* u, float h, double* xn, double* J) {
  Model M = pend_model();
  qref_mid_dyn(&M, xn, x, u, h);
  qref_mid_jac(&M, J, x, u, h);
}
static void pend_con(void* ctx, int k, double* c, const double* x, const double* u) {
  (void)ctx; (void)k; (void)u;
  c[0] = M_PI - x[0];
  c[1] = 0.0 - x[1];
}
static void pend_conjac(void* ctx, int k, double* J, const double* x, const double* u) {
  (void)ctx; (void)k; (void)x; (void)u;
  J[0] = -1.0; J[3] = -1.0; /* 2 x 3 column-major, -I on the state block */
}
/* variant 0: N=50, tf=3, unconstrained, iterations_max 20 ; variant 1: N=20, tf=2, goal equality */
int kat_pendulum_ls(int variant, int cubic, double* X, double* U, AltroRefStats* st);
int kat_pendulum(int variant, double* X, double* U, AltroRefStats* st) { return kat_pendulum_ls(variant, 0, X, U, st); }
int kat_pendulum_ls(int variant, int cubic, double* X, double* U, AltroRefStats* st) {
  enum { NMAX = 50, n = 2, m = 1 };
  const int N = variant == 0 ? 50 : 20;
  const float tf = variant == 0 ? 3.0f : 2.0f;
  const float h = tf / (double)N;
  double Q[(NMAX + 1) * n], R[(NMAX + 1) * m], xr[(NMAX + 1) * n], ur[(NMAX + 1) * m] = {0}, w[NMAX + 1] = {0};
  int p[NMAX + 1] = {0}, ct[NMAX + 1] = {0};
  for (int k = 0; k <= N; ++k) {
    Q[k * n] = Q[k * n + 1] = k < N ? 1e-2 : 1.0;
    R[k] = 1e-3;
    xr[k * n] = M_PI; xr[k * n + 1] = 0.0;
  }
  if (variant == 1) { p[N] = 2; ct[N] = ALTRO_REF_EQUALITY; }
  double x0[2] = {0, 0};
  Model M = pend_model();
  AltroRefProblem P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.n = n; P.m = m; P.h = h; P.ctx = &M; P.dyn = qref_mid_dyn; P.jac = qref_mid_jac;
  P.Q = Q; P.R = R; P.xref = xr; P.uref = ur; P.w = w; P.p = p; P.ctype = ct;
  P.con = pend_con; P.conjac = pend_conjac; P.x0 = x0;
  AltroRefOptions o;
  altro_ref_default_options(&o);
  o.iterations_max = variant == 0 ? 20 : 100;
  if (cubic) o.use_backtracking_linesearch = 0;
  for (int k = 0; k < N; ++k) U[k] = 0.1;
  return altro_ref_solve(&P, &o, X, U, st);
}

/* ------------------------------------------------------------------------------------------------
 * Quaternion-MPC goldens:
 *   which = 0: TestAltroQuatMpc.cpp (4 feet, stand)      -> quat_mpc_test.json
 *   which = 1: TestAltroTrotQuatMpc.cpp (2 feet, m = 6)  -> trot_quat_mpc_test.json
 * Both: N = 20, h = 0.01, iterations_max = 10, default penalties, inertia = (12.84/5.204) I_trunk,
 * gravity (0,0,-9.81) (identity attitude), cone rows without rotation.
 */
int kat_quat_golden(int which, double* X, double* U, AltroRefStats* st) {
  enum { N = 20, n = 13 };
  const int nf = which == 0 ? 4 : 2, m = 3 * nf;
  const double h = 0.01, mass = 12.84, torso_mass = 5.204;
  Model M;
  memset(&M, 0, sizeof(M));
  M.n = n; M.m = m; M.nf = nf; M.f = qref_quat_ct_dyn; M.df = qref_quat_ct_jac; M.mass = mass;
  if (which == 0) {
    const double fp[12] = {0.2104, 0.13, -0.325, 0.2104, -0.13, -0.325,
                           -0.1658, 0.13, -0.325, -0.1658, -0.13, -0.325}; /* TestAltroQuatMpc.cpp:41-44 */
    memcpy(M.foot, fp, sizeof(fp));
  } else {
    const double fp[6] = {0.17, 0.13, -0.3, -0.17, -0.13, -0.3};           /* TestAltroTrotQuatMpc.cpp:39-42 */
    memcpy(M.foot, fp, sizeof(fp));
  }
  double I[9] = {0.0168128557, 0, 0, 0, 0.063009565, 0, 0, 0, 0.0716547275};
  for (int i = 0; i < 9; ++i) I[i] *= mass / torso_mass;
  qref_inv3(I, M.Iinv);
  M.g_vec[2] = -9.81;
  const double rc[3] = {0.0223, 0.002, -0.0005};
  const double mg[3] = {0, 0, 5.204 * -9.81};
  M.tau_g[0] = rc[1] * mg[2] - rc[2] * mg[1];
  M.tau_g[1] = rc[2] * mg[0] - rc[0] * mg[2];
  M.tau_g[2] = rc[0] * mg[1] - rc[1] * mg[0];
  qref_fill_cone(&M, which == 0 ? 0.6 : 0.7, 0);
  for (int i = 0; i < nf; ++i) M.fzmax_c[i] = 200.0;

  double Q[(N + 1) * n], R[(N + 1) * 12], xr[(N + 1) * n], ur[(N + 1) * 12], w[N + 1];
  int p[N + 1], ct[N + 1];
  const double Qd0[13] = {1, 1, 1, 0, 0, 0, 0, 2, 2, 2, 1, 1, 1};
  const double Qd1[13] = {1, 1, 1, 0, 0, 0, 0, 10, 10, 10, 10, 10, 10};
  for (int k = 0; k <= N; ++k) {
    memcpy(Q + k * n, which == 0 ? Qd0 : Qd1, sizeof(Qd0));
    double* x = xr + k * n;
    memset(x, 0, sizeof(double) * n);
    x[3] = 1.0;
    if (which == 1) { x[0] = 0.5 * 0.5 * (h * k) * (h * k); x[7] = 0.5 * h * k; }
    for (int i = 0; i < m; ++i) {
      R[k * m + i] = 1e-6;
      ur[k * m + i] = (i % 3 == 2) ? mass * 9.81 / nf : 0.0;
    }
    w[k] = which == 0 ? 1.0 : 10.0;
    p[k] = k < N ? 6 * nf : 0; /* reference range [0, N+1): the terminal rows act on no variable */
    ct[k] = ALTRO_REF_INEQUALITY;
  }
  double x0[13] = {0};
  x0[3] = 1.0;
  AltroRefProblem P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.n = n; P.m = m; P.h = (float)h; P.ctx = &M; P.dyn = qref_mid_dyn; P.jac = qref_mid_jac;
  P.Q = Q; P.R = R; P.xref = xr; P.uref = ur; P.w = w; P.p = p; P.ctype = ct;
  P.con = qref_cone_con; P.conjac = qref_cone_jac; P.x0 = x0;
  AltroRefOptions o;
  altro_ref_default_options(&o);
  o.iterations_max = 10;
  o.use_quaternion = 1;
  o.quat_start_index = 3;
  for (int k = 0; k < N; ++k) memcpy(U + k * m, ur, sizeof(double) * m);
  return altro_ref_solve(&P, &o, X, U, st);
}

/* Roll a given input trajectory through the restated midpoint SRB dynamics (golden check B1/B2). */
void kat_quat_rollout(int which, const double* U, double* X) {
  enum { N = 20, n = 13 };
  double Xs[(N + 1) * n], Us[N * 12];
  AltroRefStats st;
  (void)Xs; (void)Us; (void)st;
  const int nf = which == 0 ? 4 : 2, m = 3 * nf;
  Model M;
  memset(&M, 0, sizeof(M));
  M.n = n; M.m = m; M.nf = nf; M.f = qref_quat_ct_dyn; M.df = qref_quat_ct_jac; M.mass = 12.84;
  if (which == 0) {
    const double fp[12] = {0.2104, 0.13, -0.325, 0.2104, -0.13, -0.325, -0.1658, 0.13, -0.325, -0.1658, -0.13, -0.325};
    memcpy(M.foot, fp, sizeof(fp));
  } else {
    const double fp[6] = {0.17, 0.13, -0.3, -0.17, -0.13, -0.3};
    memcpy(M.foot, fp, sizeof(fp));
  }
  double I[9] = {0.0168128557, 0, 0, 0, 0.063009565, 0, 0, 0, 0.0716547275};
  for (int i = 0; i < 9; ++i) I[i] *= 12.84 / 5.204;
  qref_inv3(I, M.Iinv);
  M.g_vec[2] = -9.81;
  M.tau_g[0] = 0.002 * (5.204 * -9.81);
  M.tau_g[1] = -0.0223 * (5.204 * -9.81);
  M.tau_g[2] = 0;
  memset(X, 0, sizeof(double) * n);
  X[3] = 1.0;
  for (int k = 0; k < N; ++k) qref_mid_dyn(&M, X + (k + 1) * n, X + k * n, U + k * m, 0.01f);
}
