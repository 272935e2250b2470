/*
 * altro_ref.h — CPU restatement (plain C, fp64) of the AL-iLQR solver the reference calls.
 *
 * TEST INFRASTRUCTURE ONLY: the product path (quaternion_mpc_b200/) never includes, links or
 * executes anything under oracle/.  Allowed users: tests/, __graft_entry__.smoke(), and
 * bench.py's cpu_baseline / --impl reference legs.
 *
 * The arithmetic restated here lives in a dependency that is ABSENT from /root/reference:
 *   ALTRO  github.com/zixinz990/altro @ b47202ffb9e09d5a2013d4661260988810e2eaef
 *   (legged_ctrl/CMakeLists.txt:34-40; fork of bjack205/altro adding the quaternion error state)
 * The API shape mirrors the calls the reference makes on altro::ALTROSolver
 * (legged_ctrl/src/mpc/QuatMpc.cpp:218-265, ConvexMpc.cpp:85,143-189) and the contract exercised
 * by legged_ctrl/src/test/test_altro/TestDoubleIntegrator.cpp / TestPendulum.cpp.
 *
 * Pinning: see tests/test_oracle_kats.py — double-integrator iteration counts 3 / 5
 * (TestDoubleIntegrator.cpp:255,374), saturated bounds (:367-372), pendulum integrator KATs and
 * swing-up end state (TestPendulum.cpp:31-42,110-114,202), and the two golden trajectories
 * quat_mpc_test.json / trot_quat_mpc_test.json.  Active-cone QuatMpc solves stopped at the
 * iteration cap are UNPINNED (no reference vector exists).
 */
#ifndef ALTRO_REF_H_
#define ALTRO_REF_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ALTRO_REF_EQUALITY 0   /* c(x,u) == 0 */
#define ALTRO_REF_INEQUALITY 1 /* c(x,u) <= 0  (TestDoubleIntegrator.cpp:296-303) */
#define ALTRO_REF_SOC 2        /* c(x,u) = (v, s) in the second-order cone |v| <= s, scalar last (:414-448); p <= 16 */

#define ALTRO_REF_SUCCESS 0
#define ALTRO_REF_MAX_ITERATIONS 1
#define ALTRO_REF_LINESEARCH_FAILED 2
#define ALTRO_REF_BACKWARD_FAILED 3
#define ALTRO_REF_NONFINITE 4

/* discrete dynamics x+ = f(x,u,h) and its Jacobian [A B], n x (n+m) COLUMN-major, exactly the
 * altro::ExplicitDynamicsFunction / ExplicitDynamicsJacobian signatures (float h included). */
typedef void (*altro_ref_dyn_fn)(void* ctx, double* xn, const double* x, const double* u, float h);
typedef void (*altro_ref_jac_fn)(void* ctx, double* jac, const double* x, const double* u, float h);
/* constraint value c[p] and Jacobian p x (ne+m) COLUMN-major in error coordinates
 * (altro::ConstraintFunction / ConstraintJacobian, QuatMpc.cpp:194-215) */
typedef void (*altro_ref_con_fn)(void* ctx, int k, double* c, const double* x, const double* u);
typedef void (*altro_ref_conjac_fn)(void* ctx, int k, double* jac, const double* x, const double* u);

typedef struct AltroRefOptions { /* altro::AltroOptions subset; defaults = altro_ref_default_options */
  int iterations_max;
  double tol_cost_intermediate, tol_primal_feasibility, tol_stationarity;
  double penalty_initial, penalty_scaling, penalty_max;
  int use_quaternion, quat_start_index;
  double ls_c1, ls_decrease; /* Armijo constant, back-tracking factor */
  int ls_iters_max;
  int use_backtracking_linesearch; /* 1: back-tracking (the MPC path); 0: strong-Wolfe cubic search (ALTRO's default) */
} AltroRefOptions;

typedef struct AltroRefProblem {
  int N, n, m;   /* horizon (knots 0..N), state dim, input dim */
  float h;       /* SetTimeStep */
  void* ctx;
  altro_ref_dyn_fn dyn;
  altro_ref_jac_fn jac;
  /* SetLQRCost / SetQuaternionCost per knot k: diag Q[k*n..], diag R[k*m..], refs, quaternion weight */
  const double *Q, *R, *xref, *uref, *w; /* sizes (N+1)*n, (N+1)*m, (N+1)*n, (N+1)*m, N+1 */
  /* one constraint block per knot: dimension p[k] (0 = none), type ctype[k] */
  const int *p, *ctype;
  altro_ref_con_fn con;
  altro_ref_conjac_fn conjac;
  const double* x0;
} AltroRefProblem;

typedef struct AltroRefStats {
  int iterations, status, ls_trials;
  double cost, max_violation, stationarity, penalty;
  double pivot_ratio; /* diagnostic: largest (max pivot / min pivot) of any Quu factored during the solve */
} AltroRefStats;

void altro_ref_default_options(AltroRefOptions* o);
/* projection onto the second-order cone {(v, s): |v| <= s} (scalar last) and, if J != NULL, its p x p Jacobian (row-major) */
void altro_ref_soc_project(const double* z, int p, double* out, double* J);

/* U (N*m): in = initial guess (SetInput), out = solution.  X ((N+1)*n): out = state trajectory.
 * Returns 0, or -1 on bad dimensions. */
int altro_ref_solve(const AltroRefProblem* prob, const AltroRefOptions* opts, double* X, double* U,
                    AltroRefStats* stats);

#ifdef __cplusplus
}
#endif
#endif
