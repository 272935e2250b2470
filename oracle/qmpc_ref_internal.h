/* qmpc_ref_internal.h — model callbacks shared between qmpc_ref.c and kat_cases.c.
 * TEST INFRASTRUCTURE ONLY (see altro_ref.h). */
#ifndef QMPC_REF_INTERNAL_H_
#define QMPC_REF_INTERNAL_H_
typedef void (*ct_fn)(void* ctx, double* xdot, const double* x, const double* u);
typedef void (*ctj_fn)(void* ctx, double* jac, const double* x, const double* u);

typedef struct Model {
  int n, m, nf;
  double foot[12];   /* 3 x nf column-major lever arms */
  double Iinv[9];    /* row-major */
  double mass, g_vec[3], tau_g[3];
  double CR[18];     /* 6x3 row-major: C_mat * R0 (QuatMpc.cpp:203) or C_mat (ConvexMpc) */
  double fzmax_c[4]; /* fz_max * plan_contacts[i] */
  const double* fzmax_ck; /* optional per-knot bounds [k][4] (contact-schedule extension); NULL = fzmax_c */
  ct_fn f;
  ctj_fn df;
} Model;


void qref_inv3(const double* A, double* B);
void qref_quat_ct_dyn(void* ctx, double* xd, const double* x, const double* u);
void qref_quat_ct_jac(void* ctx, double* J, const double* x, const double* u);
void qref_convex_ct_dyn(void* ctx, double* xd, const double* x, const double* u);
void qref_convex_ct_jac(void* ctx, double* J, const double* x, const double* u);
void qref_mid_dyn(void* ctx, double* xn, const double* x, const double* u, float h);
void qref_mid_jac(void* ctx, double* J, const double* x, const double* u, float h);
void qref_cone_con(void* ctx, int k, double* c, const double* x, const double* u);
void qref_cone_jac(void* ctx, int k, double* J, const double* x, const double* u);
void qref_fill_cone(Model* M, double mu, const double* R0);
#endif
