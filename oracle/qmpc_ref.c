/*
 * qmpc_ref.c — fp64 CPU restatement of legged_ctrl's MPC problem assembly on top of altro_ref.c.
 * TEST INFRASTRUCTURE ONLY (see altro_ref.h).  Uses the product's public structs (include/qmpc.h)
 * so that the oracle and the CUDA library are fed byte-identical inputs.
 *
 * Follows, line by line:
 *   QuatMpc::grf_update          legged_ctrl/src/mpc/QuatMpc.cpp:109-276
 *   ConvexMpc::grf_update        legged_ctrl/src/mpc/ConvexMpc.cpp:81-198
 *   ct_srb_quat_dynamics/jacobian, ct_srb_trot_quat_*, ct_srb_dynamics/jacobian
 *                                legged_ctrl/src/utils/AltroUtils.cpp:224-513
 *   midpoint_dynamics/jacobian   legged_ctrl/src/utils/AltroUtils.cpp:9-22,78-110
 *   QuaternionUtils::G / L       legged_ctrl/src/utils/QuaternionUtils.cpp:30-52
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../include/qmpc.h"
#include "altro_ref.h"

#include "qmpc_ref_internal.h"

static void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
void qref_inv3(const double* A, double* B) {
  double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  B[0] = c00 / det; B[1] = (A[2] * A[7] - A[1] * A[8]) / det; B[2] = (A[1] * A[5] - A[2] * A[4]) / det;
  B[3] = c01 / det; B[4] = (A[0] * A[8] - A[2] * A[6]) / det; B[5] = (A[2] * A[3] - A[0] * A[5]) / det;
  B[6] = c02 / det; B[7] = (A[1] * A[6] - A[0] * A[7]) / det; B[8] = (A[0] * A[4] - A[1] * A[3]) / det;
}
static void mat3v(const double* A, const double* v, double* r) {
  for (int i = 0; i < 3; ++i) r[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
/* Eigen::Quaterniond::toRotationMatrix (BaseInterface.cpp:196), row-major, q = (w,x,y,z) */
static void quat_to_rot(const double* q, double* R) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

/* ---- quaternion SRB, nf feet (AltroUtils.cpp:363-392 / 441-468) */
void qref_quat_ct_dyn(void* ctx, double* xd, const double* x, const double* u) {
  const Model* M = (const Model*)ctx;
  const double *q = x + 3, *w = x + 10;
  double mom[3] = {M->tau_g[0], M->tau_g[1], M->tau_g[2]}, fs[3] = {0, 0, 0};
  for (int i = 0; i < M->nf; ++i) {
    double c[3];
    cross3(M->foot + 3 * i, u + 3 * i, c);
    for (int j = 0; j < 3; ++j) { mom[j] += c[j]; fs[j] += u[3 * i + j]; }
  }
  xd[0] = x[7]; xd[1] = x[8]; xd[2] = x[9];
  /* 0.5 * G(q) * omega */
  xd[3] = 0.5 * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2]);
  xd[4] = 0.5 * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2]);
  xd[5] = 0.5 * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2]);
  xd[6] = 0.5 * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2]);
  for (int j = 0; j < 3; ++j) xd[7 + j] = fs[j] / M->mass + M->g_vec[j];
  mat3v(M->Iinv, mom, xd + 10);
}
/* 13 x (13+m) column-major (AltroUtils.cpp:395-439 / 471-513) */
void qref_quat_ct_jac(void* ctx, double* J, const double* x, const double* u) {
  (void)u;
  const Model* M = (const Model*)ctx;
  const int n = 13, m = M->m;
  memset(J, 0, sizeof(double) * n * (n + m));
#define JJ(i, j) J[(j) * n + (i)]
  const double *q = x + 3, *w = x + 10;
  JJ(0, 7) = 1; JJ(1, 8) = 1; JJ(2, 9) = 1;
  JJ(3, 4) = -0.5 * w[0]; JJ(3, 5) = -0.5 * w[1]; JJ(3, 6) = -0.5 * w[2];
  JJ(4, 3) = 0.5 * w[0]; JJ(5, 3) = 0.5 * w[1]; JJ(6, 3) = 0.5 * w[2];
  /* -0.5 * skew(w) */
  JJ(4, 5) = 0.5 * w[2];  JJ(4, 6) = -0.5 * w[1];
  JJ(5, 4) = -0.5 * w[2]; JJ(5, 6) = 0.5 * w[0];
  JJ(6, 4) = 0.5 * w[1];  JJ(6, 5) = -0.5 * w[0];
  JJ(3, 10) = -0.5 * q[1]; JJ(3, 11) = -0.5 * q[2]; JJ(3, 12) = -0.5 * q[3];
  JJ(4, 10) = 0.5 * q[0];  JJ(4, 11) = -0.5 * q[3]; JJ(4, 12) = 0.5 * q[2];
  JJ(5, 10) = 0.5 * q[3];  JJ(5, 11) = 0.5 * q[0];  JJ(5, 12) = -0.5 * q[1];
  JJ(6, 10) = -0.5 * q[2]; JJ(6, 11) = 0.5 * q[1];  JJ(6, 12) = 0.5 * q[0];
  for (int i = 0; i < M->nf; ++i) {
    const double* r = M->foot + 3 * i;
    double S[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
    for (int a = 0; a < 3; ++a) {
      JJ(7 + a, 13 + 3 * i + a) = 1.0 / M->mass;
      for (int b = 0; b < 3; ++b) {
        double s = 0;
        for (int l = 0; l < 3; ++l) s += M->Iinv[3 * a + l] * S[3 * l + b];
        JJ(10 + a, 13 + 3 * i + b) = s;
      }
    }
  }
#undef JJ
}

/* ---- Euler-angle SRB of ConvexMpc (AltroUtils.cpp:224-359), intended 12-state math */
typedef struct ConvexCtx {
  Model M;
  double It[9]; /* hard-coded trunk inertia, AltroUtils.cpp:268-270 */
} ConvexCtx;
static void convex_Bc(const Model* M, double yaw, double* IwinvS /* 4 blocks 3x3 row-major */) {
  double cy = cos(yaw), sy = sin(yaw);
  double Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
  const double It[3] = {0.0168128557, 0.063009565, 0.0716547275};
  double Iw[9], Iwinv[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double s = 0;
      for (int l = 0; l < 3; ++l) s += Rz[3 * a + l] * It[l] * Rz[3 * b + l];
      Iw[3 * a + b] = s;
    }
  qref_inv3(Iw, Iwinv);
  for (int i = 0; i < 4; ++i) {
    const double* r = M->foot + 3 * i;
    double S[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double s = 0;
        for (int l = 0; l < 3; ++l) s += Iwinv[3 * a + l] * S[3 * l + b];
        IwinvS[9 * i + 3 * a + b] = s;
      }
  }
}
void qref_convex_ct_dyn(void* ctx, double* xd, const double* x, const double* u) {
  const Model* M = (const Model*)ctx;
  double cy = cos(x[2]), sy = sin(x[2]);
  double BS[36];
  convex_Bc(M, x[2], BS);
  /* rpy rate = [cy sy 0; -sy cy 0; 0 0 1] * omega_world (yaw-only map, AltroUtils.cpp:257-259) */
  xd[0] = cy * x[6] + sy * x[7];
  xd[1] = -sy * x[6] + cy * x[7];
  xd[2] = x[8];
  xd[3] = x[9]; xd[4] = x[10]; xd[5] = x[11];
  for (int a = 0; a < 3; ++a) {
    double s = 0, f = 0;
    for (int i = 0; i < 4; ++i) {
      for (int b = 0; b < 3; ++b) s += BS[9 * i + 3 * a + b] * u[3 * i + b];
      f += u[3 * i + a];
    }
    xd[6 + a] = s;
    xd[9 + a] = f / 12.84; /* hard-coded mass, AltroUtils.cpp:239 */
  }
  xd[11] += -9.81;
}
void qref_convex_ct_jac(void* ctx, double* J, const double* x, const double* u) {
  (void)u;
  const Model* M = (const Model*)ctx;
  const int n = 12;
  memset(J, 0, sizeof(double) * n * 24);
#define JJ(i, j) J[(j) * n + (i)]
  double cy = cos(x[2]), sy = sin(x[2]);
  double BS[36];
  convex_Bc(M, x[2], BS);
  JJ(0, 2) = x[7] * cy - x[6] * sy;  /* AltroUtils.cpp:355 */
  JJ(1, 2) = -x[6] * cy - x[7] * sy; /* AltroUtils.cpp:356 */
  JJ(0, 6) = cy; JJ(0, 7) = sy; JJ(1, 6) = -sy; JJ(1, 7) = cy; JJ(2, 8) = 1;
  JJ(3, 9) = 1; JJ(4, 10) = 1; JJ(5, 11) = 1;
  for (int i = 0; i < 4; ++i)
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) JJ(6 + a, 12 + 3 * i + b) = BS[9 * i + 3 * a + b];
      JJ(9 + a, 12 + 3 * i + a) = 1.0 / 12.84;
    }
#undef JJ
}

/* ---- explicit midpoint (AltroUtils.cpp:9-22, 78-110); float h, h/2 evaluated in float (exact) */
void qref_mid_dyn(void* ctx, double* xn, const double* x, const double* u, float h) {
  const Model* M = (const Model*)ctx;
  const int n = M->n;
  double xm[16], hh = (double)(h / 2), hd = (double)h;
  M->f(ctx, xm, x, u);
  for (int i = 0; i < n; ++i) xm[i] = xm[i] * hh + x[i];
  M->f(ctx, xn, xm, u);
  for (int i = 0; i < n; ++i) xn[i] = x[i] + hd * xn[i];
}
void qref_mid_jac(void* ctx, double* J, const double* x, const double* u, float h) {
  const Model* M = (const Model*)ctx;
  const int n = M->n, m = M->m;
  double xm[16], hh = (double)(h / 2), hd = (double)h;
  double Jc[16 * 28], Jm[16 * 28];
  M->f(ctx, xm, x, u);
  for (int i = 0; i < n; ++i) xm[i] = x[i] + hh * xm[i];
  M->df(ctx, Jc, x, u);
  M->df(ctx, Jm, xm, u);
  /* A_d = I + h Am (I + h/2 A) ; B_d = h (Am (h/2) B + Bm) ; column-major n x (n+m) */
  for (int j = 0; j < n + m; ++j)
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int l = 0; l < n; ++l) s += Jm[l * n + i] * Jc[j * n + l];
      double v = hd * (hh * s + Jm[j * n + i]);
      if (j < n && i == j) v += 1.0;
      J[j * n + i] = v;
    }
}

/* ---- friction-cone rows (QuatMpc.cpp:194-215; ConvexMpc.cpp:14-35,130-140) */
void qref_cone_con(void* ctx, int k, double* c, const double* x, const double* u) {
  (void)x;
  const Model* M = (const Model*)ctx;
  const double* fzb = M->fzmax_ck ? M->fzmax_ck + 4 * k : M->fzmax_c;
  for (int i = 0; i < M->nf; ++i) {
    for (int r = 0; r < 6; ++r)
      c[6 * i + r] = M->CR[3 * r] * u[3 * i] + M->CR[3 * r + 1] * u[3 * i + 1] + M->CR[3 * r + 2] * u[3 * i + 2];
    c[6 * i + 4] += -fzb[i];
  }
}
void qref_cone_jac(void* ctx, int k, double* J, const double* x, const double* u) {
  (void)k; (void)x; (void)u;
  const Model* M = (const Model*)ctx;
  const int p = 6 * M->nf, ne = 12;
  for (int i = 0; i < M->nf; ++i)
    for (int r = 0; r < 6; ++r)
      for (int b = 0; b < 3; ++b) J[(ne + 3 * i + b) * p + 6 * i + r] = M->CR[3 * r + b];
}

void qref_fill_cone(Model* M, double mu, const double* R0 /* NULL = identity */) {
  const double C[18] = {1, 0, -mu, -1, 0, -mu, 0, 1, -mu, 0, -1, -mu, 0, 0, 1, 0, 0, -1};
  for (int r = 0; r < 6; ++r)
    for (int b = 0; b < 3; ++b) {
      if (!R0) { M->CR[3 * r + b] = C[3 * r + b]; continue; }
      double s = 0;
      for (int l = 0; l < 3; ++l) s += C[3 * r + l] * R0[3 * l + b];
      M->CR[3 * r + b] = s;
    }
}

static void opts_from_cfg(const QmpcConfig* cfg, AltroRefOptions* o) {
  altro_ref_default_options(o);
  o->iterations_max = cfg->iterations_max;
  o->penalty_initial = cfg->penalty_initial;
  o->penalty_scaling = cfg->penalty_scaling;
  o->penalty_max = cfg->penalty_max;
  o->tol_cost_intermediate = cfg->tol_cost_intermediate;
  o->tol_primal_feasibility = cfg->tol_primal_feasibility;
  o->tol_stationarity = cfg->tol_stationarity;
}

/* ============================================================ QuatMpc::grf_update, one problem */
/* `sched` = NULL: the reference's behaviour (one contact mask over the horizon).  Otherwise
 * QMPC_MAX_HORIZON bytes, bit i of byte k = foot i in contact at knot k: the per-step contact
 * schedule the reference flags as TODO (ConvexMpc.cpp:82; LeggedContactFSM.cpp:272-286 is the
 * unused predictor).  Extension semantics: u_traj_ref[k] and the fz bound of knot k use mask k;
 * a knot with no contact has u_ref = 0; SetInput(u_traj_ref.at(0)) is kept verbatim. */
/* `warm` (nullable): the trajectory-shift warm start of include/qmpc.h (extension; ALTRO's
 * ShiftTrajectory pattern of TestBicycle.cpp:181-199, which legged_ctrl never calls). */
typedef struct QuatAssembly {
  Model M;
  double R0[9], qd[4], uref[12], x0[13];
  double Q[(QMPC_MAX_HORIZON + 1) * 13], R[(QMPC_MAX_HORIZON + 1) * 12], xref[(QMPC_MAX_HORIZON + 1) * 13],
      ur[(QMPC_MAX_HORIZON + 1) * 12], wq[QMPC_MAX_HORIZON + 1];
  int p[QMPC_MAX_HORIZON + 1], ct[QMPC_MAX_HORIZON + 1];
  double fzk[(QMPC_MAX_HORIZON + 1) * 4];
  AltroRefProblem P;
  int n, m, nf, N;
} QuatAssembly;

/* problem assembly of QuatMpc::grf_update (QuatMpc.cpp:118-253) into the ALTRO-shaped problem */
static int quat_assemble(const QmpcConfig* cfg, const QmpcProblem* in, const unsigned char* sched, QuatAssembly* A) {
  memset(A, 0, sizeof(*A));
#define M (A->M)
#define R0 (A->R0)
#define qd (A->qd)
#define uref (A->uref)
#define x0 (A->x0)
#define Q (A->Q)
#define R (A->R)
#define xref (A->xref)
#define ur (A->ur)
#define wq (A->wq)
#define p (A->p)
#define ct (A->ct)
#define fzk (A->fzk)
#define P (A->P)

  const int N = cfg->horizon;
  const int nf = cfg->model == QMPC_MODEL_QUAT_2FOOT ? 2 : 4;
  const int n = 13, m = 3 * nf;
  if (N < 1 || N > QMPC_MAX_HORIZON) return QMPC_ERR_ARG;
  A->n = n; A->m = m; A->nf = nf; A->N = N;
  M.n = n; M.m = m; M.nf = nf;
  M.f = qref_quat_ct_dyn; M.df = qref_quat_ct_jac;
  memcpy(M.foot, in->foot_pos_body, sizeof(double) * 3 * nf);
  qref_inv3(cfg->inertia, M.Iinv);
  M.mass = cfg->robot_mass;

  quat_to_rot(in->torso_quat, R0);
  double gw[3] = {0, 0, -cfg->gravity};
  if (cfg->model == QMPC_MODEL_QUAT_4FOOT) {
    /* g_body = R0^T g_world frozen at the measured attitude (AltroUtils.cpp:370-371) */
    for (int i = 0; i < 3; ++i) M.g_vec[i] = R0[i] * gw[0] + R0[3 + i] * gw[1] + R0[6 + i] * gw[2];
  } else {
    memcpy(M.g_vec, gw, sizeof(gw)); /* ct_srb_trot_quat_dynamics does not rotate (AltroUtils.cpp:446-447) */
  }
  double mg[3] = {cfg->com_mass * M.g_vec[0], cfg->com_mass * M.g_vec[1], cfg->com_mass * M.g_vec[2]};
  cross3(cfg->com_offset, mg, M.tau_g);
  /* QuatMpc: cone rows C_mat R0 (QuatMpc.cpp:194-215); two-contact problem: C_mat unrotated
     (TestAltroTrotQuatMpc.cpp:101-110) */
  qref_fill_cone(&M, cfg->mu, cfg->model == QMPC_MODEL_QUAT_4FOOT ? R0 : NULL);

  /* references (QuatMpc.cpp:118-176) */
  int num_contacts = 0;
  for (int i = 0; i < nf; ++i) num_contacts += in->plan_contacts[i] ? 1 : 0;
  for (int i = 0; i < nf; ++i) {
    uref[3 * i + 2] = (in->plan_contacts[i] ? 1.0 : 0.0) * cfg->robot_mass * cfg->gravity / num_contacts;
    M.fzmax_c[i] = cfg->fz_max * (in->plan_contacts[i] ? 1.0 : 0.0);
  }
  /* torso_quat_d += 0.5 G(q_d) w_d * 5 ms ; renormalise (QuatMpc.cpp:128-137) */
  {
    const double *q = in->torso_quat_d, *w = in->torso_ang_vel_d_body;
    double s = 0.5 * cfg->quat_d_dt;
    qd[0] = q[0] + s * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2]);
    qd[1] = q[1] + s * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2]);
    qd[2] = q[2] + s * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2]);
    qd[3] = q[3] + s * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2]);
    double nrm = sqrt(qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3]);
    for (int i = 0; i < 4; ++i) qd[i] /= nrm;
  }
  if (sched) {
    M.fzmax_ck = fzk;
    for (int k = 0; k < N; ++k) {
      int nc = 0;
      for (int i = 0; i < nf; ++i) nc += (sched[k] >> i) & 1;
      for (int i = 0; i < nf; ++i) {
        const double c = ((sched[k] >> i) & 1) ? 1.0 : 0.0;
        ur[k * m + 3 * i] = 0; ur[k * m + 3 * i + 1] = 0;
        ur[k * m + 3 * i + 2] = nc ? c * cfg->robot_mass * cfg->gravity / nc : 0.0;
        fzk[4 * k + i] = cfg->fz_max * c;
      }
    }
    memcpy(ur + N * m, ur + (N - 1) * m, sizeof(double) * m); /* terminal knot has no input cost */
    memcpy(uref, ur, sizeof(double) * m);                     /* SetInput(u_traj_ref.at(0)) :253 */
  }
  for (int k = 0; k <= N; ++k) {
    double* xr = xref + k * n;
    memset(xr, 0, sizeof(double) * n);
    /* i * h / 1000.0 with h in ms, evaluated in double (QuatMpc.cpp:156-157) */
    xr[0] = in->torso_pos_d_body[0] + in->torso_lin_vel_d_body[0] * k * cfg->dt;
    xr[1] = in->torso_pos_d_body[1] + in->torso_lin_vel_d_body[1] * k * cfg->dt;
    xr[2] = in->torso_pos_d_body[2];
    memcpy(xr + 3, qd, sizeof(qd));
    memcpy(xr + 7, in->torso_lin_vel_d_body, sizeof(double) * 3);
    memcpy(Q + k * n, cfg->q_weights, sizeof(double) * n);
    memcpy(R + k * m, cfg->r_weights, sizeof(double) * m);
    if (!sched) memcpy(ur + k * m, uref, sizeof(double) * m);
    wq[k] = cfg->w;
    /* SetConstraint(..., 0, horizon): knots 0..N-1 (QuatMpc.cpp:229) */
    p[k] = k < N ? 6 * nf : 0;
    ct[k] = ALTRO_REF_INEQUALITY;
  }
  /* x_init (QuatMpc.cpp:231-245): omega dropped by the reference's `;` at :242 */
  memcpy(x0 + 3, in->torso_quat, sizeof(double) * 4);
  for (int i = 0; i < 3; ++i)
    x0[7 + i] = R0[i] * in->torso_lin_vel_world[0] + R0[3 + i] * in->torso_lin_vel_world[1] +
                R0[6 + i] * in->torso_lin_vel_world[2];
  if (!cfg->drop_omega0) memcpy(x0 + 10, in->torso_ang_vel_body, sizeof(double) * 3);

#undef M
#undef R0
#undef qd
#undef uref
#undef x0
#undef Q
#undef R
#undef xref
#undef ur
#undef wq
#undef p
#undef ct
#undef fzk
#undef P
  A->P.N = N; A->P.n = n; A->P.m = m; A->P.h = (float)cfg->dt; A->P.ctx = &A->M;
  A->P.dyn = qref_mid_dyn; A->P.jac = qref_mid_jac;
  A->P.Q = A->Q; A->P.R = A->R; A->P.xref = A->xref; A->P.uref = A->ur; A->P.w = A->wq;
  A->P.p = A->p; A->P.ctype = A->ct; A->P.con = qref_cone_con; A->P.conjac = qref_cone_jac; A->P.x0 = A->x0;
  return QMPC_OK;
}

static __thread double tl_last_pivot_ratio = 0.0; /* diagnostic of the calling thread's last QuatMpc solve */

int qmpc_ref_solve_one_warm(const QmpcConfig* cfg, const QmpcProblem* in, const unsigned char* sched,
                            QmpcWarmStart* warm, QmpcResult* out) {
  QuatAssembly* A = (QuatAssembly*)malloc(sizeof(QuatAssembly));
  if (!A) return QMPC_ERR_ARG;
  int rc = quat_assemble(cfg, in, sched, A);
  if (rc) { free(A); return rc; }
  const int N = A->N, nf = A->nf, m = A->m;
  const double *R0 = A->R0, *qd = A->qd, *uref = A->uref;
  AltroRefOptions o;
  opts_from_cfg(cfg, &o);
  o.use_quaternion = 1;
  o.quat_start_index = 3;

  double X[(QMPC_MAX_HORIZON + 1) * 13], U[QMPC_MAX_HORIZON * 12];
  for (int k = 0; k < N; ++k) memcpy(U + k * m, uref, sizeof(double) * m); /* SetInput(u_ref[0]) :253 */
  if (warm && warm->valid)
    for (int k = 0; k < N; ++k) memcpy(U + k * m, warm->u[k + 1 < N ? k + 1 : N - 1], sizeof(double) * m);
  AltroRefStats st;
  if (altro_ref_solve(&A->P, &o, X, U, &st)) { free(A); return QMPC_ERR_ARG; }
  tl_last_pivot_ratio = st.pivot_ratio;

  memset(out, 0, sizeof(*out));
  for (int i = 0; i < nf; ++i) {
    mat3v(R0, U + 3 * i, out->grf_world + 3 * i);            /* QuatMpc.cpp:268 */
    memcpy(out->grf_body + 3 * i, U + 3 * i, sizeof(double) * 3); /* QuatMpc.cpp:269 */
  }
  memcpy(out->torso_quat_d, qd, sizeof(double) * 4);
  out->max_violation = st.max_violation;
  out->iterations = st.iterations;
  out->status = st.status;
  if (warm) {
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < 12; ++i) warm->u[k][i] = i < m ? U[k * m + i] : 0.0;
    warm->valid = st.status != QMPC_STATUS_NONFINITE;
  }
  free(A);
  return QMPC_OK;
}


/* ---- the restated NLP itself, for solver-independent checks (tests/test_oracle_fixed_point.py): plain cost
 * (no augmented-Lagrangian terms) of the input trajectory U by single shooting, its gradient by the adjoint
 * recursion with the full-state midpoint Jacobians, and the linear cone rows  cone_A u_k + cone_b_k <= 0.
 * cone_A: (6 nf) x m row-major (the same for every knot), cone_b: N x (6 nf). */
int qmpc_ref_nlp_eval(const QmpcConfig* cfg, const QmpcProblem* in, const unsigned char* sched, const double* U, double* cost,
                      double* grad, double* X, double* cone_A, double* cone_b) {
  QuatAssembly* A = (QuatAssembly*)malloc(sizeof(QuatAssembly));
  if (!A) return QMPC_ERR_ARG;
  int rc = quat_assemble(cfg, in, sched, A);
  if (rc) { free(A); return rc; }
  const int N = A->N, n = A->n, m = A->m, nf = A->nf;
  const AltroRefProblem* P = &A->P;
  double Xl[(QMPC_MAX_HORIZON + 1) * 13];
  memcpy(Xl, A->x0, sizeof(double) * n);
  for (int k = 0; k < N; ++k) P->dyn(P->ctx, Xl + (k + 1) * n, Xl + k * n, U + k * m, P->h);
  double J = 0, lam[13], lx[13];
  for (int k = N; k >= 0; --k) {
    const double *x = Xl + k * n, *xr = P->xref + k * n, *Qk = P->Q + k * n;
    for (int i = 0; i < n; ++i) { double d = x[i] - xr[i]; J += 0.5 * Qk[i] * d * d; lx[i] = Qk[i] * d; }
    if (P->w[k] != 0.0) {
      double sdot = xr[3] * x[3] + xr[4] * x[4] + xr[5] * x[5] + xr[6] * x[6];
      J += P->w[k] * (1.0 - fabs(sdot));
      double sg = sdot >= 0 ? 1.0 : -1.0;
      for (int i = 0; i < 4; ++i) lx[3 + i] += -P->w[k] * sg * xr[3 + i];
    }
    if (k == N) { memcpy(lam, lx, sizeof(double) * n); continue; }
    const double *u = U + k * m, *urk = P->uref + k * m, *Rk = P->R + k * m;
    double Jd[13 * 25], nl[13];
    memset(Jd, 0, sizeof(Jd));
    P->jac(P->ctx, Jd, x, u, P->h); /* column-major n x (n + m) */
    for (int j = 0; j < m; ++j) {
      double d = u[j] - urk[j], g = Rk[j] * d;
      J += 0.5 * Rk[j] * d * d;
      for (int i = 0; i < n; ++i) g += Jd[(n + j) * n + i] * lam[i];
      if (grad) grad[k * m + j] = g;
    }
    for (int j = 0; j < n; ++j) {
      double sacc = lx[j];
      for (int i = 0; i < n; ++i) sacc += Jd[j * n + i] * lam[i];
      nl[j] = sacc;
    }
    memcpy(lam, nl, sizeof(double) * n);
  }
  if (cost) *cost = J;
  if (X) memcpy(X, Xl, sizeof(double) * (N + 1) * n);
  if (cone_A) {
    memset(cone_A, 0, sizeof(double) * 6 * nf * m);
    for (int i = 0; i < nf; ++i)
      for (int r = 0; r < 6; ++r)
        for (int b = 0; b < 3; ++b) cone_A[(6 * i + r) * m + 3 * i + b] = A->M.CR[3 * r + b];
  }
  if (cone_b) {
    double zero_u[12] = {0}, c[24];
    for (int k = 0; k < N; ++k) {
      P->con(P->ctx, k, c, Xl + k * n, zero_u);
      memcpy(cone_b + k * 6 * nf, c, sizeof(double) * 6 * nf);
    }
  }
  free(A);
  return QMPC_OK;
}

int qmpc_ref_solve_one_sched(const QmpcConfig* cfg, const QmpcProblem* in, const unsigned char* sched,
                             QmpcResult* out) {
  return qmpc_ref_solve_one_warm(cfg, in, sched, NULL, out);
}

int qmpc_ref_solve_one(const QmpcConfig* cfg, const QmpcProblem* in, QmpcResult* out) {
  return qmpc_ref_solve_one_sched(cfg, in, NULL, out);
}

/* ============================================================ ConvexMpc::grf_update, one problem */
int qmpc_ref_solve_one_convex_sched(const QmpcConfig* cfg, const QmpcConvexProblem* in, const unsigned char* sched,
                                    QmpcResult* out) {
  const int N = cfg->horizon, n = 12, m = 12;
  if (N < 1 || N > QMPC_MAX_HORIZON) return QMPC_ERR_ARG;
  Model M;
  memset(&M, 0, sizeof(M));
  M.n = n; M.m = m; M.nf = 4;
  M.f = qref_convex_ct_dyn; M.df = qref_convex_ct_jac;
  memcpy(M.foot, in->foot_pos_abs_com, sizeof(double) * 12);
  qref_fill_cone(&M, cfg->mu, NULL);
  int num_contacts = 0;
  for (int i = 0; i < 4; ++i) num_contacts += in->plan_contacts[i] ? 1 : 0;
  double uref[12] = {0};
  for (int i = 0; i < 4; ++i) {
    uref[3 * i + 2] = cfg->robot_mass * cfg->gravity / num_contacts * (in->plan_contacts[i] ? 1.0 : 0.0);
    M.fzmax_c[i] = cfg->fz_max * (in->plan_contacts[i] ? 1.0 : 0.0);
  }
  double Q[(QMPC_MAX_HORIZON + 1) * 12], R[(QMPC_MAX_HORIZON + 1) * 12], xref[(QMPC_MAX_HORIZON + 1) * 12],
      ur[(QMPC_MAX_HORIZON + 1) * 12], wq[QMPC_MAX_HORIZON + 1];
  int p[QMPC_MAX_HORIZON + 1], ct[QMPC_MAX_HORIZON + 1];
  double fzk[(QMPC_MAX_HORIZON + 1) * 4] = {0};
  if (sched) { /* per-step contacts: the TODO at ConvexMpc.cpp:82 (same extension rules as QuatMpc) */
    M.fzmax_ck = fzk;
    for (int k = 0; k < N; ++k) {
      int nc = 0;
      for (int i = 0; i < 4; ++i) nc += (sched[k] >> i) & 1;
      for (int i = 0; i < 4; ++i) {
        const double c = ((sched[k] >> i) & 1) ? 1.0 : 0.0;
        ur[k * m + 3 * i] = 0; ur[k * m + 3 * i + 1] = 0;
        ur[k * m + 3 * i + 2] = nc ? cfg->robot_mass * cfg->gravity / nc * c : 0.0;
        fzk[4 * k + i] = cfg->fz_max * c;
      }
    }
    memcpy(ur + N * m, ur + (N - 1) * m, sizeof(double) * m);
    memcpy(uref, ur, sizeof(double) * m);
  }
  for (int k = 0; k <= N; ++k) {
    double* xr = xref + k * n; /* ConvexMpc.cpp:95-108 */
    memset(xr, 0, sizeof(double) * n);
    xr[2] = in->torso_euler[2] + in->yaw_rate_d * cfg->dt * k;
    xr[3] = in->torso_pos_d_world[0]; xr[4] = in->torso_pos_d_world[1]; xr[5] = in->torso_pos_d_world[2];
    xr[8] = in->yaw_rate_d;
    xr[9] = in->torso_lin_vel_d_world[0]; xr[10] = in->torso_lin_vel_d_world[1];
    memcpy(Q + k * n, cfg->q_weights, sizeof(double) * n);
    memcpy(R + k * m, cfg->r_weights, sizeof(double) * m);
    if (!sched) memcpy(ur + k * m, uref, sizeof(double) * m);
    wq[k] = 0.0;
    /* reference range is [0, N+1) (ConvexMpc.cpp:153-154); at knot N the rows depend on no
       optimisation variable (u_N does not exist, Jx = 0) so they are a no-op and are dropped */
    p[k] = k < N ? 24 : 0;
    ct[k] = ALTRO_REF_INEQUALITY;
  }
  double x0[12];
  memcpy(x0, in->torso_euler, 24); memcpy(x0 + 3, in->torso_pos_world, 24);
  memcpy(x0 + 6, in->torso_ang_vel_world, 24); memcpy(x0 + 9, in->torso_lin_vel_world, 24);

  AltroRefProblem P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.n = n; P.m = m; P.h = (float)cfg->dt; P.ctx = &M;
  P.dyn = qref_mid_dyn; P.jac = qref_mid_jac; /* ConvexMpc.cpp:123-124 uses the midpoint rule */
  P.Q = Q; P.R = R; P.xref = xref; P.uref = ur; P.w = wq;
  P.p = p; P.ctype = ct; P.con = qref_cone_con; P.conjac = qref_cone_jac; P.x0 = x0;
  AltroRefOptions o;
  opts_from_cfg(cfg, &o);
  double X[(QMPC_MAX_HORIZON + 1) * 12], U[QMPC_MAX_HORIZON * 12];
  for (int k = 0; k < N; ++k) memcpy(U + k * m, uref, sizeof(double) * m);
  AltroRefStats st;
  if (altro_ref_solve(&P, &o, X, U, &st)) return QMPC_ERR_ARG;
  memset(out, 0, sizeof(*out));
  for (int i = 0; i < 4; ++i) {
    /* optimized_input = R0^T u (ConvexMpc.cpp:190-192); u itself is the world-frame GRF */
    const double *R0 = in->torso_rot_mat, *f = U + 3 * i;
    for (int a = 0; a < 3; ++a) out->grf_body[3 * i + a] = R0[a] * f[0] + R0[3 + a] * f[1] + R0[6 + a] * f[2];
    memcpy(out->grf_world + 3 * i, f, sizeof(double) * 3);
  }
  out->torso_quat_d[0] = 1.0;
  out->max_violation = st.max_violation;
  out->iterations = st.iterations;
  out->status = st.status;
  return QMPC_OK;
}

int qmpc_ref_solve_one_convex(const QmpcConfig* cfg, const QmpcConvexProblem* in, QmpcResult* out) {
  return qmpc_ref_solve_one_convex_sched(cfg, in, NULL, out);
}

/* ============================================================ batch drivers (one worker per core) */
typedef struct Job {
  double* pivot_ratio; /* NULL or batch diagnostics */
  const QmpcConfig* cfg;
  const void* in;
  const unsigned char* sched; /* NULL or batch x QMPC_MAX_HORIZON mask bytes */
  QmpcWarmStart* warm;        /* NULL or batch warm-start buffers (quat models) */
  QmpcResult* out;
  int lo, hi, convex, rc;
} Job;
static void* worker(void* arg) {
  Job* j = (Job*)arg;
  for (int i = j->lo; i < j->hi; ++i) {
    const unsigned char* sc = j->sched ? j->sched + (size_t)i * QMPC_MAX_HORIZON : NULL;
    int rc = j->convex ? qmpc_ref_solve_one_convex_sched(j->cfg, (const QmpcConvexProblem*)j->in + i, sc, j->out + i)
                       : qmpc_ref_solve_one_warm(j->cfg, (const QmpcProblem*)j->in + i, sc, j->warm ? j->warm + i : NULL,
                                                 j->out + i);
    if (rc) j->rc = rc;
    if (j->pivot_ratio && !j->convex) j->pivot_ratio[i] = tl_last_pivot_ratio;
  }
  return NULL;
}
static int run_batch_diag(const QmpcConfig* cfg, const void* in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                          QmpcResult* out, int nthreads, int convex, double* pivot_ratio);
static int run_batch(const QmpcConfig* cfg, const void* in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                     QmpcResult* out,
                     int nthreads, int convex) {
  return run_batch_diag(cfg, in, sched, warm, batch, out, nthreads, convex, NULL);
}
int qmpc_ref_solve_batch_diag(const QmpcConfig* cfg, const QmpcProblem* in, const unsigned char* sched, QmpcWarmStart* warm,
                              int batch, QmpcResult* out, double* pivot_ratio, int nthreads) {
  return run_batch_diag(cfg, in, sched, warm, batch, out, nthreads, 0, pivot_ratio);
}
static int run_batch_diag(const QmpcConfig* cfg, const void* in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                          QmpcResult* out, int nthreads, int convex, double* pivot_ratio) {
  if (!cfg || !in || !out || batch < 0) return QMPC_ERR_ARG;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads > batch) nthreads = batch > 0 ? batch : 1;
  pthread_t th[256];
  Job jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].pivot_ratio = pivot_ratio;
    jobs[t].cfg = cfg; jobs[t].in = in; jobs[t].sched = sched; jobs[t].warm = warm; jobs[t].out = out; jobs[t].convex = convex; jobs[t].rc = 0;
    jobs[t].lo = (int)((long long)batch * t / nthreads);
    jobs[t].hi = (int)((long long)batch * (t + 1) / nthreads);
    if (nthreads == 1) worker(&jobs[t]);
    else pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  int rc = 0;
  for (int t = 0; t < nthreads; ++t) {
    if (nthreads > 1) pthread_join(th[t], NULL);
    if (jobs[t].rc) rc = jobs[t].rc;
  }
  return rc;
}
int qmpc_ref_solve_batch(const QmpcConfig* cfg, const QmpcProblem* in, int batch, QmpcResult* out, int nthreads) {
  return run_batch(cfg, in, NULL, NULL, batch, out, nthreads, 0);
}
int qmpc_ref_solve_batch_convex(const QmpcConfig* cfg, const QmpcConvexProblem* in, int batch, QmpcResult* out,
                                int nthreads) {
  return run_batch(cfg, in, NULL, NULL, batch, out, nthreads, 1);
}
int qmpc_ref_solve_batch_sched(const QmpcConfig* cfg, const QmpcProblem* in, const unsigned char* sched, int batch,
                               QmpcResult* out, int nthreads) {
  return run_batch(cfg, in, sched, NULL, batch, out, nthreads, 0);
}
int qmpc_ref_solve_batch_warm(const QmpcConfig* cfg, const QmpcProblem* in, const unsigned char* sched,
                              QmpcWarmStart* warm, int batch, QmpcResult* out, int nthreads) {
  return run_batch(cfg, in, sched, warm, batch, out, nthreads, 0);
}
int qmpc_ref_solve_batch_convex_sched(const QmpcConfig* cfg, const QmpcConvexProblem* in, const unsigned char* sched,
                                      int batch, QmpcResult* out, int nthreads) {
  return run_batch(cfg, in, sched, NULL, batch, out, nthreads, 1);
}
