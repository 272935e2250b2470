// qmpc_models.cuh — device-side single-rigid-body models and per-problem set-up.
//
// What the reference computes here (all fp64):
//   QuadrupedModel::ct_srb_quat_dynamics / ct_srb_quat_jacobian   legged_ctrl/src/utils/AltroUtils.cpp:363-439
//   QuadrupedModel::ct_srb_trot_quat_*  (2 feet)                  AltroUtils.cpp:441-513
//   QuadrupedModel::ct_srb_dynamics / ct_srb_jacobian (Euler)     AltroUtils.cpp:224-359
//   QuaternionUtils::L / G                                        legged_ctrl/src/utils/QuaternionUtils.cpp:30-52
//   problem assembly of QuatMpc::grf_update                       legged_ctrl/src/mpc/QuatMpc.cpp:109-253
//   problem assembly of ConvexMpc::grf_update                     legged_ctrl/src/mpc/ConvexMpc.cpp:81-175
// Written from the mathematical statement in SURVEY.md appendix A; one thread owns one problem.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define QMPC_HD __host__ __device__
#else
// host-only translation units (tests/emul: runs the very same solver bodies on the CPU so they can
// be debugged in a GPU-less container; never part of libqmpc_b200.so)
#define QMPC_HD
struct double2 { double x, y; };
#endif

#include "../../include/qmpc.h"

namespace qmpc {

// sin and cos of one angle.  On the device one sincos() call: a single argument reduction for both (the Euler model's
// roll-outs evaluate the pair twice per knot); same values as sin() / cos() of the CUDA math library (checked on the
// device against the previous build, profiles/r02_run2*_bitcheck.log).  -DQMPC_SINCOS_SEPARATE restores the two calls.
QMPC_HD inline void qmpc_sincos(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__) && !defined(QMPC_SINCOS_SEPARATE)
  sincos(a, s, c);
#else
  *s = sin(a); *c = cos(a);
#endif
}

QMPC_HD inline void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

QMPC_HD inline void inv3(const double* A, double* B) {
  double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  B[0] = c00 / det; B[1] = (A[2] * A[7] - A[1] * A[8]) / det; B[2] = (A[1] * A[5] - A[2] * A[4]) / det;
  B[3] = c01 / det; B[4] = (A[0] * A[8] - A[2] * A[6]) / det; B[5] = (A[2] * A[3] - A[0] * A[5]) / det;
  B[6] = c02 / det; B[7] = (A[1] * A[6] - A[0] * A[7]) / det; B[8] = (A[0] * A[4] - A[1] * A[3]) / det;
}

// Eigen::Quaterniond::toRotationMatrix (BaseInterface.cpp:196), row-major, q = (w,x,y,z)
QMPC_HD inline void quat_to_rot(const double* q, double* R) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

// G(q) = L(q) H, 4x3 row-major
QMPC_HD inline void quat_G(const double* q, double* G) {
  G[0] = -q[1]; G[1] = -q[2]; G[2] = -q[3];
  G[3] = q[0];  G[4] = -q[3]; G[5] = q[2];
  G[6] = q[3];  G[7] = q[0];  G[8] = -q[1];
  G[9] = -q[2]; G[10] = q[1]; G[11] = q[0];
}

// Friction-cone rows per foot: C_mat * Rc with C_mat from QuatMpc.cpp:47-52
QMPC_HD inline void fill_cone(double mu, const double* Rc /* nullptr = identity */, double* CR) {
  const double C[18] = {1, 0, -mu, -1, 0, -mu, 0, 1, -mu, 0, -1, -mu, 0, 0, 1, 0, 0, -1};
  for (int r = 0; r < 6; ++r)
    for (int b = 0; b < 3; ++b) {
      if (!Rc) { CR[3 * r + b] = C[3 * r + b]; continue; }
      double s = 0;
      for (int l = 0; l < 3; ++l) s += C[3 * r + l] * Rc[3 * l + b];
      CR[3 * r + b] = s;
    }
}


// Per-knot contact schedule shared by the models.  byte k: bits 0..3 = foot i in contact at knot
// k, bits 4..6 = number of feet in contact.  The reference holds ONE mask over the whole horizon
// (QuatMpc.cpp:119-125, 202; ConvexMpc.cpp:82 "TODO"); a null `sched` reproduces exactly that.
// With a schedule (LeggedContactFSM::predict_contact_state) knot k uses its own mask
// for u_ref and the fz bounds; a knot without any contact gets u_ref = 0 instead of the reference's
// 0/0 (which a constant all-swing mask still reproduces: QuatMpc.cpp:122).
struct ContactPlan {
  unsigned char cm[QMPC_MAX_HORIZON];
  double wz[5];   // u_ref z of a stance foot when n feet are in contact; wz[0] = NaN (plain) or 0 (schedule)
  double fzmax;
  QMPC_HD void fill(int nf, int N, const int32_t* plan_contacts, const unsigned char* sched, double weight,
                    double fz_max, bool weight_first) {
    int m0 = 0;
    for (int i = 0; i < nf; ++i) m0 |= plan_contacts[i] ? (1 << i) : 0;
    for (int k = 0; k < QMPC_MAX_HORIZON; ++k) {
      int mk = (sched && k < N) ? (sched[k] & ((1 << nf) - 1)) : m0;
      int nc = (mk & 1) + ((mk >> 1) & 1) + ((mk >> 2) & 1) + ((mk >> 3) & 1);
      cm[k] = (unsigned char)(mk | (nc << 4));
    }
    // QuatMpc.cpp:122 evaluates c * m * 9.81 / nc, ConvexMpc.cpp:93 m * 9.81 / nc * c (c = 1 here)
    for (int n = 1; n <= 4; ++n) wz[n] = weight_first ? weight / n * 1.0 : 1.0 * weight / n;
    wz[0] = sched ? 0.0 : NAN;   // the reference's 0 * m g / 0 (QuatMpc.cpp:122) and m g / 0 * 0 (ConvexMpc.cpp:93)
    fzmax = fz_max;
  }
  QMPC_HD bool in_contact(int k, int f) const { return (cm[k] >> f) & 1; }
  QMPC_HD double urefz(int k, int f) const {
    const int c = cm[k], nc = c >> 4;
    return (((c >> f) & 1) || nc == 0) ? wz[nc] : 0.0;
  }
  QMPC_HD double fzc(int k, int f) const { return ((cm[k] >> f) & 1) ? fzmax : 0.0; }
};

// Warm start (include/qmpc.h QmpcWarmStart): initial input of knot k = previous solution shifted by one knot
QMPC_HD inline const double* warm_row(const QmpcWarmStart* w, int k, int N) {
  return w->u[k + 1 < N ? k + 1 : N - 1];
}

// ------------------------------------------------------------------------------------------------
// Quaternion SRB with NF feet (QuatMpc: NF = 4; 2-contact model: NF = 2)
template <int NF>
struct QuatModel {
  static constexpr int NX = 13, NE = 12, NU = 3 * NF, NC = 6 * NF, QI = 3;
  static constexpr bool kQuat = true;
  using Problem = QmpcProblem;
  // structure seen by the cooperative kernel (qmpc_coop.cuh): error-state blocks in memory order
  // [position, attitude, linear velocity, angular velocity] (kSwap = 0), three knot-dependent 3x3 blocks
  // (Aff, Afw, Cf), the (angular velocity, moment) block of M is h I
  static constexpr int kFeet = NF, kSwap = 0, NLIN = 27;
  static constexpr bool kDw = false;
  // index of the first weight of error-state block b in q_weights (the attitude block has its own Hessian)
  QMPC_HD static constexpr int qoff(int b) { return b == 0 ? 0 : (b == 2 ? 7 : 10); }

  double foot[3 * NF];
  double IS[9 * NF];  // Iinv * skew(r_i), 3x3 row-major per foot  (AltroUtils.cpp:433)
  double Iinv[9];
  double inv_mass, g[3], tau_g[3];
  double CR[18];
  ContactPlan cp;
  double qd[4], pd[3], vd[3];
  double R0[9];
  double dtk;  // reference time step for x_ref (double, QuatMpc.cpp:156)

  QMPC_HD double urefz(int k, int f) const { return cp.urefz(k, f); }
  QMPC_HD double uref_at(int k, int i) const { return (i % 3 == 2) ? cp.urefz(k, i / 3) : 0.0; }
  QMPC_HD double fzc(int k, int f) const { return cp.fzc(k, f); }

  QMPC_HD void setup(const QmpcConfig& cfg, const QmpcProblem& in, const unsigned char* sched, double* x0) {
    for (int i = 0; i < 3 * NF; ++i) foot[i] = in.foot_pos_body[i];
    inv3(cfg.inertia, Iinv);
    inv_mass = 1.0 / cfg.robot_mass;
    quat_to_rot(in.torso_quat, R0);
    const double gw[3] = {0, 0, -cfg.gravity};
    if (NF == 4) {
      for (int i = 0; i < 3; ++i) g[i] = R0[i] * gw[0] + R0[3 + i] * gw[1] + R0[6 + i] * gw[2];
    } else {
      for (int i = 0; i < 3; ++i) g[i] = gw[i];
    }
    double mg[3] = {cfg.com_mass * g[0], cfg.com_mass * g[1], cfg.com_mass * g[2]};
    cross3(cfg.com_offset, mg, tau_g);
    // QuatMpc rotates the cone rows into the body frame, C_mat R0 (QuatMpc.cpp:194-215); the reference's
    // two-contact problem uses C_mat as it is (TestAltroTrotQuatMpc.cpp:101-110)
    fill_cone(cfg.mu, NF == 4 ? R0 : nullptr, CR);
    for (int i = 0; i < NF; ++i) {
      const double* r = foot + 3 * i;
      const double S[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          double s = 0;
          for (int l = 0; l < 3; ++l) s += Iinv[3 * a + l] * S[3 * l + b];
          IS[9 * i + 3 * a + b] = s;
        }
    }
    cp.fill(NF, cfg.horizon, in.plan_contacts, sched, cfg.robot_mass * cfg.gravity, cfg.fz_max, false);
    {
      const double *q = in.torso_quat_d, *w = in.torso_ang_vel_d_body;
      double s = 0.5 * cfg.quat_d_dt;
      qd[0] = q[0] + s * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2]);
      qd[1] = q[1] + s * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2]);
      qd[2] = q[2] + s * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2]);
      qd[3] = q[3] + s * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2]);
      double nrm = sqrt(qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3]);
      for (int i = 0; i < 4; ++i) qd[i] /= nrm;
    }
    for (int i = 0; i < 3; ++i) { pd[i] = in.torso_pos_d_body[i]; vd[i] = in.torso_lin_vel_d_body[i]; }
    dtk = cfg.dt;
    for (int i = 0; i < NX; ++i) x0[i] = 0;
    for (int i = 0; i < 4; ++i) x0[3 + i] = in.torso_quat[i];
    for (int i = 0; i < 3; ++i)
      x0[7 + i] = R0[i] * in.torso_lin_vel_world[0] + R0[3 + i] * in.torso_lin_vel_world[1] +
                  R0[6 + i] * in.torso_lin_vel_world[2];
    if (!cfg.drop_omega0)
      for (int i = 0; i < 3; ++i) x0[10 + i] = in.torso_ang_vel_body[i];
  }

  QMPC_HD void xref(int k, double* xr) const {
    xr[0] = pd[0] + vd[0] * k * dtk;
    xr[1] = pd[1] + vd[1] * k * dtk;
    xr[2] = pd[2];
    for (int i = 0; i < 4; ++i) xr[3 + i] = qd[i];
    for (int i = 0; i < 3; ++i) { xr[7 + i] = vd[i]; xr[10 + i] = 0; }
  }

  QMPC_HD void ct_dyn(const double* x, const double* u, double* xd) const {
    const double *q = x + 3, *w = x + 10;
    double mom[3] = {0, 0, 0}, fs[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      double c[3];
      cross3(foot + 3 * i, u + 3 * i, c);
      for (int j = 0; j < 3; ++j) { mom[j] += c[j]; fs[j] += u[3 * i + j]; }
    }
    for (int j = 0; j < 3; ++j) mom[j] += tau_g[j];
    xd[0] = x[7]; xd[1] = x[8]; xd[2] = x[9];
    xd[3] = 0.5 * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2]);
    xd[4] = 0.5 * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2]);
    xd[5] = 0.5 * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2]);
    xd[6] = 0.5 * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2]);
    for (int j = 0; j < 3; ++j) xd[7 + j] = fs[j] * inv_mass + g[j];
    for (int a = 0; a < 3; ++a) xd[10 + a] = Iinv[3 * a] * mom[0] + Iinv[3 * a + 1] * mom[1] + Iinv[3 * a + 2] * mom[2];
  }

  // One explicit-midpoint step driven by the net wrench of the feet (fs = sum f_i, mom = sum r_i x f_i): the
  // roll-outs of the cooperative kernel never form u as an array.  Same expressions, in the same order, as
  // ct_dyn / mid_dyn (AltroUtils.cpp:9-22, 383-391).
  QMPC_HD void wrench_step(double* x, double fs0, double fs1, double fs2, double mom0, double mom1, double mom2,
                           double hd, double hh) const {
    mom0 += tau_g[0]; mom1 += tau_g[1]; mom2 += tau_g[2];
    const double al0 = fs0 * inv_mass + g[0], al1 = fs1 * inv_mass + g[1], al2 = fs2 * inv_mass + g[2];
    const double aw0 = Iinv[0] * mom0 + Iinv[1] * mom1 + Iinv[2] * mom2;
    const double aw1 = Iinv[3] * mom0 + Iinv[4] * mom1 + Iinv[5] * mom2;
    const double aw2 = Iinv[6] * mom0 + Iinv[7] * mom1 + Iinv[8] * mom2;
    double xm[NX];
    {
      const double *q = x + 3, *w = x + 10;
      xm[0] = x[7] * hh + x[0]; xm[1] = x[8] * hh + x[1]; xm[2] = x[9] * hh + x[2];
      xm[3] = (0.5 * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2])) * hh + q[0];
      xm[4] = (0.5 * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2])) * hh + q[1];
      xm[5] = (0.5 * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2])) * hh + q[2];
      xm[6] = (0.5 * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2])) * hh + q[3];
      xm[7] = al0 * hh + x[7]; xm[8] = al1 * hh + x[8]; xm[9] = al2 * hh + x[9];
      xm[10] = aw0 * hh + x[10]; xm[11] = aw1 * hh + x[11]; xm[12] = aw2 * hh + x[12];
    }
    {
      const double *q = xm + 3, *w = xm + 10;
      const double qd0 = 0.5 * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2]);
      const double qd1 = 0.5 * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2]);
      const double qd2 = 0.5 * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2]);
      const double qd3 = 0.5 * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2]);
      x[0] = x[0] + hd * xm[7]; x[1] = x[1] + hd * xm[8]; x[2] = x[2] + hd * xm[9];
      x[3] = x[3] + hd * qd0; x[4] = x[4] + hd * qd1; x[5] = x[5] + hd * qd2; x[6] = x[6] + hd * qd3;
      x[7] = x[7] + hd * al0; x[8] = x[8] + hd * al1; x[9] = x[9] + hd * al2;
      x[10] = x[10] + hd * aw0; x[11] = x[11] + hd * aw1; x[12] = x[12] + hd * aw2;
    }
  }

  // dense column-major NX x (NX+NU)
  QMPC_HD void ct_jac(const double* x, const double* u, double* J) const {
    (void)u;
    for (int i = 0; i < NX * (NX + NU); ++i) J[i] = 0;
#define JJ(i, j) J[(j) * NX + (i)]
    const double *q = x + 3, *w = x + 10;
    JJ(0, 7) = 1; JJ(1, 8) = 1; JJ(2, 9) = 1;
    JJ(3, 4) = -0.5 * w[0]; JJ(3, 5) = -0.5 * w[1]; JJ(3, 6) = -0.5 * w[2];
    JJ(4, 3) = 0.5 * w[0]; JJ(5, 3) = 0.5 * w[1]; JJ(6, 3) = 0.5 * w[2];
    JJ(4, 5) = 0.5 * w[2];  JJ(4, 6) = -0.5 * w[1];
    JJ(5, 4) = -0.5 * w[2]; JJ(5, 6) = 0.5 * w[0];
    JJ(6, 4) = 0.5 * w[1];  JJ(6, 5) = -0.5 * w[0];
    JJ(3, 10) = -0.5 * q[1]; JJ(3, 11) = -0.5 * q[2]; JJ(3, 12) = -0.5 * q[3];
    JJ(4, 10) = 0.5 * q[0];  JJ(4, 11) = -0.5 * q[3]; JJ(4, 12) = 0.5 * q[2];
    JJ(5, 10) = 0.5 * q[3];  JJ(5, 11) = 0.5 * q[0];  JJ(5, 12) = -0.5 * q[1];
    JJ(6, 10) = -0.5 * q[2]; JJ(6, 11) = 0.5 * q[1];  JJ(6, 12) = 0.5 * q[0];
    for (int i = 0; i < NF; ++i)
      for (int a = 0; a < 3; ++a) {
        JJ(7 + a, 13 + 3 * i + a) = inv_mass;
        for (int b = 0; b < 3; ++b) JJ(10 + a, 13 + 3 * i + b) = IS[9 * i + 3 * a + b];
      }
#undef JJ
  }

  QMPC_HD void write_result(const double* u0, QmpcResult& out) const {
    for (int i = 0; i < 12; ++i) { out.grf_body[i] = 0; out.grf_world[i] = 0; }
    for (int i = 0; i < NF; ++i) {
      const double* f = u0 + 3 * i;
      for (int a = 0; a < 3; ++a) {
        out.grf_world[3 * i + a] = R0[3 * a] * f[0] + R0[3 * a + 1] * f[1] + R0[3 * a + 2] * f[2];
        out.grf_body[3 * i + a] = f[a];
      }
    }
    for (int i = 0; i < 4; ++i) out.torso_quat_d[i] = qd[i];
  }
};

// ------------------------------------------------------------------------------------------------
// Euler-angle SRB of ConvexMpc: x = [rpy, p_w, omega_w, v_w]
struct ConvexModel {
  static constexpr int NX = 12, NE = 12, NU = 12, NC = 24, QI = -1;
  static constexpr bool kQuat = false;
  using Problem = QmpcConvexProblem;
  // structure seen by the cooperative kernel: state blocks in memory order [attitude (rpy), position, angular
  // velocity, linear velocity] = the quaternion model's roles with neighbours swapped (kSwap = 1); FOUR
  // knot-dependent 3x3 blocks: Aff, Afw, Cf and Dw = h (Rz I Rz^T)^-1 at the midpoint yaw, the (angular
  // velocity, moment) block of M (the inertia is yaw-dependent here, so it cannot be folded into W)
  static constexpr int kFeet = 4, kSwap = 1, NLIN = 36;
  static constexpr bool kDw = true;
  QMPC_HD static constexpr int qoff(int b) { return 3 * b; }

  double foot[12];
  double IS[36];   // skew(r_i), 3x3 row-major per foot: the moment rows of W
  double inv_mass;
  double CR[18];
  ContactPlan cp;
  double xr0[12], yaw_rate, dtk;
  double R0[9];

  QMPC_HD double urefz(int k, int f) const { return cp.urefz(k, f); }
  QMPC_HD double uref_at(int k, int i) const { return (i % 3 == 2) ? cp.urefz(k, i / 3) : 0.0; }
  QMPC_HD double fzc(int k, int f) const { return cp.fzc(k, f); }

  QMPC_HD void setup(const QmpcConfig& cfg, const QmpcConvexProblem& in, const unsigned char* sched, double* x0) {
    for (int i = 0; i < 12; ++i) foot[i] = in.foot_pos_abs_com[i];
    for (int i = 0; i < 4; ++i) {
      const double* r = foot + 3 * i;
      const double S[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
      for (int a = 0; a < 9; ++a) IS[9 * i + a] = S[a];
    }
    inv_mass = 1.0 / 12.84;   // hard-coded mass of ct_srb_dynamics, AltroUtils.cpp:239
    for (int i = 0; i < 9; ++i) R0[i] = in.torso_rot_mat[i];
    fill_cone(cfg.mu, nullptr, CR);
    cp.fill(4, cfg.horizon, in.plan_contacts, sched, cfg.robot_mass * cfg.gravity, cfg.fz_max, true);
    for (int i = 0; i < 12; ++i) xr0[i] = 0;
    xr0[2] = in.torso_euler[2];
    for (int i = 0; i < 3; ++i) xr0[3 + i] = in.torso_pos_d_world[i];
    xr0[8] = in.yaw_rate_d;
    xr0[9] = in.torso_lin_vel_d_world[0];
    xr0[10] = in.torso_lin_vel_d_world[1];
    yaw_rate = in.yaw_rate_d;
    dtk = cfg.dt;
    for (int i = 0; i < 3; ++i) {
      x0[i] = in.torso_euler[i]; x0[3 + i] = in.torso_pos_world[i];
      x0[6 + i] = in.torso_ang_vel_world[i]; x0[9 + i] = in.torso_lin_vel_world[i];
    }
  }

  QMPC_HD void xref(int k, double* xr) const {
    for (int i = 0; i < 12; ++i) xr[i] = xr0[i];
    xr[2] = xr0[2] + yaw_rate * dtk * k;
  }

  // (Rz I_t Rz^T)^-1 skew(r_i), 4 blocks 3x3 row-major (AltroUtils.cpp:268-288)
  QMPC_HD void Bc(double yaw, double* BS) const {
    double sy, cy;
    sy = sin(yaw); cy = cos(yaw);
    const double Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
    const double It[3] = {0.0168128557, 0.063009565, 0.0716547275};
    double Iw[9], Iwinv[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double s = 0;
        for (int l = 0; l < 3; ++l) s += Rz[3 * a + l] * It[l] * Rz[3 * b + l];
        Iw[3 * a + b] = s;
      }
    inv3(Iw, Iwinv);
    for (int i = 0; i < 4; ++i) {
      const double* r = foot + 3 * i;
      const double S[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          double s = 0;
          for (int l = 0; l < 3; ++l) s += Iwinv[3 * a + l] * S[3 * l + b];
          BS[9 * i + 3 * a + b] = s;
        }
    }
  }

  // (Rz I_t Rz^T)^-1 = Rz I_t^-1 Rz^T in closed form (I_t diagonal, AltroUtils.cpp:268-270): symmetric, block
  // diagonal; returns i00, i01, i11, i22.  Used by the cooperative kernel's roll-outs and linearisation.
  QMPC_HD static void Iw_inv(double sy, double cy, double* iw) {
    const double ia = 1.0 / 0.0168128557, ib = 1.0 / 0.063009565, ic = 1.0 / 0.0716547275;
    iw[0] = ia * cy * cy + ib * sy * sy;
    iw[1] = (ia - ib) * sy * cy;
    iw[2] = ia * sy * sy + ib * cy * cy;
    iw[3] = ic;
  }
  // continuous dynamics from the net wrench (world frame): xd = f(x, fs, mom)
  QMPC_HD void wrench_dyn(const double* x, double fs0, double fs1, double fs2, double mom0, double mom1, double mom2,
                          double* xd) const {
    double sy, cy, iw[4];
    qmpc_sincos(x[2], &sy, &cy);
    Iw_inv(sy, cy, iw);
    xd[0] = cy * x[6] + sy * x[7];
    xd[1] = -sy * x[6] + cy * x[7];
    xd[2] = x[8];
    xd[3] = x[9]; xd[4] = x[10]; xd[5] = x[11];
    xd[6] = iw[0] * mom0 + iw[1] * mom1;
    xd[7] = iw[1] * mom0 + iw[2] * mom1;
    xd[8] = iw[3] * mom2;
    xd[9] = fs0 * inv_mass; xd[10] = fs1 * inv_mass; xd[11] = fs2 * inv_mass + -9.81;
  }
  QMPC_HD void wrench_step(double* x, double fs0, double fs1, double fs2, double mom0, double mom1, double mom2,
                           double hd, double hh) const {
    double xd[NX], xm[NX];
    wrench_dyn(x, fs0, fs1, fs2, mom0, mom1, mom2, xd);
#pragma unroll
    for (int i = 0; i < NX; ++i) xm[i] = xd[i] * hh + x[i];
    wrench_dyn(xm, fs0, fs1, fs2, mom0, mom1, mom2, xd);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = x[i] + hd * xd[i];
  }

  QMPC_HD void ct_dyn(const double* x, const double* u, double* xd) const {
    double sy, cy, BS[36];
    sy = sin(x[2]); cy = cos(x[2]);
    Bc(x[2], BS);
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
      xd[9 + a] = f / 12.84;
    }
    xd[11] += -9.81;
  }

  QMPC_HD void ct_jac(const double* x, const double* u, double* J) const {
    (void)u;
    for (int i = 0; i < NX * (NX + NU); ++i) J[i] = 0;
#define JJ(i, j) J[(j) * NX + (i)]
    double sy, cy, BS[36];
    sy = sin(x[2]); cy = cos(x[2]);
    Bc(x[2], BS);
    JJ(0, 2) = x[7] * cy - x[6] * sy;
    JJ(1, 2) = -x[6] * cy - x[7] * sy;
    JJ(0, 6) = cy; JJ(0, 7) = sy; JJ(1, 6) = -sy; JJ(1, 7) = cy; JJ(2, 8) = 1;
    JJ(3, 9) = 1; JJ(4, 10) = 1; JJ(5, 11) = 1;
    for (int i = 0; i < 4; ++i)
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) JJ(6 + a, 12 + 3 * i + b) = BS[9 * i + 3 * a + b];
        JJ(9 + a, 12 + 3 * i + a) = 1.0 / 12.84;
      }
#undef JJ
  }

  QMPC_HD void write_result(const double* u0, QmpcResult& out) const {
    for (int i = 0; i < 4; ++i) {
      const double* f = u0 + 3 * i;
      for (int a = 0; a < 3; ++a) {
        out.grf_body[3 * i + a] = R0[a] * f[0] + R0[3 + a] * f[1] + R0[6 + a] * f[2];  // R0^T u
        out.grf_world[3 * i + a] = f[a];
      }
    }
    out.torso_quat_d[0] = 1; out.torso_quat_d[1] = 0; out.torso_quat_d[2] = 0; out.torso_quat_d[3] = 0;
  }
};

}  // namespace qmpc
