// qmpc_periph.cuh — the data either side of the solve, batched.
//
//   N1  LeggedContactFSM::predict_contact_state over the shipped gait tables
//         legged_ctrl/src/utils/LeggedContactFSM.cpp:87-206 (tables), :272-286 (predictor)
//   N2  A1Kinematics::fk / jac for the four legs      legged_ctrl/src/utils/A1Kinematics.cpp:10-21
//         as called from BaseInterface.cpp:204-212 (foot_pos_body, jac_foot)
//       joint torque targets tau_i = -J_i^T f_i       BaseInterface.cpp:343-405
//
// All three are streaming, HBM-bound element-wise kernels (tens of bytes per robot): one thread per
// robot (N1) or per (robot, leg) (N2), consecutive threads touch consecutive addresses, grid sized
// to the batch.  The bodies are QMPC_HD so tests/emul can run them on the host.
#pragma once
#include <cstddef>
#include "qmpc_models.cuh"

namespace qmpc {

// no-FMA arithmetic where the result feeds a comparison that must match the reference bit for bit
QMPC_HD inline double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
QMPC_HD inline double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

QMPC_HD inline void sincos_rn(double a, double* s, double* c) {
#ifdef __CUDA_ARCH__
  sincos(a, s, c);   // one argument reduction for both (these kernels are fp64-trig bound, not HBM bound, otherwise)
#else
  *s = sin(a); *c = cos(a);
#endif
}

// cos / sin of the yaw angle of q = (w, x, y, z) without going through the angle: Utils::quat_to_euler
// returns yaw = atan2(t3, t4) (Utils.cpp:29-31) and BaseInterface.cpp:200 builds AngleAxis(yaw, UnitZ) from
// it; (t4, t3) / |(t4, t3)| is the same pair to 1 ulp and saves an fp64 atan2 + sincos per robot.
QMPC_HD inline void quat_yaw_cs(const double* q, double* cy, double* sy) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double t3 = +2.0 * (w * z + x * y);
  const double t4 = +1.0 - 2.0 * (y * y + z * z);
  const double r = sqrt(t3 * t3 + t4 * t4);
  if (r > 0.0) { *cy = t4 / r; *sy = t3 / r; }
  else { *cy = 1.0; *sy = 0.0; }   // atan2(0, 0) = 0
}

// One leg's gait pattern: up to 3 segments (switch time, STANCE?) over the unit gait cycle.
struct GaitLegPattern {
  int n;
  double sw[3];
  int stance[3];
};

// LeggedContactFSM.cpp:87-206.  leg: 0 FL, 1 FR, 2 RL, 3 RR.
QMPC_HD inline GaitLegPattern gait_pattern(int gait, int leg) {
  GaitLegPattern p{};
  const bool diagA = (leg == 0 || leg == 3);
  switch (gait) {
    case QMPC_GAIT_TROT:   // FL/RR stance first, FR/RL swing first, switch at half cycle
      p.n = 2; p.sw[0] = 0.5; p.sw[1] = 1.0;
      p.stance[0] = diagA ? 1 : 0; p.stance[1] = diagA ? 0 : 1;
      break;
    case QMPC_GAIT_TROT_WITH_STAND:
      if (diagA) { p.n = 2; p.sw[0] = 0.6; p.sw[1] = 1.0; p.stance[0] = 1; p.stance[1] = 0; }
      else { p.n = 3; p.sw[0] = 0.1; p.sw[1] = 0.5; p.sw[2] = 1.0; p.stance[0] = 1; p.stance[1] = 0; p.stance[2] = 1; }
      break;
    case QMPC_GAIT_CRAWL:
      if (leg == 0) { p.n = 2; p.sw[0] = 0.25; p.sw[1] = 1.0; p.stance[0] = 0; p.stance[1] = 1; }
      else if (leg == 1) { p.n = 3; p.sw[0] = 0.25; p.sw[1] = 0.5; p.sw[2] = 1.0; p.stance[0] = 1; p.stance[1] = 0; p.stance[2] = 1; }
      else if (leg == 2) { p.n = 3; p.sw[0] = 0.5; p.sw[1] = 0.75; p.sw[2] = 1.0; p.stance[0] = 1; p.stance[1] = 0; p.stance[2] = 1; }
      else { p.n = 2; p.sw[0] = 0.75; p.sw[1] = 1.0; p.stance[0] = 1; p.stance[1] = 0; }
      break;
    default:  // QMPC_GAIT_STAND
      p.n = 1; p.sw[0] = 1.0; p.stance[0] = 1;
      break;
  }
  // unused entries never match (ph <= -inf is false for every ph): predict_contact tests all three without a loop
  // over p.n, so the pattern stays in registers (the indexed loop put it in local memory: 3 LDL / 5 STL per leg)
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i >= p.n) { p.sw[i] = -INFINITY; p.stance[i] = 1; }
  return p;
}

// predict_contact_state(dt) of one leg (LeggedContactFSM.cpp:272-286); falls through to STANCE
QMPC_HD inline int predict_contact(const GaitLegPattern& p, double gait_phase, double gait_freq, double dt) {
  double ph = add_rn(gait_phase, mul_rn(gait_freq, dt));
  // the reference wraps with `while (ph > 1.0) ph -= 1.0` (LeggedContactFSM.cpp:275-277): for ph < 2^53 every
  // subtraction is exact, so the loop ends at ph - floor(ph), or at 1.0 when ph is a whole number >= 1 - computed
  // here in closed form (bit-identical), because one robot record with an Inf / NaN / 1e12 phase must not hang the
  // launch (and every later solve on the stream); non-finite and absurd phases fall through to STANCE
  if (ph > 1.0) {
    if (!(ph < 9.0e15)) return 1;
    const double r = ph - floor(ph);
    ph = (r == 0.0) ? 1.0 : r;
  }
  if (ph <= p.sw[0]) return p.stance[0];
  if (ph <= p.sw[1]) return p.stance[1];
  if (ph <= p.sw[2]) return p.stance[2];
  return 1;
}

QMPC_HD inline void predict_schedule_one(const QmpcGaitState& g, int N, double dt, QmpcContactSchedule& out) {
  GaitLegPattern pat[4];
  for (int leg = 0; leg < 4; ++leg) pat[leg] = gait_pattern(g.gait, leg);
  for (int k = 0; k < QMPC_MAX_HORIZON; ++k) {
    int m = 0;
    if (k < N) {
      const double t = mul_rn((double)k, dt);
      for (int leg = 0; leg < 4; ++leg) m |= predict_contact(pat[leg], g.gait_phase[leg], g.gait_freq, t) << leg;
    }
    out.mask[k] = (uint8_t)m;
  }
}

// Forward kinematics + Jacobian of one leg in closed form.  q = (hip, thigh, calf);
// rho_fix = (ox, oy, d, lt, lc), rho_opt = (cx, cy, cz) foot-contact offset.  With
// a = lc - cz, s12 = sin(q1 + q2), c12 = cos(q1 + q2), L = lt cos q1 + cx s12 + a c12 :
//   p = [ ox - lt sin q1 - a s12 + cx c12 ;  oy + (cy + d) cos q0 + L sin q0 ;  (cy + d) sin q0 - L cos q0 ]
// which is the polynomial A1Kinematics::autoFunc_fk_derive expands term by term
// (A1Kinematics.cpp:40-75); J = dp/dq follows by differentiation (autoFunc_d_fk_dq, :77-140).
// jac is written column-major (Eigen): jac[3 * col + row].
QMPC_HD inline void leg_fk_jac(const double* q, const double* rf, const double* ro, double* p, double* jac) {
  const double ox = rf[0], oy = rf[1], d = rf[2], lt = rf[3], lc = rf[4];
  const double cx = ro[0], cy = ro[1], cz = ro[2];
  double s0, c0, s1, c1, s12, c12;
  sincos_rn(q[0], &s0, &c0);
  sincos_rn(q[1], &s1, &c1);
  sincos_rn(q[1] + q[2], &s12, &c12);
  const double a = lc - cz, e = cy + d;
  const double L = lt * c1 + cx * s12 + a * c12;
  const double dL2 = cx * c12 - a * s12;   // dL/dq2
  const double dL1 = dL2 - lt * s1;        // dL/dq1
  if (p) {
    p[0] = ox - lt * s1 - a * s12 + cx * c12;
    p[1] = oy + e * c0 + L * s0;
    p[2] = e * s0 - L * c0;
  }
  if (jac) {
    jac[0] = 0.0;                    jac[3] = -L;        jac[6] = -(a * c12 + cx * s12);
    jac[1] = -e * s0 + L * c0;       jac[4] = s0 * dL1;  jac[7] = s0 * dL2;
    jac[2] = e * c0 + L * s0;        jac[5] = -c0 * dL1; jac[8] = -c0 * dL2;
  }
}

// tau = -J^T f (BaseInterface.cpp:379, :398); zero for a planned swing leg when movement_mode > 0 (:381)
QMPC_HD inline void leg_torque(const double* jac /* 3x3 column-major */, const double* f, bool active, double* tau) {
  for (int j = 0; j < 3; ++j) {
    double t = 0;
    for (int a = 0; a < 3; ++a) t += -jac[3 * j + a] * f[a];
    tau[j] = active ? t : 0.0;
  }
}

// ---- N3: QuatMpc::goal_update (QuatMpc.cpp:68-107) with its per-robot state on the device ----------
// State layout: element-major [field][robot] doubles (stride = robots the handle was created for), so
// consecutive threads touch consecutive addresses.  Fields: desired world position (3), its
// initialised flag, samples seen n, then for each of the six MovingWindowFilter(100) channels
// (lin-vel x,y,z then pos x,y,z; QuatMpc.cpp:9-12): Neumaier sum, correction, ring of 100 samples.
constexpr int kGoalWindow = 100;
constexpr int kGoalFields = 5 + 6 * (2 + kGoalWindow);
struct GoalStateRef {   // view of one robot's state
  double* base;
  size_t stride;
  QMPC_HD double& at(int field) const { return base[(size_t)field * stride]; }
};

// Utils::quat_to_euler yaw (Utils.cpp:29-31), q = (w, x, y, z)
QMPC_HD inline double quat_yaw(const double* q) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double t3 = +2.0 * (w * z + x * y);
  const double t4 = +1.0 - 2.0 * (y * y + z * z);
  return atan2(t3, t4);
}

// MovingWindowFilter::CalculateAverage (MovingWindowFilter.hpp:26-39): Neumaier sum over the last 100 samples
QMPC_HD inline void neumaier_add(double& sum, double& corr, double v) {
  const double ns = add_rn(sum, v);
  if (fabs(sum) >= fabs(v)) corr = add_rn(corr, add_rn(add_rn(sum, -ns), v));
  else corr = add_rn(corr, add_rn(add_rn(v, -ns), sum));
  sum = ns;
}
// One channel: the caller has loaded (sum, corr, oldest sample) and stores them back after the update
QMPC_HD inline double window_average(double& sum, double& corr, double oldest, bool full, double v) {
  if (full) neumaier_add(sum, corr, -oldest);   // the left-most value leaves the window first
  neumaier_add(sum, corr, v);
  return add_rn(sum, corr) / (double)kGoalWindow;
}

// All loads first, all stores last: the state / input / output pointers may alias as far as the compiler
// knows, and interleaved loads and stores would serialise six DRAM round trips per robot.
QMPC_HD inline void goal_update_one(const GoalStateRef& s, const QmpcGoalInput& gin, QmpcProblem& out) {
  const QmpcGoalInput in = gin;
  double pdw[3] = {s.at(0), s.at(1), s.at(2)};
  const bool inited = s.at(3) != 0.0;
  const long long n = (long long)s.at(4);
  const int slot = (int)(n % kGoalWindow);
  const bool full = n >= kGoalWindow;
  double sum[6], corr[6], oldest[6], val[6], avg[6];
#pragma unroll
  for (int ch = 0; ch < 6; ++ch) {
    const int f0 = 5 + ch * (2 + kGoalWindow);
    sum[ch] = s.at(f0);
    corr[ch] = s.at(f0 + 1);
    oldest[ch] = s.at(f0 + 2 + slot);
  }
  double R[9], cy, sy;
  quat_to_rot(in.torso_quat, R);                       // fbk.torso_rot_mat   BaseInterface.cpp:196
  quat_yaw_cs(in.torso_quat, &cy, &sy);                // torso_rot_mat_z = AngleAxis(fbk.torso_euler[2], UnitZ)  :197-200
  if (!inited) {                                       // torso_pos_d_world_init  QuatMpc.cpp:74-77
    for (int i = 0; i < 3; ++i) pdw[i] = in.torso_pos_world[i];
  }
  const double vrel[3] = {in.joy_vel[0], in.joy_vel[1], 0.0};                    // :80-82
  const double vw[3] = {cy * vrel[0] - sy * vrel[1], sy * vrel[0] + cy * vrel[1], vrel[2]};   // :84
  for (int i = 0; i < 3; ++i) val[i] = R[i] * vw[0] + R[3 + i] * vw[1] + R[6 + i] * vw[2];    // R^T v  :85
  pdw[0] = pdw[0] + vw[0] * 5.0 / 1000.0;              // :98-100
  pdw[1] = pdw[1] + vw[1] * 5.0 / 1000.0;
  pdw[2] = in.joy_body_height;
  const double dp[3] = {pdw[0] - in.torso_pos_world[0], pdw[1] - in.torso_pos_world[1], pdw[2] - in.torso_pos_world[2]};
  for (int i = 0; i < 3; ++i) val[3 + i] = R[i] * dp[0] + R[3 + i] * dp[1] + R[6 + i] * dp[2];    // :102
#pragma unroll
  for (int ch = 0; ch < 6; ++ch) avg[ch] = window_average(sum[ch], corr[ch], oldest[ch], full, val[ch]);   // :86-89, :103-106
  // ---- stores
  for (int i = 0; i < 3; ++i) s.at(i) = pdw[i];
  s.at(3) = 1.0;
  s.at(4) = (double)(n + 1);
#pragma unroll
  for (int ch = 0; ch < 6; ++ch) {
    const int f0 = 5 + ch * (2 + kGoalWindow);
    s.at(f0) = sum[ch];
    s.at(f0 + 1) = corr[ch];
    s.at(f0 + 2 + slot) = val[ch];
  }
  for (int i = 0; i < 4; ++i) out.torso_quat[i] = in.torso_quat[i];
  for (int i = 0; i < 3; ++i) out.torso_lin_vel_world[i] = in.torso_lin_vel_world[i];
  for (int i = 0; i < 3; ++i) out.torso_pos_d_body[i] = avg[3 + i];
  for (int i = 0; i < 3; ++i) out.torso_lin_vel_d_body[i] = avg[i];
  for (int i = 0; i < 3; ++i) out.torso_ang_vel_d_body[i] = in.joy_ang_rate[i];                  // :93-95
}

// Raibert heuristic foot-hold targets (BaseInterface.cpp:265-288)
QMPC_HD inline void raibert_one(const QmpcRaibertParams& rp, const QmpcGoalInput& in, double* tgt_world, double* tgt_rel) {
  double R[9];
  quat_to_rot(in.torso_quat, R);
  double cy, sy;
  quat_yaw_cs(in.torso_quat, &cy, &sy);
  // torso_lin_vel_rel = Rz^T v_world
  const double v0 = cy * in.torso_lin_vel_world[0] + sy * in.torso_lin_vel_world[1];
  const double v1 = -sy * in.torso_lin_vel_world[0] + cy * in.torso_lin_vel_world[1];
  const double k = sqrt(fabs(in.torso_pos_world[2]) / 9.81);
  double d0 = k * (v0 - in.joy_vel[0]) + (1.0 / rp.gait_freq) / 2.0 * in.joy_vel[0];
  if (d0 < -rp.delta_x_limit) d0 = -rp.delta_x_limit;
  if (d0 > rp.delta_x_limit) d0 = rp.delta_x_limit;
  double d1 = k * (v1 - in.joy_vel[1]) + (1.0 / rp.gait_freq) / 2.0 * in.joy_vel[1];
  if (d1 < -rp.delta_y_limit) d1 = -rp.delta_y_limit;
  if (d1 > rp.delta_y_limit) d1 = rp.delta_y_limit;
  const double a0 = cy * d0 - sy * d1, a1 = sy * d0 + cy * d1;       // raibert_delta_abs = Rz delta_rel
  for (int i = 0; i < 4; ++i) {
    const double* p = rp.default_foot_pos_rel + 3 * i;
    const double abs0 = cy * p[0] - sy * p[1] + a0, abs1 = sy * p[0] + cy * p[1] + a1, abs2 = p[2];
    if (tgt_rel)
      for (int a = 0; a < 3; ++a) tgt_rel[3 * i + a] = R[a] * abs0 + R[3 + a] * abs1 + R[6 + a] * abs2;
    if (tgt_world) {
      tgt_world[3 * i] = abs0 + in.torso_pos_world[0];
      tgt_world[3 * i + 1] = abs1 + in.torso_pos_world[1];
      tgt_world[3 * i + 2] = abs2 + in.torso_pos_world[2];
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Row N3, gait-FSM half: LeggedContactFSM::update / reset for the four legs of every robot
// (LeggedContactFSM.cpp:10-78, 208-260) as driven by QuatMpc::foot_update (QuatMpc.cpp:278-305), and the quintic
// swing curve (Utils.cpp:236-293).  Per-leg state lives on the device, element-major [field][4 robot + leg].
constexpr int kFsmFields = 27;
namespace fsm {
constexpr int s = 0, phase = 1, idx = 2, prev = 3, start = 4, end = 5, called = 6, swing_start = 7, swing_end = 10,
              swing_extend = 13, pos = 16, vel = 19, acc = 22, terrain = 25, gait = 26;
}
struct FsmRef {
  double* p;
  size_t stride;
  QMPC_HD double& at(int f) const { return p[(size_t)f * stride]; }
};

// C matrix of QuinticCurve::get_foot_swing_target (Utils.cpp:238-244): T is a FLOAT there, its powers are
// float products, each entry is then stored as double.  Row-major 6x6.
QMPC_HD inline void quintic_C(float T, double* C) {
  const double c[36] = {1, 0, 0, 0, 0, 0,
                        1, T, T * T, T * T * T, T * T * T * T, T * T * T * T * T,
                        0, 1, 0, 0, 0, 0,
                        0, 1, 2 * T, 3 * T * T, 4 * T * T * T, 5 * T * T * T * T,
                        1, T / 2, T * T / 4, T * T * T / 8, T * T * T * T / 16, T * T * T * T * T / 32,
                        0, 1, T, 3 * T * T / 4, 4 * T * T * T / 8, 5 * T * T * T * T / 16};
  for (int i = 0; i < 36; ++i) C[i] = c[i];
}
// C.inverse() (Eigen: PartialPivLU, then solve against the identity): host-side helper, evaluated once per launch
// - C depends on the gait frequency only
inline bool quintic_C_inverse(float T, double* Cinv) {
  double A[36];
  int perm[6];
  quintic_C(T, A);
  for (int i = 0; i < 6; ++i) perm[i] = i;
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = fabs(A[6 * k + k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[6 * i + k]) > best) { best = fabs(A[6 * i + k]); piv = i; }
    if (best == 0.0) return false;
    if (piv != k) {
      for (int j = 0; j < 6; ++j) { const double t = A[6 * k + j]; A[6 * k + j] = A[6 * piv + j]; A[6 * piv + j] = t; }
      const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    for (int i = k + 1; i < 6; ++i) {
      A[6 * i + k] /= A[6 * k + k];
      for (int j = k + 1; j < 6; ++j) A[6 * i + j] -= A[6 * i + k] * A[6 * k + j];
    }
  }
  for (int c = 0; c < 6; ++c) {
    double x[6];
    for (int i = 0; i < 6; ++i) x[i] = perm[i] == c ? 1.0 : 0.0;
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < i; ++j) x[i] -= A[6 * i + j] * x[j];
    for (int i = 5; i >= 0; --i) {
      for (int j = i + 1; j < 6; ++j) x[i] -= A[6 * i + j] * x[j];
      x[i] /= A[6 * i + i];
    }
    for (int i = 0; i < 6; ++i) Cinv[6 * i + c] = x[i];
  }
  return true;
}
struct QuinticInv { double m[36]; };

// position / velocity / acceleration of one axis: a = Cinv con, then the three polynomials exactly as written
// at Utils.cpp:262-264 (no FMA: the products and sums of the reference, one rounding each)
QMPC_HD inline void quintic_axis(const QuinticInv& ci, const double* con, double t, double* p, double* v, double* a) {
  double c[6];
  for (int i = 0; i < 6; ++i) {
    double acc = 0.0;
    for (int j = 0; j < 6; ++j) acc = add_rn(acc, mul_rn(ci.m[6 * i + j], con[j]));
    c[i] = acc;
  }
  const double t2 = mul_rn(t, t);   // the reference evaluates a*t*t*... left to right: ((a t) t) t
  (void)t2;
  auto pw = [&](double coef, int n) { double r = coef; for (int i = 0; i < n; ++i) r = mul_rn(r, t); return r; };
  *p = add_rn(add_rn(add_rn(add_rn(add_rn(c[0], pw(c[1], 1)), pw(c[2], 2)), pw(c[3], 3)), pw(c[4], 4)), pw(c[5], 5));
  *v = add_rn(add_rn(add_rn(add_rn(c[1], pw(mul_rn(2, c[2]), 1)), pw(mul_rn(3, c[3]), 2)), pw(mul_rn(4, c[4]), 3)), pw(mul_rn(5, c[5]), 4));
  *a = add_rn(add_rn(add_rn(mul_rn(2, c[2]), pw(mul_rn(6, c[3]), 1)), pw(mul_rn(12, c[4]), 2)), pw(mul_rn(20, c[5]), 3));
}

// QuinticCurve::get_foot_swing_target(t, T, start, final) -> pos[3], vel[3], acc[3]
QMPC_HD inline void quintic_swing_target(const QuinticInv& ci, float tf, float T, const double* p0, const double* pT,
                                         double* pos, double* vel, double* acc) {
  const double t = (double)tf;
  const double dx = pT[0] - p0[0], dy = pT[1] - p0[1];
  const double k = 1.26 / (double)T;
  const double v_xy_mid = mul_rn(k, sqrt(add_rn(mul_rn(dx, dx), mul_rn(dy, dy))));
  const double theta = atan2(fabs(dy), fabs(dx));
  const double v_x_mid = mul_rn(mul_rn((dx >= 0 ? 1.0 : -1.0), v_xy_mid), cos(theta));
  const double v_y_mid = mul_rn(mul_rn((dy >= 0 ? 1.0 : -1.0), v_xy_mid), sin(theta));
  const double zc[6] = {p0[2], pT[2], 0.1, -0.1, 0.1, 0.0};
  const double xc[6] = {p0[0], pT[0], 0.0, 0.0, add_rn(p0[0], pT[0]) / 2, v_x_mid};
  const double yc[6] = {p0[1], pT[1], 0.0, 0.0, add_rn(p0[1], pT[1]) / 2, v_y_mid};
  quintic_axis(ci, xc, t, pos + 0, vel + 0, acc + 0);
  quintic_axis(ci, yc, t, pos + 1, vel + 1, acc + 1);
  quintic_axis(ci, zc, t, pos + 2, vel + 2, acc + 2);
}

QMPC_HD inline double fsm_percent(double phase, double start, double end) {
  double pct = add_rn(phase, -start) / add_rn(end, -start);
  if (pct < 0.0) pct = 0.0;
  else if (pct > 1.0) pct = 1.0;
  return pct;
}

// reset_params (trot unless `gait` says otherwise) + reset: a freshly constructed controller
QMPC_HD inline void leg_fsm_init_one(const FsmRef& st, int leg, int gait) {
  const GaitLegPattern pat = gait_pattern(gait, leg);
  for (int f = 0; f < kFsmFields; ++f) st.at(f) = 0.0;
  st.at(fsm::gait) = (double)gait;
  st.at(fsm::prev) = (double)(pat.n - 1);
  st.at(fsm::end) = pat.sw[0];
  st.at(fsm::s) = (double)pat.stance[0];
}

// One tick of QuatMpc::foot_update for one leg (QuatMpc.cpp:278-305): movement_mode 0 -> reset() and
// plan_contact = true; else update(dt, gait_freq, cur, target, foot_force_flag).
QMPC_HD inline void leg_fsm_tick_one(const FsmRef& st, const QuinticInv& ci, int leg, int movement_mode, double dt, double gait_freq,
                                     const double* cur, const double* tgt, bool flag, double* pos, double* vel, double* acc,
                                     double* gait_counter, int* contact) {
  const GaitLegPattern pat = gait_pattern((int)st.at(fsm::gait), leg);
  int s = (int)st.at(fsm::s), idx = (int)st.at(fsm::idx), prev = (int)st.at(fsm::prev);
  double phase = st.at(fsm::phase), start = st.at(fsm::start), end = st.at(fsm::end);
  bool called = st.at(fsm::called) != 0.0;
  double sw0[3], sw1[3], ext[3], P[3], V[3], A[3];
  for (int a = 0; a < 3; ++a) {
    sw0[a] = st.at(fsm::swing_start + a); sw1[a] = st.at(fsm::swing_end + a); ext[a] = st.at(fsm::swing_extend + a);
    P[a] = st.at(fsm::pos + a); V[a] = st.at(fsm::vel + a); A[a] = st.at(fsm::acc + a);
  }
  double terrain = st.at(fsm::terrain);
  if (movement_mode == 0) {
    // LeggedContactFSM::reset (LeggedContactFSM.cpp:10-31)
    phase = 0; idx = 0; prev = pat.n - 1; start = 0; end = pat.sw[0];
    if (s == 0) { for (int a = 0; a < 3; ++a) { P[a] = sw1[a]; V[a] = 0.0; } }
    s = pat.stance[0];
    called = false;
    *gait_counter = phase;      // ctrl.gait_counter is not written in this branch; reported as the reset phase
    *contact = 1;               // QuatMpc.cpp:288
  } else {
    // LeggedContactFSM::update (LeggedContactFSM.cpp:33-78)
    if (!called) {
      for (int a = 0; a < 3; ++a) { sw0[a] = cur[a]; sw1[a] = tgt[a]; P[a] = tgt[a]; V[a] = 0.0; }
      called = true;
    }
    phase = add_rn(phase, mul_rn(gait_freq, dt));
    auto common_enter = [&]() {   // :208-223
      prev = idx;
      idx = (idx + 1) % pat.n;
      if (idx < prev) phase = add_rn(phase, -1.0);
      start = phase;
      end = pat.sw[idx];
    };
    if (s == 1) {
      if (phase >= end) {
        terrain = cur[2];                                            // stance_exit :80-84
        common_enter();                                              // swing_enter :225-229
        for (int a = 0; a < 3; ++a) { sw0[a] = cur[a]; ext[a] = 0.0; }
        s = 0;
      }
    } else {
      const double pct = fsm_percent(phase, start, end);
      if ((pct > 0.9 && flag) || pct >= 1.0) {                       // early contact / end of swing :53-64
        s = 1;
        common_enter();                                              // stance_enter :231-235
        for (int a = 0; a < 3; ++a) { P[a] = cur[a]; V[a] = 0.0; }
      }
    }
    if (s == 0) {                                                    // swing_update :237-246
      const double t = fsm_percent(phase, start, end);
      const double fin[3] = {add_rn(tgt[0], ext[0]), add_rn(tgt[1], ext[1]), add_rn(tgt[2], ext[2])};
      quintic_swing_target(ci, (float)(mul_rn(0.5, t) / gait_freq), (float)(0.5 / gait_freq), sw0, fin, P, V, A);
    }                                                                // stance_update is empty (:248-259)
    *gait_counter = phase;
    *contact = s;
  }
  for (int a = 0; a < 3; ++a) { pos[a] = P[a]; vel[a] = V[a]; acc[a] = A[a]; }
  st.at(fsm::s) = (double)s; st.at(fsm::idx) = (double)idx; st.at(fsm::prev) = (double)prev;
  st.at(fsm::phase) = phase; st.at(fsm::start) = start; st.at(fsm::end) = end; st.at(fsm::called) = called ? 1.0 : 0.0;
  for (int a = 0; a < 3; ++a) {
    st.at(fsm::swing_start + a) = sw0[a]; st.at(fsm::swing_end + a) = sw1[a]; st.at(fsm::swing_extend + a) = ext[a];
    st.at(fsm::pos + a) = P[a]; st.at(fsm::vel + a) = V[a]; st.at(fsm::acc + a) = A[a];
  }
  st.at(fsm::terrain) = terrain;
}

#ifdef __CUDACC__
// one thread per (robot, leg)
__global__ void __launch_bounds__(256)
qmpc_leg_fsm_init_kernel(double* __restrict__ state, size_t stride, const int32_t* __restrict__ gait, int batch) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * batch) return;
  leg_fsm_init_one(FsmRef{state + t, stride}, t & 3, gait ? gait[t >> 2] : QMPC_GAIT_TROT);
}

__global__ void __launch_bounds__(256)
qmpc_foot_update_kernel(double* __restrict__ state, size_t stride, QuinticInv ci, const QmpcFootUpdateInput* __restrict__ in,
                        double dt, double gait_freq, int batch, QmpcFootUpdateOutput* __restrict__ out,
                        QmpcProblem* __restrict__ problems, QmpcGaitState* __restrict__ gait_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * batch) return;
  const int b = t >> 2, leg = t & 3;
  const QmpcFootUpdateInput* r = in + b;
  double cur[3], tgt[3], pos[3], vel[3], acc[3], gc;
  int contact;
#pragma unroll
  for (int a = 0; a < 3; ++a) { cur[a] = r->foot_pos_world[3 * leg + a]; tgt[a] = r->foot_pos_target_world[3 * leg + a]; }
  const FsmRef st{state + t, stride};
  leg_fsm_tick_one(st, ci, leg, r->movement_mode, dt, gait_freq, cur, tgt, r->foot_contact_flag[leg] != 0, pos, vel, acc, &gc,
                   &contact);
  QmpcFootUpdateOutput* o = out + b;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    o->foot_pos_target[3 * leg + a] = pos[a]; o->foot_vel_target[3 * leg + a] = vel[a]; o->foot_acc_target[3 * leg + a] = acc[a];
  }
  o->gait_counter[leg] = gc;
  o->plan_contacts[leg] = contact;
  if (problems) problems[b].plan_contacts[leg] = contact;
  if (gait_out) {
    gait_out[b].gait_phase[leg] = gc;
    if (leg == 0) { gait_out[b].gait_freq = gait_freq; gait_out[b].gait = (int32_t)st.at(fsm::gait); gait_out[b].pad_ = 0; }
  }
}

// One thread per robot, 128 robots per block; the filter state is element-major (consecutive threads, consecutive
// addresses).  The array-of-structures sides go through shared memory like the Raibert kernel's: the block's 128 input
// records come in with consecutive threads on consecutive addresses, and the 16 words goal_update writes into each
// QmpcProblem (three runs: words 0..6, 22..27, 32..34 of the 37) go out run by run instead of field by field.
constexpr int kGoalBlock = 128;
__global__ void __launch_bounds__(kGoalBlock)
qmpc_goal_update_kernel(double* __restrict__ state, size_t stride, const QmpcGoalInput* __restrict__ in, int batch,
                        QmpcProblem* __restrict__ problems) {
  constexpr int kIn = (int)(sizeof(QmpcGoalInput) / 8), kInPad = kIn + 1, kOut = 16, kOutPad = 17;
  constexpr int kProb = (int)(sizeof(QmpcProblem) / 8);
  static_assert(sizeof(QmpcProblem) % 8 == 0 && offsetof(QmpcProblem, torso_quat) == 0 &&
                offsetof(QmpcProblem, torso_lin_vel_world) == 32 && offsetof(QmpcProblem, torso_pos_d_body) == 176 &&
                offsetof(QmpcProblem, torso_lin_vel_d_body) == 200 && offsetof(QmpcProblem, torso_ang_vel_d_body) == 256,
                "QmpcProblem layout changed: update the output runs below");
  __shared__ double s_in[kGoalBlock * kInPad];
  __shared__ double s_out[kGoalBlock * kOutPad];
  const size_t i0 = (size_t)blockIdx.x * kGoalBlock;
  const int cnt = (int)((size_t)batch - i0 < (size_t)kGoalBlock ? (size_t)batch - i0 : (size_t)kGoalBlock);
  const double* src = reinterpret_cast<const double*>(in + i0);
  for (int e = threadIdx.x; e < cnt * kIn; e += kGoalBlock) s_in[(e / kIn) * kInPad + e % kIn] = src[e];
  __syncthreads();
  if ((int)threadIdx.x < cnt) {
    QmpcGoalInput rec;
    double* r = reinterpret_cast<double*>(&rec);
#pragma unroll
    for (int e = 0; e < kIn; ++e) r[e] = s_in[threadIdx.x * kInPad + e];
    QmpcProblem tmp;
    goal_update_one(GoalStateRef{state + i0 + threadIdx.x, stride}, rec, tmp);
    double* o = s_out + threadIdx.x * kOutPad;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = tmp.torso_quat[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      o[4 + j] = tmp.torso_lin_vel_world[j];
      o[7 + j] = tmp.torso_pos_d_body[j];
      o[10 + j] = tmp.torso_lin_vel_d_body[j];
      o[13 + j] = tmp.torso_ang_vel_d_body[j];
    }
  }
  __syncthreads();
  double* dst = reinterpret_cast<double*>(problems + i0);
  for (int e = threadIdx.x; e < cnt * kOut; e += kGoalBlock) {
    const int r = e / kOut, j = e % kOut;
    const int word = j < 7 ? j : (j < 13 ? 22 + (j - 7) : 32 + (j - 13));
    dst[(size_t)r * kProb + word] = s_out[r * kOutPad + j];
  }
}

// One thread per robot, 128 robots per block.  Records in (128-byte QmpcGoalInput) and out (two 96-byte target sets) are
// array-of-structures: read / written by the threads themselves every access scatters 32 x 8 bytes over 4 KB / 3 KB.
// The block instead copies its 128 input records into shared memory with consecutive threads on consecutive
// addresses (row stride padded to 17 doubles: conflict-free), every thread works on its record there, and the
// 128 x 12 results of each output go out the same way.
constexpr int kRaibertBlock = 128;
__global__ void __launch_bounds__(kRaibertBlock)
qmpc_raibert_kernel(QmpcRaibertParams rp, const QmpcGoalInput* __restrict__ in, int batch,
                    double* __restrict__ tgt_world, double* __restrict__ tgt_rel) {
  constexpr int kIn = (int)(sizeof(QmpcGoalInput) / 8), kInPad = kIn + 1;
  static_assert(sizeof(QmpcGoalInput) % 8 == 0, "QmpcGoalInput is copied as 8-byte words");
  __shared__ double s_in[kRaibertBlock * kInPad];
  __shared__ double s_w[kRaibertBlock * 13], s_r[kRaibertBlock * 13];
  const size_t i0 = (size_t)blockIdx.x * kRaibertBlock;
  const int cnt = (int)((size_t)batch - i0 < (size_t)kRaibertBlock ? (size_t)batch - i0 : (size_t)kRaibertBlock);
  const double* src = reinterpret_cast<const double*>(in + i0);
  for (int e = threadIdx.x; e < cnt * kIn; e += kRaibertBlock) s_in[(e / kIn) * kInPad + e % kIn] = src[e];
  __syncthreads();
  if ((int)threadIdx.x < cnt) {
    QmpcGoalInput rec;
    double* r = reinterpret_cast<double*>(&rec);
#pragma unroll
    for (int e = 0; e < kIn; ++e) r[e] = s_in[threadIdx.x * kInPad + e];
    raibert_one(rp, rec, s_w + 13 * threadIdx.x, s_r + 13 * threadIdx.x);
  }
  __syncthreads();
  if (tgt_world)
    for (int e = threadIdx.x; e < cnt * 12; e += kRaibertBlock) tgt_world[12 * i0 + e] = s_w[(e / 12) * 13 + e % 12];
  if (tgt_rel)
    for (int e = threadIdx.x; e < cnt * 12; e += kRaibertBlock) tgt_rel[12 * i0 + e] = s_r[(e / 12) * 13 + e % 12];
}

// One thread per (robot, leg): 32 stance bits of its leg, exchanged inside the quad by shuffles; lane `leg`
// of the quad then stores knots 8 leg .. 8 leg + 7 (one 8-byte store, 32 contiguous bytes per robot).
__global__ void __launch_bounds__(256)
qmpc_predict_schedule_kernel(const QmpcGaitState* __restrict__ g, int batch, int N, double dt,
                             QmpcContactSchedule* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int leg = t & 3;
  const int robot = (t >> 2) < batch ? (t >> 2) : batch - 1;   // tail threads shadow the last robot (full-warp shuffles)
  const QmpcGaitState* gs = g + robot;
  const GaitLegPattern pat = gait_pattern(gs->gait, leg);
  const double phase = gs->gait_phase[leg], freq = gs->gait_freq;
  unsigned bits = 0;
  for (int k = 0; k < N; ++k) bits |= (unsigned)predict_contact(pat, phase, freq, mul_rn((double)k, dt)) << k;
  const unsigned b0 = __shfl_sync(0xffffffffu, bits, (threadIdx.x & ~3) + 0);
  const unsigned b1 = __shfl_sync(0xffffffffu, bits, (threadIdx.x & ~3) + 1);
  const unsigned b2 = __shfl_sync(0xffffffffu, bits, (threadIdx.x & ~3) + 2);
  const unsigned b3 = __shfl_sync(0xffffffffu, bits, (threadIdx.x & ~3) + 3);
  unsigned long long w = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = 8 * leg + j;
    const unsigned m = ((b0 >> k) & 1u) | (((b1 >> k) & 1u) << 1) | (((b2 >> k) & 1u) << 2) | (((b3 >> k) & 1u) << 3);
    w |= (unsigned long long)m << (8 * j);
  }
  if ((t >> 2) < batch) reinterpret_cast<unsigned long long*>(out + robot)[leg] = w;
}

// One thread per (robot, leg).  The 3 + 9 results of a thread are 24- and 72-byte records: written by the threads
// themselves every store instruction scatters 32 x 8 bytes over 768 / 2304 bytes (4 partially written sectors per
// sector's worth of data; the kernel sat at 0.40 of the HBM roof).  They are staged in shared memory instead and the
// block writes its 768 + 2304 contiguous doubles with consecutive threads on consecutive addresses.
__global__ void __launch_bounds__(256)
qmpc_leg_kinematics_kernel(QmpcLegParams lp, const double* __restrict__ joint_pos, int batch,
                           double* __restrict__ foot_pos_body, double* __restrict__ jac_foot) {
  __shared__ double sf[256 * 3];
  __shared__ double sj[256 * 9];
  const size_t t0 = (size_t)blockIdx.x * blockDim.x;
  const size_t t = t0 + threadIdx.x;   // (robot, leg)
  const size_t total = (size_t)batch * 4;
  const int cnt = (int)(total - t0 < blockDim.x ? total - t0 : blockDim.x);
  if (t < total) {
    const int leg = (int)(t & 3);
    const double q[3] = {joint_pos[3 * t], joint_pos[3 * t + 1], joint_pos[3 * t + 2]};
    double p[3], J[9];
    leg_fk_jac(q, lp.rho_fix[leg], lp.rho_opt[leg], p, J);
#pragma unroll
    for (int a = 0; a < 3; ++a) sf[3 * threadIdx.x + a] = p[a];
#pragma unroll
    for (int a = 0; a < 9; ++a) sj[9 * threadIdx.x + a] = J[a];
  }
  __syncthreads();
  if (foot_pos_body)
    for (int i = threadIdx.x; i < 3 * cnt; i += blockDim.x) foot_pos_body[3 * t0 + i] = sf[i];
  if (jac_foot)
    for (int i = threadIdx.x; i < 9 * cnt; i += blockDim.x) jac_foot[9 * t0 + i] = sj[i];
}

__global__ void __launch_bounds__(256)
qmpc_joint_torque_kernel(const QmpcResult* __restrict__ res, const double* __restrict__ jac_foot,
                         const int32_t* __restrict__ plan_contacts, int movement_mode, int batch,
                         double* __restrict__ tau) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // (robot, leg)
  if (t >= batch * 4) return;
  const int b = t >> 2, leg = t & 3;
  double J[9], f[3], out[3];
#pragma unroll
  for (int a = 0; a < 9; ++a) J[a] = jac_foot[9 * (size_t)t + a];
#pragma unroll
  for (int a = 0; a < 3; ++a) f[a] = res[b].grf_body[3 * leg + a];   // ctrl.optimized_input[3 leg ..]
  const bool active = movement_mode <= 0 || !plan_contacts || plan_contacts[t] != 0;
  leg_torque(J, f, active, out);
#pragma unroll
  for (int a = 0; a < 3; ++a) tau[3 * (size_t)t + a] = out[a];
}
#endif

}  // namespace qmpc
