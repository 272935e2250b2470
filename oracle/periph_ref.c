/* periph_ref.c — CPU restatement of the reference code either side of the MPC solve
 * (SURVEY.md 8f rows N1, N2).  TEST INFRASTRUCTURE ONLY (see altro_ref.h): loaded by tests/ and
 * bench.py's checker legs, never by the product.
 *
 *   N1  LeggedContactFSM::predict_contact_state      legged_ctrl/src/utils/LeggedContactFSM.cpp:272-286
 *       gait pattern tables                          LeggedContactFSM.cpp:87-206
 *   N2  A1Kinematics::fk / jac (symbolic-toolbox polynomials, restated term by term in the
 *       reference's summation order)                 legged_ctrl/src/utils/A1Kinematics.cpp:40-129
 *       BaseInterface::tau_ctrl_update torque rows   legged_ctrl/src/interfaces/BaseInterface.cpp:343-405
 */
#include <math.h>
#include <string.h>

#include "../include/qmpc.h"

/* ---------------------------------------------------------------- N1: gait FSM predictor */
enum { SWING = 0, STANCE = 1 };
typedef struct LegFsm {
  int gait_pattern_size;
  int gait_state_pattern[3];
  double gait_switch_time[3];
  double gait_phase, gait_freq;
} LegFsm;

static void push(LegFsm* f, int state, double sw) {
  f->gait_state_pattern[f->gait_pattern_size] = state;
  f->gait_switch_time[f->gait_pattern_size] = sw;
  f->gait_pattern_size++;
}

/* set_default_gait_pattern / set_trot_with_stand_gait_pattern / set_crawl_gait_pattern /
 * set_default_stand_pattern, keyed by leg_id (LeggedContactFSM.cpp:87-206) */
static void set_pattern(LegFsm* f, int gait, int leg_id) {
  f->gait_pattern_size = 0;
  if (gait == QMPC_GAIT_TROT) {
    if (leg_id == 0 || leg_id == 3) { push(f, STANCE, 0.5); push(f, SWING, 1.0); }
    else { push(f, SWING, 0.5); push(f, STANCE, 1.0); }
  } else if (gait == QMPC_GAIT_TROT_WITH_STAND) {
    if (leg_id == 0 || leg_id == 3) { push(f, STANCE, 0.6); push(f, SWING, 1.0); }
    else { push(f, STANCE, 0.1); push(f, SWING, 0.5); push(f, STANCE, 1.0); }
  } else if (gait == QMPC_GAIT_CRAWL) {
    if (leg_id == 0) { push(f, SWING, 0.25); push(f, STANCE, 1.0); }
    else if (leg_id == 1) { push(f, STANCE, 0.25); push(f, SWING, 0.5); push(f, STANCE, 1.0); }
    else if (leg_id == 2) { push(f, STANCE, 0.5); push(f, SWING, 0.75); push(f, STANCE, 1.0); }
    else { push(f, STANCE, 0.75); push(f, SWING, 1.0); }
  } else {
    push(f, STANCE, 1.0);
  }
}

/* LeggedContactFSM::predict_contact_state (LeggedContactFSM.cpp:272-286) */
static int predict_contact_state(const LegFsm* f, double dt) {
  double predicted_gait_phase = f->gait_phase + f->gait_freq * dt;
  while (predicted_gait_phase > 1.0) predicted_gait_phase -= 1.0;
  for (int i = 0; i < f->gait_pattern_size; i++)
    if (predicted_gait_phase <= f->gait_switch_time[i]) return f->gait_state_pattern[i];
  return STANCE;
}

int qmpc_ref_predict_schedule(const QmpcConfig* cfg, const QmpcGaitState* g, int batch, QmpcContactSchedule* out) {
  if (!cfg || !g || !out || batch < 0) return QMPC_ERR_ARG;
  for (int b = 0; b < batch; ++b) {
    LegFsm leg[4];
    for (int i = 0; i < 4; ++i) {
      set_pattern(&leg[i], g[b].gait, i);
      leg[i].gait_phase = g[b].gait_phase[i];
      leg[i].gait_freq = g[b].gait_freq;
    }
    memset(&out[b], 0, sizeof(out[b]));
    for (int k = 0; k < cfg->horizon; ++k) {
      int m = 0;
      for (int i = 0; i < 4; ++i)
        if (predict_contact_state(&leg[i], k * cfg->dt) == STANCE) m |= 1 << i;
      out[b].mask[k] = (uint8_t)m;
    }
  }
  return QMPC_OK;
}

/* ---------------------------------------------------------------- N2: leg kinematics */
/* fk: A1Kinematics.cpp:40-75.  c = rho_opt (3), r = rho_fix (5) */
static void fk_ref(const double q[3], const double c[3], const double r[5], double p[3]) {
  const double c0 = cos(q[0]), c1 = cos(q[1]), c2 = cos(q[2]);
  const double s0 = sin(q[0]), s1 = sin(q[1]), s2 = sin(q[2]);
  const double q12 = q[1] + q[2];
  const double s12 = sin(q12);
  p[0] = (((r[0] + c[2] * s12) - r[4] * s12) - s1 * r[3]) + c[0] * cos(q12);
  p[1] = ((((((((r[1] + c[1] * c0) + r[2] * c0) + c1 * s0 * r[3]) + c[0] * c1 * s0 * s2) + c[0] * c2 * s0 * s1) -
            c[2] * c1 * c2 * s0) + c[2] * s0 * s1 * s2) + r[4] * c1 * c2 * s0) - r[4] * s0 * s1 * s2;
  const double a = c[0] * c0, b = c[2] * c0, e = r[4] * c0;
  p[2] = (((((((c[1] * s0 + r[2] * s0) - c0 * c1 * r[3]) - a * c1 * s2) - a * c2 * s1) + b * c1 * c2) - b * s1 * s2) -
          e * c1 * c2) + e * s1 * s2;
}

/* jac (column-major 3x3): A1Kinematics.cpp:77-129 */
static void jac_ref(const double q[3], const double c[3], const double r[5], double J[9]) {
  const double c0 = cos(q[0]), c1 = cos(q[1]), c2 = cos(q[2]);
  const double s0 = sin(q[0]), s1 = sin(q[1]), s2 = sin(q[2]);
  const double q12 = q[1] + q[2];
  const double c12 = cos(q12), s12 = sin(q12);
  const double t12 = c[0] * c12, t16 = c[2] * s12, t17 = r[4] * s12;
  const double t22 = (t12 + t16) + -t17;
  J[0] = 0.0;
  double a = c[0] * c0, b = c[2] * c0, e = r[4] * c0;
  J[1] = (((((((-c[1] * s0 - r[2] * s0) + c0 * c1 * r[3]) + a * c1 * s2) + a * c2 * s1) - b * c1 * c2) + b * s1 * s2) +
          e * c1 * c2) - e * s1 * s2;
  J[2] = (((((((c[1] * c0 + r[2] * c0) + c1 * s0 * r[3]) + c[0] * c1 * s0 * s2) + c[0] * c2 * s0 * s1) -
            c[2] * c1 * c2 * s0) + c[2] * s0 * s1 * s2) + r[4] * c1 * c2 * s0) - r[4] * s0 * s1 * s2;
  a = (c[2] * c12 + -(r[4] * c12)) + -(c[0] * s12);
  J[3] = a - c1 * r[3];
  b = ((s1 * r[3] - t12) - t16) + t17;
  J[4] = -s0 * b;
  J[5] = c0 * b;
  J[6] = a;
  J[7] = s0 * t22;
  J[8] = -c0 * t22;
}

/* BaseInterface.cpp:204-212 for a batch: joint_pos batch x 12 -> foot_pos_body batch x 12 (3x4 col-major),
 * jac_foot batch x 36 (3x12 col-major) */
int qmpc_ref_leg_kinematics(const QmpcLegParams* lp, const double* joint_pos, int batch, double* foot, double* jac) {
  if (!lp || !joint_pos || batch < 0) return QMPC_ERR_ARG;
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < 4; ++i) {
      const double* q = joint_pos + 12 * (size_t)b + 3 * i;
      if (foot) fk_ref(q, lp->rho_opt[i], lp->rho_fix[i], foot + 12 * (size_t)b + 3 * i);
      if (jac) jac_ref(q, lp->rho_opt[i], lp->rho_fix[i], jac + 36 * (size_t)b + 9 * i);
    }
  return QMPC_OK;
}

/* BaseInterface::tau_ctrl_update torque rows (:379-381, :398): tau_i = -jac_i^T * optimized_input_i */
int qmpc_ref_joint_torques(const QmpcResult* res, const double* jac_foot, const int32_t* plan_contacts,
                           int movement_mode, int batch, double* tau) {
  if (!res || !jac_foot || !tau || batch < 0) return QMPC_ERR_ARG;
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < 4; ++i) {
      const double* J = jac_foot + 36 * (size_t)b + 9 * i; /* column-major 3x3 block */
      const double* f = res[b].grf_body + 3 * i;
      double* t = tau + 12 * (size_t)b + 3 * i;
      const int stance = !plan_contacts || plan_contacts[4 * (size_t)b + i];
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int a = 0; a < 3; ++a) s += -J[3 * j + a] * f[a];
        t[j] = (movement_mode > 0 && !stance) ? 0.0 : s;
      }
    }
  return QMPC_OK;
}

/* ---------------------------------------------------------------- N3: reference generation */
/* MovingWindowFilter (include/utils/MovingWindowFilter.hpp:16-71): Neumaier moving-window average.
 * The std::deque is restated as a FIFO ring: front() is the oldest of the window_size_ samples. */
#define MWF_WINDOW 100
typedef struct MovingWindowFilter {
  double sum_, correction_;
  double value_deque_[MWF_WINDOW];
  int size_, front_;
} MovingWindowFilter;

static void UpdateNeumaierSum(MovingWindowFilter* f, double value) {
  double new_sum = f->sum_ + value;
  if (fabs(f->sum_) >= fabs(value)) f->correction_ += (f->sum_ - new_sum) + value;
  else f->correction_ += (value - new_sum) + f->sum_;
  f->sum_ = new_sum;
}
static double CalculateAverage(MovingWindowFilter* f, double new_value) {
  if (f->size_ < MWF_WINDOW) {
    /* pass */
  } else {
    UpdateNeumaierSum(f, -f->value_deque_[f->front_]); /* -value_deque_.front(); pop_front() */
    f->front_ = (f->front_ + 1) % MWF_WINDOW;
    f->size_--;
  }
  UpdateNeumaierSum(f, new_value);
  f->value_deque_[(f->front_ + f->size_) % MWF_WINDOW] = new_value; /* push_back */
  f->size_++;
  return (f->sum_ + f->correction_) / (double)MWF_WINDOW;
}

/* The members of QuatMpc + LeggedState.ctrl that goal_update keeps between ticks */
typedef struct QmpcRefGoalState {
  MovingWindowFilter torso_lin_vel_d_body_filter[3], torso_pos_d_body_filter[3]; /* QuatMpc.cpp:9-12 */
  double torso_pos_d_world[3];
  int torso_pos_d_world_init;
} QmpcRefGoalState;

int qmpc_ref_goal_state_bytes(void) { return (int)sizeof(QmpcRefGoalState); }

/* Eigen::Quaterniond::toRotationMatrix (BaseInterface.cpp:196), row-major */
static void quat_to_rot_ref(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
/* Utils::quat_to_euler, yaw component (Utils.cpp:29-31) */
static double quat_yaw_ref(const double* q) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  double y_sqr = y * y;
  double t3 = +2.0 * (w * z + x * y);
  double t4 = +1.0 - 2.0 * (y_sqr + z * z);
  return atan2(t3, t4);
}

/* QuatMpc::goal_update (QuatMpc.cpp:68-107) for every robot of the batch; state[i] persists across calls
 * (zero-filled = freshly constructed QuatMpc with torso_pos_d_world_init == false) */
int qmpc_ref_goal_update(QmpcRefGoalState* state, const QmpcGoalInput* in, int batch, QmpcProblem* out) {
  if (!state || !in || !out || batch < 0) return QMPC_ERR_ARG;
  for (int b = 0; b < batch; ++b) {
    QmpcRefGoalState* s = &state[b];
    const QmpcGoalInput* g = &in[b];
    double R[9];
    quat_to_rot_ref(g->torso_quat, R);
    const double yaw = quat_yaw_ref(g->torso_quat);
    const double Rz[9] = {cos(yaw), -sin(yaw), 0, sin(yaw), cos(yaw), 0, 0, 0, 1}; /* AngleAxisd(yaw, UnitZ) */
    if (!s->torso_pos_d_world_init) { /* :74-77 */
      memcpy(s->torso_pos_d_world, g->torso_pos_world, sizeof(double) * 3);
      s->torso_pos_d_world_init = 1;
    }
    const double v_rel[3] = {g->joy_vel[0], g->joy_vel[1], 0.0}; /* :80-82 */
    double v_world[3], v_body[3], p_body[3];
    for (int i = 0; i < 3; ++i) v_world[i] = Rz[3 * i] * v_rel[0] + Rz[3 * i + 1] * v_rel[1] + Rz[3 * i + 2] * v_rel[2];
    for (int i = 0; i < 3; ++i) v_body[i] = R[i] * v_world[0] + R[3 + i] * v_world[1] + R[6 + i] * v_world[2];
    for (int i = 0; i < 3; ++i) out[b].torso_lin_vel_d_body[i] = CalculateAverage(&s->torso_lin_vel_d_body_filter[i], v_body[i]);
    for (int i = 0; i < 3; ++i) out[b].torso_ang_vel_d_body[i] = g->joy_ang_rate[i]; /* :93-95 */
    s->torso_pos_d_world[0] += v_world[0] * 5.0 / 1000.0; /* :98-100 */
    s->torso_pos_d_world[1] += v_world[1] * 5.0 / 1000.0;
    s->torso_pos_d_world[2] = g->joy_body_height;
    double dp[3];
    for (int i = 0; i < 3; ++i) dp[i] = s->torso_pos_d_world[i] - g->torso_pos_world[i];
    for (int i = 0; i < 3; ++i) p_body[i] = R[i] * dp[0] + R[3 + i] * dp[1] + R[6 + i] * dp[2]; /* :102 */
    for (int i = 0; i < 3; ++i) out[b].torso_pos_d_body[i] = CalculateAverage(&s->torso_pos_d_body_filter[i], p_body[i]);
    memcpy(out[b].torso_quat, g->torso_quat, sizeof(double) * 4);
    memcpy(out[b].torso_lin_vel_world, g->torso_lin_vel_world, sizeof(double) * 3);
  }
  return QMPC_OK;
}

/* Raibert heuristic (BaseInterface.cpp:265-288) */
int qmpc_ref_raibert_targets(const QmpcRaibertParams* rp, const QmpcGoalInput* in, int batch, double* tgt_world,
                             double* tgt_rel) {
  if (!rp || !in || batch < 0) return QMPC_ERR_ARG;
  for (int b = 0; b < batch; ++b) {
    const QmpcGoalInput* g = &in[b];
    double R[9];
    quat_to_rot_ref(g->torso_quat, R);
    const double yaw = quat_yaw_ref(g->torso_quat);
    const double c = cos(yaw), s = sin(yaw);
    const double vrel0 = c * g->torso_lin_vel_world[0] + s * g->torso_lin_vel_world[1]; /* Rz^T v */
    const double vrel1 = -s * g->torso_lin_vel_world[0] + c * g->torso_lin_vel_world[1];
    const double k = sqrt(fabs(g->torso_pos_world[2]) / 9.81);
    double d[2];
    d[0] = k * (vrel0 - g->joy_vel[0]) + (1.0 / rp->gait_freq) / 2.0 * g->joy_vel[0];
    if (d[0] < -rp->delta_x_limit) d[0] = -rp->delta_x_limit;
    if (d[0] > rp->delta_x_limit) d[0] = rp->delta_x_limit;
    d[1] = k * (vrel1 - g->joy_vel[1]) + (1.0 / rp->gait_freq) / 2.0 * g->joy_vel[1];
    if (d[1] < -rp->delta_y_limit) d[1] = -rp->delta_y_limit;
    if (d[1] > rp->delta_y_limit) d[1] = rp->delta_y_limit;
    const double abs_d[2] = {c * d[0] - s * d[1], s * d[0] + c * d[1]};
    for (int i = 0; i < 4; ++i) {
      const double* p = rp->default_foot_pos_rel + 3 * i;
      double a[3] = {c * p[0] - s * p[1], s * p[0] + c * p[1], p[2]}; /* Rz * default_foot_pos_rel */
      a[0] += abs_d[0];
      a[1] += abs_d[1];
      if (tgt_rel)
        for (int r = 0; r < 3; ++r) tgt_rel[12 * (size_t)b + 3 * i + r] = R[r] * a[0] + R[3 + r] * a[1] + R[6 + r] * a[2];
      if (tgt_world)
        for (int r = 0; r < 3; ++r) tgt_world[12 * (size_t)b + 3 * i + r] = a[r] + g->torso_pos_world[r];
    }
  }
  return QMPC_OK;
}
