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

/* ---------------------------------------------------------------- N3, gait-FSM half
 * LeggedContactFSM as an object with the reference's members and methods
 * (legged_ctrl/include/utils/LeggedContactFSM.h:66-107, src/utils/LeggedContactFSM.cpp:4-78, 208-270),
 * QuinticCurve::get_foot_swing_target (src/utils/Utils.cpp:236-293) and QuatMpc::foot_update
 * (src/mpc/QuatMpc.cpp:278-305).  Eigen's C.inverse() for a dynamic 6x6 is PartialPivLU solved against the
 * identity; restated as such. */
typedef struct RefLegFsm {
  int leg_id, s;
  double gait_phase, gait_freq;
  int gait_freeze, gait_freeze_counter;
  LegFsm pat; /* gait_state_pattern / gait_switch_time / gait_pattern_size */
  int gait_pattern_index, prev_gait_pattern_index;
  double cur_state_start_time, cur_state_end_time;
  int not_first_call;
  double swing_start_foot_pos_world[3], swing_end_foot_pos_world[3], swing_extend_foot_pos_world[3];
  double terrain_height;
  double FSM_foot_pos_target_world[3], FSM_foot_vel_target_world[3], FSM_foot_acc_target_world[3];
  int gait;
} RefLegFsm;
typedef struct QmpcRefFsmState { RefLegFsm leg[4]; } QmpcRefFsmState;
int qmpc_ref_leg_fsm_state_bytes(void) { return (int)sizeof(QmpcRefFsmState); }

static void fsm_set_pattern(RefLegFsm* f, int gait) {
  set_pattern(&f->pat, gait, f->leg_id);
  f->gait = gait;
  f->gait_pattern_index = 0;
  f->prev_gait_pattern_index = f->pat.gait_pattern_size - 1;
  f->cur_state_start_time = 0.0;
  f->cur_state_end_time = f->pat.gait_switch_time[f->gait_pattern_index];
}
static void fsm_reset_params(RefLegFsm* f, double gait_freq, int leg_id, int gait) { /* :4-9 (+ optional set_*_pattern) */
  memset(f, 0, sizeof(*f));
  f->leg_id = leg_id;
  f->gait_freq = gait_freq;
  f->s = STANCE;
  fsm_set_pattern(f, gait);
}
static void fsm_reset(RefLegFsm* f) { /* :10-31 */
  f->gait_phase = 0;
  f->gait_freeze = 0;
  f->gait_freeze_counter = 0;
  f->gait_pattern_index = 0;
  f->prev_gait_pattern_index = f->pat.gait_pattern_size - 1;
  f->cur_state_start_time = 0;
  f->cur_state_end_time = f->pat.gait_switch_time[f->gait_pattern_index];
  if (f->s == SWING) {
    memcpy(f->FSM_foot_pos_target_world, f->swing_end_foot_pos_world, sizeof(double) * 3);
    memset(f->FSM_foot_vel_target_world, 0, sizeof(double) * 3);
  }
  f->s = f->pat.gait_state_pattern[f->gait_pattern_index];
  f->not_first_call = 0;
}
static double fsm_percent_in_state(const RefLegFsm* f) { /* :261-270 */
  double percent = (f->gait_phase - f->cur_state_start_time) / (f->cur_state_end_time - f->cur_state_start_time);
  if (percent < 0.0) percent = 0.0;
  else if (percent > 1.0) percent = 1.0;
  return percent;
}
static void fsm_common_enter(RefLegFsm* f) { /* :208-223 */
  f->prev_gait_pattern_index = f->gait_pattern_index;
  f->gait_pattern_index = (f->gait_pattern_index + 1) % f->pat.gait_pattern_size;
  if (f->gait_pattern_index < f->prev_gait_pattern_index) f->gait_phase -= 1.0;
  f->cur_state_start_time = f->gait_phase;
  f->cur_state_end_time = f->pat.gait_switch_time[f->gait_pattern_index];
  f->gait_freeze = 0;
  f->gait_freeze_counter = 0;
}
/* Eigen PartialPivLU of the 6x6 C, then inverse = solve(Identity) */
static void quintic_C_inverse_ref(float T, double* Cinv) {
  double A[36] = {1, 0, 0, 0, 0, 0,
                  1, T, T * T, T * T * T, T * T * T * T, T * T * T * T * T,
                  0, 1, 0, 0, 0, 0,
                  0, 1, 2 * T, 3 * T * T, 4 * T * T * T, 5 * T * T * T * T,
                  1, T / 2, T * T / 4, T * T * T / 8, T * T * T * T / 16, T * T * T * T * T / 32,
                  0, 1, T, 3 * T * T / 4, 4 * T * T * T / 8, 5 * T * T * T * T / 16};
  int perm[6] = {0, 1, 2, 3, 4, 5};
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = fabs(A[6 * k + k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[6 * i + k]) > best) { best = fabs(A[6 * i + k]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < 6; ++j) { double t = A[6 * k + j]; A[6 * k + j] = A[6 * piv + j]; A[6 * piv + j] = t; }
      int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
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
}
/* QuinticCurve::get_foot_swing_target (Utils.cpp:236-293); out = [pos(3), vel(3), acc(3)] */
static void get_foot_swing_target(float t, float T, const double* foot_pos_start, const double* foot_pos_final, double* out) {
  double Cinv[36];
  quintic_C_inverse_ref(T, Cinv);
  double dx = foot_pos_final[0] - foot_pos_start[0];
  double dy = foot_pos_final[1] - foot_pos_start[1];
  double k = 1.26 / T;
  double v_xy_mid = k * sqrt(dx * dx + dy * dy);
  double theta = atan2(fabs(dy), fabs(dx));
  double v_x_mid = (dx >= 0 ? 1 : -1) * v_xy_mid * cos(theta);
  double v_y_mid = (dy >= 0 ? 1 : -1) * v_xy_mid * sin(theta);
  const double con[3][6] = {
      {foot_pos_start[0], foot_pos_final[0], 0.0, 0.0, (foot_pos_start[0] + foot_pos_final[0]) / 2, v_x_mid},
      {foot_pos_start[1], foot_pos_final[1], 0.0, 0.0, (foot_pos_start[1] + foot_pos_final[1]) / 2, v_y_mid},
      {foot_pos_start[2], foot_pos_final[2], 0.1, -0.1, 0.1, 0.0}};
  for (int ax = 0; ax < 3; ++ax) {
    double a[6];
    for (int i = 0; i < 6; ++i) {
      double s = 0;
      for (int j = 0; j < 6; ++j) s += Cinv[6 * i + j] * con[ax][j];
      a[i] = s;
    }
    out[ax] = a[0] + a[1] * t + a[2] * t * t + a[3] * t * t * t + a[4] * t * t * t * t + a[5] * t * t * t * t * t;
    out[3 + ax] = a[1] + 2 * a[2] * t + 3 * a[3] * t * t + 4 * a[4] * t * t * t + 5 * a[5] * t * t * t * t;
    out[6 + ax] = 2 * a[2] + 6 * a[3] * t + 12 * a[4] * t * t + 20 * a[5] * t * t * t;
  }
}
/* LeggedContactFSM::update (:33-78) */
static double fsm_update(RefLegFsm* f, double dt, double gait_freq, const double* foot_pos_cur_world,
                         const double* foot_pos_target_world, int foot_force_flag) {
  if (!f->not_first_call) {
    memcpy(f->swing_start_foot_pos_world, foot_pos_cur_world, sizeof(double) * 3);
    memcpy(f->swing_end_foot_pos_world, foot_pos_target_world, sizeof(double) * 3);
    memcpy(f->FSM_foot_pos_target_world, foot_pos_target_world, sizeof(double) * 3);
    memset(f->FSM_foot_vel_target_world, 0, sizeof(double) * 3);
    f->not_first_call = 1;
  }
  f->gait_phase += gait_freq * dt;
  if (f->s == STANCE) {
    if (f->gait_phase >= f->cur_state_end_time) {
      f->terrain_height = foot_pos_cur_world[2]; /* stance_exit */
      fsm_common_enter(f);                       /* swing_enter */
      memcpy(f->swing_start_foot_pos_world, foot_pos_cur_world, sizeof(double) * 3);
      memset(f->swing_extend_foot_pos_world, 0, sizeof(double) * 3);
      f->s = SWING;
    }
  } else if (f->s == SWING) {
    if (fsm_percent_in_state(f) > 0.9 && foot_force_flag) {
      f->s = STANCE;
      fsm_common_enter(f); /* stance_enter */
      memcpy(f->FSM_foot_pos_target_world, foot_pos_cur_world, sizeof(double) * 3);
      memset(f->FSM_foot_vel_target_world, 0, sizeof(double) * 3);
    } else if (fsm_percent_in_state(f) >= 1.0) {
      f->s = STANCE;
      fsm_common_enter(f);
      memcpy(f->FSM_foot_pos_target_world, foot_pos_cur_world, sizeof(double) * 3);
      memset(f->FSM_foot_vel_target_world, 0, sizeof(double) * 3);
    }
  }
  if (f->s == SWING) { /* swing_update :237-246 ; stance_update is commented out in the reference */
    double t = fsm_percent_in_state(f);
    double fin[3], out[9];
    for (int a = 0; a < 3; ++a) fin[a] = foot_pos_target_world[a] + f->swing_extend_foot_pos_world[a];
    get_foot_swing_target((float)(0.5 * t / gait_freq), (float)(0.5 / gait_freq), f->swing_start_foot_pos_world, fin, out);
    memcpy(f->FSM_foot_pos_target_world, out, sizeof(double) * 3);
    memcpy(f->FSM_foot_vel_target_world, out + 3, sizeof(double) * 3);
    memcpy(f->FSM_foot_acc_target_world, out + 6, sizeof(double) * 3);
  }
  return f->gait_phase;
}

int qmpc_ref_leg_fsm_init(QmpcRefFsmState* st, const int32_t* gait, double gait_freq, int batch) {
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < 4; ++i) {
      fsm_reset_params(&st[b].leg[i], gait_freq, i, gait ? gait[b] : QMPC_GAIT_TROT);
      fsm_reset(&st[b].leg[i]);
    }
  return 0;
}
/* QuatMpc::foot_update (QuatMpc.cpp:278-305) + the FSM passthrough of grf_update (:270-272) */
int qmpc_ref_foot_update(QmpcRefFsmState* st, const QmpcFootUpdateInput* in, double dt, double gait_freq, int batch,
                         QmpcFootUpdateOutput* out) {
  for (int b = 0; b < batch; ++b) {
    if (in[b].movement_mode == 0) {
      for (int i = 0; i < 4; ++i) {
        fsm_reset(&st[b].leg[i]);
        out[b].plan_contacts[i] = 1;
        out[b].gait_counter[i] = st[b].leg[i].gait_phase;
      }
    } else {
      for (int i = 0; i < 4; ++i)
        out[b].gait_counter[i] = fsm_update(&st[b].leg[i], dt, gait_freq, in[b].foot_pos_world + 3 * i,
                                            in[b].foot_pos_target_world + 3 * i, in[b].foot_contact_flag[i]);
      for (int i = 0; i < 4; ++i) out[b].plan_contacts[i] = st[b].leg[i].s;
    }
    for (int i = 0; i < 4; ++i) {
      memcpy(out[b].foot_pos_target + 3 * i, st[b].leg[i].FSM_foot_pos_target_world, sizeof(double) * 3);
      memcpy(out[b].foot_vel_target + 3 * i, st[b].leg[i].FSM_foot_vel_target_world, sizeof(double) * 3);
      memcpy(out[b].foot_acc_target + 3 * i, st[b].leg[i].FSM_foot_acc_target_world, sizeof(double) * 3);
    }
  }
  return 0;
}
