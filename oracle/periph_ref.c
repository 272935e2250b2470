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
