// CudaConvexMpc.cpp — see CudaConvexMpc.h.  Host-side packing only; no solver arithmetic here.
#include "CudaConvexMpc.h"

#include <cstdio>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>

namespace legged {

CudaConvexMpc::CudaConvexMpc(LeggedState& state, int device) {
  n = 12;
  m = 12;
  h = state.param.mpc_update_period;  // milliseconds, as in the reference (ConvexMpc.cpp:8)
  horizon = state.param.mpc_horizon;
  for (int i = 0; i < NUM_LEG; i++) leg_FSM[i].reset_params(state, i);
  num_contacts = 0;

  // solver configuration: ConvexMpc.cpp:36-38 (5 iterations, back-tracking line search) and the ROS
  // parameters the lambdas capture (mu, fz_max, weights, mass).  The model's own mass / inertia are the
  // constants QuadrupedModel::ct_srb_dynamics hard-codes (AltroUtils.cpp:239, 268-270) = the defaults.
  if (qmpc_default_config(QMPC_MODEL_EULER_CONVEX, horizon, &cfg_) != QMPC_OK)
    throw std::runtime_error("CudaConvexMpc: unsupported mpc_horizon");
  cfg_.dt = h / 1000.0;
  for (int i = 0; i < 12; ++i) cfg_.q_weights[i] = state.param.q_weights[i];
  cfg_.q_weights[12] = 0.0;
  for (int i = 0; i < 12; ++i) cfg_.r_weights[i] = state.param.r_weights[i];
  cfg_.mu = state.param.mu;
  cfg_.fz_max = state.param.fz_max;
  cfg_.robot_mass = state.param.robot_mass;
  std::memset(&prob_, 0, sizeof(prob_));
  std::memset(&last_, 0, sizeof(last_));
  const int rc = qmpc_create(&cfg_, /*max_batch=*/1, device, &handle_);
  if (rc != QMPC_OK) {
    std::string msg = std::string("CudaConvexMpc: qmpc_create failed: ") + qmpc_last_error(handle_);
    qmpc_destroy(handle_);
    handle_ = nullptr;
    throw std::runtime_error(msg);
  }
}

CudaConvexMpc::~CudaConvexMpc() { qmpc_destroy(handle_); }

bool CudaConvexMpc::update(LeggedState& state) {
  goal_update(state);
  foot_update(state);
  grf_update(state);
  if (state.param.terrain_adpt_state == 1) terrain_update(state);
  return true;
}

bool CudaConvexMpc::goal_update(LeggedState& state) {
  if (!state.estimator_init) {
    std::cout << "Estimator is not initialized!" << std::endl;
    return true;
  }
  state.ctrl.torso_pos_d_world[0] = state.joy.body_x;
  state.ctrl.torso_pos_d_world[1] = state.joy.body_y;
  state.ctrl.torso_pos_d_world[2] = state.joy.body_height;
  // forward velocity ramps towards the joystick value by 1 m/s^2; lateral velocity follows it directly
  const double step = 1.0 * h / 1000.0;
  if (state.ctrl.torso_lin_vel_d_rel[0] < state.joy.velx) state.ctrl.torso_lin_vel_d_rel[0] += step;
  else if (state.ctrl.torso_lin_vel_d_rel[0] > state.joy.velx) state.ctrl.torso_lin_vel_d_rel[0] -= step;
  state.ctrl.torso_lin_vel_d_rel[1] = state.joy.vely;
  state.ctrl.torso_lin_vel_d_rel[2] = 0.0;
  state.ctrl.torso_lin_vel_d_world = state.fbk.torso_rot_mat_z * state.ctrl.torso_lin_vel_d_rel;
  state.ctrl.torso_ang_vel_d_body[2] = state.joy.yaw_rate;
  return true;
}

bool CudaConvexMpc::foot_update(LeggedState& state) {
  if (state.ctrl.movement_mode == 0) {
    for (int i = 0; i < NUM_LEG; i++) {
      leg_FSM[i].reset();
      state.ctrl.plan_contacts[i] = true;
    }
  } else {
    for (int i = 0; i < NUM_LEG; i++)
      state.ctrl.gait_counter[i] =
          leg_FSM[i].update(h / 1000.0, state.param.gait_freq, state.fbk.foot_pos_world.col(i),
                            state.ctrl.foot_pos_target_world.col(i), state.fbk.foot_contact_flag[i]);
    for (int i = 0; i < NUM_LEG; i++) state.ctrl.plan_contacts[i] = leg_FSM[i].get_contact_state();
  }
  for (int leg = 0; leg < NUM_LEG; ++leg)
    for (int i = 0; i < 3; ++i) {
      state.ctrl.optimized_state[6 + 3 * leg + i] = leg_FSM[leg].FSM_foot_pos_target_world[i];
      state.ctrl.optimized_input[12 + 3 * leg + i] = leg_FSM[leg].FSM_foot_vel_target_world[i];
      state.ctrl.optimized_input[24 + 3 * leg + i] = leg_FSM[leg].FSM_foot_acc_target_world[i];
    }
  return true;
}

bool CudaConvexMpc::grf_update(LeggedState& state) {
  // ---- pack exactly the fields ConvexMpc::grf_update reads (ConvexMpc.cpp:91-117, 155-166, 191)
  QmpcConvexProblem& p = prob_;
  for (int i = 0; i < 3; ++i) {
    p.torso_euler[i] = state.fbk.torso_euler[i];
    p.torso_pos_world[i] = state.fbk.torso_pos_world[i];
    p.torso_ang_vel_world[i] = state.fbk.torso_ang_vel_world[i];
    p.torso_lin_vel_world[i] = state.fbk.torso_lin_vel_world[i];
    p.torso_pos_d_world[i] = state.ctrl.torso_pos_d_world[i];
    p.torso_lin_vel_d_world[i] = state.ctrl.torso_lin_vel_d_world[i];
    for (int j = 0; j < 3; ++j) p.torso_rot_mat[3 * i + j] = state.fbk.torso_rot_mat(i, j);
  }
  p.yaw_rate_d = state.ctrl.torso_ang_vel_d_body[2];
  num_contacts = 0;
  for (int leg = 0; leg < NUM_LEG; ++leg) {
    for (int i = 0; i < 3; ++i) p.foot_pos_abs_com[3 * leg + i] = state.fbk.foot_pos_abs_com(i, leg);
    p.plan_contacts[leg] = state.ctrl.plan_contacts[leg] ? 1 : 0;
    num_contacts += p.plan_contacts[leg];
  }
  // ---- solve on the GPU (batch = 1; H2D, kernel, D2H, sync inside the call)
  QmpcResult r;
  int rc;
  if (use_schedule_ && state.ctrl.movement_mode != 0) {
    std::memset(&sched_, 0, sizeof(sched_));
    for (int k = 0; k < horizon && k < QMPC_MAX_HORIZON; ++k)
      for (int leg = 0; leg < NUM_LEG; ++leg)
        if (leg_FSM[leg].predict_contact_state(k * h / 1000.0) == STANCE) sched_.mask[k] |= (uint8_t)(1u << leg);
    rc = qmpc_solve_batch_convex_sched_host(handle_, &p, &sched_, 1, &r);
  } else {
    rc = qmpc_solve_batch_convex_host(handle_, &p, 1, &r);
  }
  const bool bad = rc != QMPC_OK || r.status == QMPC_STATUS_NONFINITE || r.status == QMPC_STATUS_BACKWARD_FAILED;
  last_rc_ = rc;
  last_tick_failed_ = bad;
  if (bad) {   // update() keeps returning true (the reference's contract), but the failure is counted and reported
    ++failure_count_;
    if (rc != QMPC_OK) std::snprintf(last_error_, sizeof(last_error_), "qmpc rc=%d: %s", rc, qmpc_last_error(handle_));
    else std::snprintf(last_error_, sizeof(last_error_), "solver status %d (%s)", r.status, qmpc_status_string(r.status));
    if (failure_count_ == 1) std::fprintf(stderr, "[CudaConvexMpc] solve failed: %s (further failures are counted)\n", last_error_);
    if (zero_on_failure_) for (int i = 0; i < 12; ++i) state.ctrl.optimized_input[i] = 0.0;
    if (rc == QMPC_OK) last_ = r;
    return true;
  }
  last_ = r;
  // ---- unpack what ConvexMpc::grf_update writes (ConvexMpc.cpp:190-195)
  for (int i = 0; i < 12; ++i) state.ctrl.optimized_input[i] = r.grf_body[i];   // R^T u per leg
  for (int i = 0; i < 3; ++i) {
    state.ctrl.optimized_state[i] = state.ctrl.torso_pos_d_world[i];
    state.ctrl.optimized_state[3 + i] = state.ctrl.torso_euler_d[i];
  }
  return true;
}

}  // namespace legged
