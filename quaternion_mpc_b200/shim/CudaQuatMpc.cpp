// CudaQuatMpc.cpp — see CudaQuatMpc.h.  Host-side packing only; no solver arithmetic here.
#include "CudaQuatMpc.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

namespace legged {

double CudaQuatMpc::Avg100::push(double v) {
  auto add = [this](double x) {
    const double ns = sum + x;
    corr += (std::fabs(sum) >= std::fabs(x)) ? (sum - ns) + x : (x - ns) + sum;
    sum = ns;
  };
  if (q.size() >= 100) {
    add(-q.front());
    q.pop_front();
  }
  add(v);
  q.push_back(v);
  return (sum + corr) / 100.0;
}

CudaQuatMpc::CudaQuatMpc(LeggedState& state, int device) {
  // desired position starts at the measured one (QuatMpc.cpp:13-19)
  state.ctrl.torso_pos_d_world = state.fbk.torso_pos_world;
  const double nrm = std::sqrt(state.ctrl.torso_pos_d_world[0] * state.ctrl.torso_pos_d_world[0] +
                               state.ctrl.torso_pos_d_world[1] * state.ctrl.torso_pos_d_world[1] +
                               state.ctrl.torso_pos_d_world[2] * state.ctrl.torso_pos_d_world[2]);
  pos_d_world_init_ = !(nrm < 0.001);

  n = 13;
  m = 12;
  h = state.param.mpc_update_period;  // milliseconds, as in the reference
  horizon = state.param.mpc_horizon;
  for (int i = 0; i < NUM_LEG; i++) leg_FSM[i].reset_params(state, i);

  // solver + robot configuration from the ROS parameters (QuatMpc.cpp:21-38, 182, 227)
  if (qmpc_default_config(QMPC_MODEL_QUAT_4FOOT, horizon, &cfg_) != QMPC_OK)
    throw std::runtime_error("CudaQuatMpc: unsupported mpc_horizon");
  cfg_.dt = h / 1000.0;
  for (int i = 0; i < 13; ++i) cfg_.q_weights[i] = state.param.q_weights[i];
  for (int i = 0; i < 12; ++i) cfg_.r_weights[i] = state.param.r_weights[i];
  cfg_.w = state.param.w;
  cfg_.mu = state.param.mu;
  cfg_.fz_max = state.param.fz_max;
  cfg_.robot_mass = state.param.robot_mass;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cfg_.inertia[3 * r + c] = 1.2 * state.param.trunk_inertia(r, c);
  std::memset(&prob_, 0, sizeof(prob_));
  std::memset(&last_, 0, sizeof(last_));
  const int rc = qmpc_create(&cfg_, /*max_batch=*/1, device, &handle_);
  if (rc != QMPC_OK) {
    std::string msg = std::string("CudaQuatMpc: qmpc_create failed: ") + qmpc_last_error(handle_);
    qmpc_destroy(handle_);
    handle_ = nullptr;
    throw std::runtime_error(msg);
  }
}

CudaQuatMpc::~CudaQuatMpc() { qmpc_destroy(handle_); }

bool CudaQuatMpc::update(LeggedState& state) {
  goal_update(state);
  foot_update(state);
  grf_update(state);
  return true;
}

bool CudaQuatMpc::goal_update(LeggedState& state) {
  if (!state.estimator_init) return true;
  if (!pos_d_world_init_) {
    state.ctrl.torso_pos_d_world = state.fbk.torso_pos_world;
    pos_d_world_init_ = true;
  }
  // joystick -> desired velocity in the yaw frame, world frame, body frame; 100-tick average
  state.ctrl.torso_lin_vel_d_rel[0] = state.joy.velx;
  state.ctrl.torso_lin_vel_d_rel[1] = state.joy.vely;
  state.ctrl.torso_lin_vel_d_rel[2] = 0.0;
  state.ctrl.torso_lin_vel_d_world = state.fbk.torso_rot_mat_z * state.ctrl.torso_lin_vel_d_rel;
  state.ctrl.torso_lin_vel_d_body = state.fbk.torso_rot_mat.transpose() * state.ctrl.torso_lin_vel_d_world;
  for (int i = 0; i < 3; ++i) vel_d_body_filt_[i] = vel_filt_[i].push(state.ctrl.torso_lin_vel_d_body[i]);

  state.ctrl.torso_ang_vel_d_body[0] = state.joy.roll_rate;
  state.ctrl.torso_ang_vel_d_body[1] = state.joy.pitch_rate;
  state.ctrl.torso_ang_vel_d_body[2] = state.joy.yaw_rate;

  // desired position integrates the desired velocity over the fixed 5 ms tick
  state.ctrl.torso_pos_d_world[0] += state.ctrl.torso_lin_vel_d_world[0] * 5.0 / 1000.0;
  state.ctrl.torso_pos_d_world[1] += state.ctrl.torso_lin_vel_d_world[1] * 5.0 / 1000.0;
  state.ctrl.torso_pos_d_world[2] = state.joy.body_height;
  state.ctrl.torso_pos_d_body =
      state.fbk.torso_rot_mat.transpose() * (state.ctrl.torso_pos_d_world - state.fbk.torso_pos_world);
  for (int i = 0; i < 3; ++i) pos_d_body_filt_[i] = pos_filt_[i].push(state.ctrl.torso_pos_d_body[i]);
  return true;
}

bool CudaQuatMpc::foot_update(LeggedState& state) {
  if (state.ctrl.movement_mode == 0) {
    for (int i = 0; i < NUM_LEG; i++) {
      leg_FSM[i].reset();
      state.ctrl.plan_contacts[i] = true;
    }
  } else {
    for (int i = 0; i < NUM_LEG; i++)
      state.ctrl.gait_counter[i] =
          leg_FSM[i].update(5.0 / 1000.0, state.param.gait_freq, state.fbk.foot_pos_world.col(i),
                            state.ctrl.foot_pos_target_world.col(i), state.fbk.foot_contact_flag[i]);
    for (int i = 0; i < NUM_LEG; i++) state.ctrl.plan_contacts[i] = leg_FSM[i].get_contact_state();
  }
  return true;
}

// A failed solve (see FailurePolicy in the header): count it, keep the error text, log the first one, write the
// GRF outputs the policy asks for, and keep the desired attitude integrating (QuatMpc.cpp:128-137 runs before
// the solve in the reference, so it must not stall with it).
void CudaQuatMpc::on_failure(LeggedState& state, int rc, int status) {
  ++failure_count_;
  last_rc_ = rc;
  last_tick_failed_ = true;
  if (rc != QMPC_OK) std::snprintf(last_error_, sizeof(last_error_), "qmpc rc=%d: %s", rc, qmpc_last_error(handle_));
  else std::snprintf(last_error_, sizeof(last_error_), "solver status %d (%s)", status, qmpc_status_string(status));
  if (failure_count_ == 1) std::fprintf(stderr, "[CudaQuatMpc] solve failed: %s (policy %d; further failures are counted)\n",
                                        last_error_, (int)failure_policy_);
  if (failure_policy_ != FailurePolicy::kHoldLast) {
    int nc = 0;
    for (int leg = 0; leg < NUM_LEG; ++leg) nc += state.ctrl.plan_contacts[leg] ? 1 : 0;
    for (int leg = 0; leg < NUM_LEG; ++leg)
      for (int i = 0; i < 3; ++i) {
        double f = 0.0;
        if (failure_policy_ == FailurePolicy::kWeightShare && i == 2 && nc > 0 && state.ctrl.plan_contacts[leg])
          f = state.param.robot_mass * 9.81 / nc;
        state.ctrl.optimized_input[3 * leg + i] = f;     // body frame; the weight share is along body z as in u_ref
      }
    for (int leg = 0; leg < NUM_LEG; ++leg) {             // mpc_grf_world = R0 u (QuatMpc.cpp:268)
      const double fz = state.ctrl.optimized_input[3 * leg + 2];
      for (int i = 0; i < 3; ++i) state.ctrl.mpc_grf_world[3 * leg + i] = state.fbk.torso_rot_mat(i, 2) * fz;
    }
  }
  if (rc != QMPC_OK) {   // no result came back: integrate the desired attitude here (QuatMpc.cpp:128-137)
    const double q[4] = {prob_.torso_quat_d[0], prob_.torso_quat_d[1], prob_.torso_quat_d[2], prob_.torso_quat_d[3]};
    const double* w = prob_.torso_ang_vel_d_body;
    const double sc = 0.5 * cfg_.quat_d_dt;
    double qd[4] = {q[0] + sc * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2]), q[1] + sc * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2]),
                    q[2] + sc * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2]), q[3] + sc * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2])};
    const double nrm = std::sqrt(qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3]);
    if (nrm > 0 && std::isfinite(nrm)) {
      state.ctrl.torso_quat_d.w() = qd[0] / nrm; state.ctrl.torso_quat_d.x() = qd[1] / nrm;
      state.ctrl.torso_quat_d.y() = qd[2] / nrm; state.ctrl.torso_quat_d.z() = qd[3] / nrm;
    }
  }
}

bool CudaQuatMpc::grf_update(LeggedState& state) {
  const auto t0 = std::chrono::high_resolution_clock::now();
  // ---- pack exactly the fields QuatMpc::grf_update reads (QuatMpc.cpp:118-246)
  QmpcProblem& p = prob_;
  p.torso_quat[0] = state.fbk.torso_quat.w();
  p.torso_quat[1] = state.fbk.torso_quat.x();
  p.torso_quat[2] = state.fbk.torso_quat.y();
  p.torso_quat[3] = state.fbk.torso_quat.z();
  p.torso_quat_d[0] = state.ctrl.torso_quat_d.w();
  p.torso_quat_d[1] = state.ctrl.torso_quat_d.x();
  p.torso_quat_d[2] = state.ctrl.torso_quat_d.y();
  p.torso_quat_d[3] = state.ctrl.torso_quat_d.z();
  for (int i = 0; i < 3; ++i) {
    p.torso_lin_vel_world[i] = state.fbk.torso_lin_vel_world[i];
    p.torso_ang_vel_body[i] = state.fbk.torso_ang_vel_body[i];
    p.torso_pos_d_body[i] = pos_d_body_filt_[i];
    p.torso_lin_vel_d_body[i] = vel_d_body_filt_[i];
    p.torso_ang_vel_d_body[i] = state.ctrl.torso_ang_vel_d_body[i];
  }
  for (int leg = 0; leg < NUM_LEG; ++leg) {
    for (int i = 0; i < 3; ++i) p.foot_pos_body[3 * leg + i] = state.fbk.foot_pos_body(i, leg);
    p.plan_contacts[leg] = state.ctrl.plan_contacts[leg] ? 1 : 0;
  }
  // "sin ang vel test" (QuatMpc.cpp:139-146): after the integration step the reference REPLACES torso_quat_d by
  // euler_to_quat of a sine Euler trajectory (Utils.cpp:76-99).  Reproduced by handing the solve that quaternion
  // with a zero desired rate: its own integration + renormalisation (QuatMpc.cpp:132-133) is then the identity.
  if (state.joy.sin_ang_vel) {
    const double e = 3.14 / 8 * std::sin(2 * 3.14 / 900 * attitude_traj_count_);
    state.ctrl.torso_euler_d[0] = e; state.ctrl.torso_euler_d[1] = e; state.ctrl.torso_euler_d[2] = e;
    attitude_traj_count_ += 1;
    const double cr = std::cos(e / 2), sr = std::sin(e / 2);   // roll = pitch = yaw = e
    p.torso_quat_d[0] = cr * cr * cr + sr * sr * sr;
    p.torso_quat_d[1] = cr * cr * sr - sr * sr * cr;
    p.torso_quat_d[2] = cr * sr * cr + sr * cr * sr;
    p.torso_quat_d[3] = sr * cr * cr - cr * sr * sr;
    for (int i = 0; i < 3; ++i) p.torso_ang_vel_d_body[i] = 0.0;
  }
  // ---- solve on the GPU (batch = 1; H2D, kernel, D2H, sync inside the call)
  QmpcResult r;
  int rc;
  if (use_schedule_ && state.ctrl.movement_mode != 0) {
    std::memset(&sched_, 0, sizeof(sched_));
    for (int k = 0; k < horizon && k < QMPC_MAX_HORIZON; ++k)
      for (int leg = 0; leg < NUM_LEG; ++leg)
        if (leg_FSM[leg].predict_contact_state(k * h / 1000.0) == STANCE) sched_.mask[k] |= (uint8_t)(1u << leg);
    rc = qmpc_solve_batch_sched_host(handle_, &p, &sched_, 1, &r);
  } else {
    rc = qmpc_solve_batch_host(handle_, &p, 1, &r);
  }
  if (rc != QMPC_OK) {               // the reference's contract: update() returns true whatever happens (Main.cpp:107)
    on_failure(state, rc, -1);
    return true;
  }
  last_ = r;
  last_rc_ = QMPC_OK;
  last_tick_failed_ = false;
  if (r.status == QMPC_STATUS_NONFINITE || r.status == QMPC_STATUS_BACKWARD_FAILED) {
    on_failure(state, QMPC_OK, r.status);   // GRFs by policy; torso_quat_d below is still the solver's (valid) one
    // what is unpacked below: the policy's forces (hold-last: the state's current values, i.e. unchanged)
    for (int i = 0; i < 12; ++i) { r.grf_body[i] = state.ctrl.optimized_input[i]; r.grf_world[i] = state.ctrl.mpc_grf_world[i]; }
    if (!std::isfinite(r.torso_quat_d[0] + r.torso_quat_d[1] + r.torso_quat_d[2] + r.torso_quat_d[3])) {
      r.torso_quat_d[0] = p.torso_quat_d[0]; r.torso_quat_d[1] = p.torso_quat_d[1];
      r.torso_quat_d[2] = p.torso_quat_d[2]; r.torso_quat_d[3] = p.torso_quat_d[3];
    }
  }
  // ---- unpack what QuatMpc::grf_update writes (QuatMpc.cpp:133-137, 231, 261-272)
  state.ctrl.torso_quat_d.w() = r.torso_quat_d[0];
  state.ctrl.torso_quat_d.x() = r.torso_quat_d[1];
  state.ctrl.torso_quat_d.y() = r.torso_quat_d[2];
  state.ctrl.torso_quat_d.z() = r.torso_quat_d[3];
  state.fbk.torso_lin_vel_body = state.fbk.torso_rot_mat.transpose() * state.fbk.torso_lin_vel_world;
  for (int leg = 0; leg < NUM_LEG; ++leg) {
    for (int i = 0; i < 3; ++i) {
      state.ctrl.mpc_grf_world[3 * leg + i] = r.grf_world[3 * leg + i];
      state.ctrl.optimized_input[3 * leg + i] = r.grf_body[3 * leg + i];
      state.ctrl.optimized_state[6 + 3 * leg + i] = leg_FSM[leg].FSM_foot_pos_target_world[i];
      state.ctrl.optimized_input[12 + 3 * leg + i] = leg_FSM[leg].FSM_foot_vel_target_world[i];
      state.ctrl.optimized_input[24 + 3 * leg + i] = leg_FSM[leg].FSM_foot_acc_target_world[i];
    }
  }
  const auto t1 = std::chrono::high_resolution_clock::now();
  state.fbk.mpc_time = std::chrono::duration<double, std::milli>(t1 - t0).count();
  return true;
}

}  // namespace legged
