// Test driver: builds a LeggedState from a QmpcProblem-like description, runs CudaQuatMpc::update
// through the LeggedMpc base pointer (as Main.cpp:107 does) and returns what it wrote.
#include <cstring>
#include <memory>
#include <stdexcept>

#include "CudaConvexMpc.h"
#include "CudaQuatMpc.h"

extern "C" int shim_construct_only(int device, char* err, int errlen) {
  legged::LeggedState st;
  try {
    legged::CudaQuatMpc mpc(st, device);
  } catch (const std::exception& e) {
    std::strncpy(err, e.what(), errlen - 1);
    err[errlen - 1] = 0;
    return 1;
  }
  return 0;
}

// in: QmpcProblem (desired quantities are fed through joystick-free state: see below); ticks >= 1
// out_problem: the QmpcProblem the shim packed on the last tick; out_grf_body/out_grf_world: 12 each
static int shim_run_impl(const QmpcProblem* in, int horizon, int ticks, QmpcProblem* out_problem,
                         double* out_grf_body, double* out_grf_world, int* status_iters, double gait_phase,
                         unsigned char* out_sched) {
  using namespace legged;
  LeggedState st;
  st.param.mpc_horizon = horizon;
  const double Q[13] = {2.5, 2.5, 10.0, 0, 0, 0, 0, 0.1, 0.1, 0.1, 0.15, 0.15, 0.15};
  for (int i = 0; i < 13; ++i) st.param.q_weights[i] = Q[i];
  for (int i = 0; i < 12; ++i) st.param.r_weights[i] = 1e-6;
  st.param.trunk_inertia(0, 0) = 0.0168128557;
  st.param.trunk_inertia(1, 1) = 0.063009565;
  st.param.trunk_inertia(2, 2) = 0.0716547275;
  st.fbk.torso_quat.w() = in->torso_quat[0]; st.fbk.torso_quat.x() = in->torso_quat[1];
  st.fbk.torso_quat.y() = in->torso_quat[2]; st.fbk.torso_quat.z() = in->torso_quat[3];
  {  // toRotationMatrix (BaseInterface.cpp:196)
    double w = in->torso_quat[0], x = in->torso_quat[1], y = in->torso_quat[2], z = in->torso_quat[3];
    auto& R = st.fbk.torso_rot_mat;
    R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - w * z); R(0, 2) = 2 * (x * z + w * y);
    R(1, 0) = 2 * (x * y + w * z); R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - w * x);
    R(2, 0) = 2 * (x * z - w * y); R(2, 1) = 2 * (y * z + w * x); R(2, 2) = 1 - 2 * (x * x + y * y);
    for (int i = 0; i < 3; ++i) st.fbk.torso_rot_mat_z(i, i) = 1.0;
  }
  st.fbk.torso_pos_world[2] = 0.3;
  for (int i = 0; i < 3; ++i) {
    st.fbk.torso_lin_vel_world[i] = in->torso_lin_vel_world[i];
    st.fbk.torso_ang_vel_body[i] = in->torso_ang_vel_body[i];
  }
  for (int leg = 0; leg < 4; ++leg)
    for (int i = 0; i < 3; ++i) st.fbk.foot_pos_body(i, leg) = in->foot_pos_body[3 * leg + i];
  st.ctrl.torso_quat_d.w() = in->torso_quat_d[0]; st.ctrl.torso_quat_d.x() = in->torso_quat_d[1];
  st.ctrl.torso_quat_d.y() = in->torso_quat_d[2]; st.ctrl.torso_quat_d.z() = in->torso_quat_d[3];
  st.joy.velx = 0.3; st.joy.vely = -0.05; st.joy.yaw_rate = 0.2; st.joy.body_height = 0.31;
  st.ctrl.movement_mode = out_sched ? 1 : 0;  // 0 = stand: foot_update sets all plan_contacts
  std::unique_ptr<LeggedMpc> mpc_ptr;
  try {
    mpc_ptr = std::make_unique<CudaQuatMpc>(st, 0);
  } catch (const std::exception&) {
    return 1;
  }
  auto* q = static_cast<CudaQuatMpc*>(mpc_ptr.get());
  if (out_sched) {   // walking: plan with the FSM's predicted per-knot contacts
    q->enable_contact_schedule(true);
    for (int leg = 0; leg < 4; ++leg) q->leg_fsm(leg).gait_phase = gait_phase;
  }
  for (int t = 0; t < ticks; ++t) mpc_ptr->update(st);
  if (out_sched) std::memcpy(out_sched, q->last_schedule().mask, QMPC_MAX_HORIZON);
  *out_problem = q->last_problem();
  for (int i = 0; i < 12; ++i) {
    out_grf_body[i] = st.ctrl.optimized_input[i];
    out_grf_world[i] = st.ctrl.mpc_grf_world[i];
  }
  status_iters[0] = q->last_status();
  status_iters[1] = q->last_iterations();
  return 0;
}

extern "C" int shim_run(const QmpcProblem* in, int horizon, int ticks, QmpcProblem* out_problem, double* out_grf_body,
                        double* out_grf_world, int* status_iters) {
  return shim_run_impl(in, horizon, ticks, out_problem, out_grf_body, out_grf_world, status_iters, 0.0, nullptr);
}
// same with enable_contact_schedule(true): also returns the schedule the shim built from leg_FSM[i].predict_contact_state
extern "C" int shim_run_sched(const QmpcProblem* in, int horizon, int ticks, double gait_phase, QmpcProblem* out_problem,
                              unsigned char* out_sched, double* out_grf_body, double* out_grf_world, int* status_iters) {
  return shim_run_impl(in, horizon, ticks, out_problem, out_grf_body, out_grf_world, status_iters, gait_phase, out_sched);
}

// ---- CudaConvexMpc: same idea.  `in` carries the measured state; desired quantities come from the joystick
// through the shim's own goal_update.  out_problem: what the shim packed on the last tick.
extern "C" int shim_convex_construct_only(int device, char* err, int errlen) {
  legged::LeggedState st;
  try {
    legged::CudaConvexMpc mpc(st, device);
  } catch (const std::exception& e) {
    std::strncpy(err, e.what(), errlen - 1);
    err[errlen - 1] = 0;
    return 1;
  }
  return 0;
}

extern "C" int shim_convex_run(const QmpcConvexProblem* in, int horizon, int ticks, int walking, double gait_phase,
                               QmpcConvexProblem* out_problem, unsigned char* out_sched, double* out_grf_body,
                               double* out_state6, int* status_iters) {
  using namespace legged;
  LeggedState st;
  st.param.mpc_horizon = horizon;
  st.param.mpc_update_period = 5.0;   // gazebo_go1_convex_mpc.yaml:36
  const double Q[12] = {3.0, 3.0, 3.0, 1.0, 1.0, 20.0, 0.0, 0.0, 3.0, 2.0, 3.0, 2.0};
  for (int i = 0; i < 12; ++i) st.param.q_weights[i] = Q[i];
  for (int i = 0; i < 12; ++i) st.param.r_weights[i] = 1e-6;
  st.param.mu = 0.6;
  st.param.fz_max = 200.0;
  for (int i = 0; i < 3; ++i) {
    st.fbk.torso_euler[i] = in->torso_euler[i];
    st.fbk.torso_pos_world[i] = in->torso_pos_world[i];
    st.fbk.torso_ang_vel_world[i] = in->torso_ang_vel_world[i];
    st.fbk.torso_lin_vel_world[i] = in->torso_lin_vel_world[i];
    for (int j = 0; j < 3; ++j) st.fbk.torso_rot_mat(i, j) = in->torso_rot_mat[3 * i + j];
    st.fbk.torso_rot_mat_z(i, i) = 1.0;
    st.ctrl.torso_euler_d[i] = 0.01 * (i + 1);
  }
  for (int leg = 0; leg < 4; ++leg)
    for (int i = 0; i < 3; ++i) st.fbk.foot_pos_abs_com(i, leg) = in->foot_pos_abs_com[3 * leg + i];
  st.joy.velx = 0.3; st.joy.vely = -0.05; st.joy.yaw_rate = 0.2; st.joy.body_height = 0.29;
  st.joy.body_x = in->torso_pos_world[0] + 0.01; st.joy.body_y = in->torso_pos_world[1] - 0.01;
  st.ctrl.movement_mode = walking ? 1 : 0;
  std::unique_ptr<LeggedMpc> mpc_ptr;
  try {
    mpc_ptr = std::make_unique<CudaConvexMpc>(st, 0);
  } catch (const std::exception&) {
    return 1;
  }
  auto* c = static_cast<CudaConvexMpc*>(mpc_ptr.get());
  if (walking) {
    c->enable_contact_schedule(true);
    for (int leg = 0; leg < 4; ++leg) c->leg_fsm(leg).gait_phase = gait_phase;
  }
  for (int t = 0; t < ticks; ++t) mpc_ptr->update(st);
  if (out_sched) std::memcpy(out_sched, c->last_schedule().mask, QMPC_MAX_HORIZON);
  *out_problem = c->last_problem();
  for (int i = 0; i < 12; ++i) out_grf_body[i] = st.ctrl.optimized_input[i];
  for (int i = 0; i < 6; ++i) out_state6[i] = st.ctrl.optimized_state[i];
  status_iters[0] = c->last_status();
  status_iters[1] = c->last_iterations();
  return 0;
}

// ---- failure policy and the sine attitude trajectory of CudaQuatMpc
static void fill_state(legged::LeggedState& st, const QmpcProblem* in) {
  const double Q[13] = {2.5, 2.5, 10.0, 0, 0, 0, 0, 0.1, 0.1, 0.1, 0.15, 0.15, 0.15};
  st.param.mpc_horizon = 10;
  for (int i = 0; i < 13; ++i) st.param.q_weights[i] = Q[i];
  for (int i = 0; i < 12; ++i) st.param.r_weights[i] = 1e-6;
  st.param.trunk_inertia(0, 0) = 0.0168128557; st.param.trunk_inertia(1, 1) = 0.063009565; st.param.trunk_inertia(2, 2) = 0.0716547275;
  st.fbk.torso_quat.w() = in->torso_quat[0]; st.fbk.torso_quat.x() = in->torso_quat[1];
  st.fbk.torso_quat.y() = in->torso_quat[2]; st.fbk.torso_quat.z() = in->torso_quat[3];
  double w = in->torso_quat[0], x = in->torso_quat[1], y = in->torso_quat[2], z = in->torso_quat[3];
  auto& R = st.fbk.torso_rot_mat;
  R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - w * z); R(0, 2) = 2 * (x * z + w * y);
  R(1, 0) = 2 * (x * y + w * z); R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - w * x);
  R(2, 0) = 2 * (x * z - w * y); R(2, 1) = 2 * (y * z + w * x); R(2, 2) = 1 - 2 * (x * x + y * y);
  for (int i = 0; i < 3; ++i) st.fbk.torso_rot_mat_z(i, i) = 1.0;
  st.fbk.torso_pos_world[2] = 0.3;
  for (int i = 0; i < 3; ++i) st.fbk.torso_lin_vel_world[i] = in->torso_lin_vel_world[i];
  for (int leg = 0; leg < 4; ++leg)
    for (int i = 0; i < 3; ++i) st.fbk.foot_pos_body(i, leg) = in->foot_pos_body[3 * leg + i];
  st.ctrl.torso_quat_d.w() = 1.0;
  st.joy.body_height = 0.3;
  st.ctrl.movement_mode = 0;
}

// tick 1: a normal solve; tick 2: a NaN velocity -> the solver reports NONFINITE.  policy: 0 hold-last, 1 weight share, 2 zero.
// out: grf_body after tick 1 (12) and after tick 2 (12); info = {failure_count, last_tick_failed, last_status, update() return}
extern "C" int shim_failure_policy(const QmpcProblem* in, int policy, double* grf_tick1, double* grf_tick2, int* info, char* err,
                                   int errlen) {
  using namespace legged;
  LeggedState st;
  fill_state(st, in);
  std::unique_ptr<LeggedMpc> mpc_ptr;
  try { mpc_ptr = std::make_unique<CudaQuatMpc>(st, 0); } catch (const std::exception&) { return 1; }
  auto* q = static_cast<CudaQuatMpc*>(mpc_ptr.get());
  q->set_failure_policy(policy == 0 ? CudaQuatMpc::FailurePolicy::kHoldLast
                                    : (policy == 1 ? CudaQuatMpc::FailurePolicy::kWeightShare : CudaQuatMpc::FailurePolicy::kZero));
  bool ok = mpc_ptr->update(st);
  for (int i = 0; i < 12; ++i) grf_tick1[i] = st.ctrl.optimized_input[i];
  if (q->failure_count() != 0 || q->last_tick_failed()) return 2;
  st.fbk.torso_lin_vel_world[1] = std::nan("");
  ok = mpc_ptr->update(st) && ok;
  for (int i = 0; i < 12; ++i) grf_tick2[i] = st.ctrl.optimized_input[i];
  info[0] = (int)q->failure_count(); info[1] = q->last_tick_failed(); info[2] = q->last_status(); info[3] = ok;
  std::strncpy(err, q->last_error(), errlen - 1);
  err[errlen - 1] = 0;
  return 0;
}

// joy.sin_ang_vel: `ticks` updates; returns the torso_quat_d the shim packed on the last tick and the one it wrote back
extern "C" int shim_sin_ang_vel(const QmpcProblem* in, int ticks, double* packed_quat_d, double* state_quat_d, double* euler_d) {
  using namespace legged;
  LeggedState st;
  fill_state(st, in);
  st.joy.sin_ang_vel = true;
  st.joy.yaw_rate = 0.4;   // must NOT move the desired attitude in this mode: the sine trajectory replaces it
  std::unique_ptr<LeggedMpc> mpc_ptr;
  try { mpc_ptr = std::make_unique<CudaQuatMpc>(st, 0); } catch (const std::exception&) { return 1; }
  auto* q = static_cast<CudaQuatMpc*>(mpc_ptr.get());
  for (int t = 0; t < ticks; ++t) mpc_ptr->update(st);
  for (int i = 0; i < 4; ++i) packed_quat_d[i] = q->last_problem().torso_quat_d[i];
  state_quat_d[0] = st.ctrl.torso_quat_d.w(); state_quat_d[1] = st.ctrl.torso_quat_d.x();
  state_quat_d[2] = st.ctrl.torso_quat_d.y(); state_quat_d[3] = st.ctrl.torso_quat_d.z();
  for (int i = 0; i < 3; ++i) euler_d[i] = st.ctrl.torso_euler_d[i];
  return 0;
}
