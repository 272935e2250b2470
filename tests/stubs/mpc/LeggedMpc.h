// TEST STUB of the reference headers the shim compiles against (mpc/LeggedMpc.h, LeggedState.h,
// LeggedParams.h, utils/LeggedContactFSM.h).  Only what quaternion_mpc_b200/shim uses; a tiny
// fixed-size vector/matrix/quaternion stands in for Eigen.  The real build uses the reference's
// headers unchanged.
#pragma once
#include <array>
#include <cmath>

#define NUM_LEG 4

namespace stubeig {
template <int N>
struct Vec {
  std::array<double, N> v{};
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  Vec operator-(const Vec& o) const { Vec r; for (int i = 0; i < N; ++i) r.v[i] = v[i] - o.v[i]; return r; }
};
template <int R, int C>
struct Mat {
  std::array<double, R * C> a{};  // row-major
  double& operator()(int r, int c) { return a[r * C + c]; }
  double operator()(int r, int c) const { return a[r * C + c]; }
  Mat<C, R> transpose() const { Mat<C, R> t; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c); return t; }
  Vec<R> operator*(const Vec<C>& x) const { Vec<R> y; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) y.v[r] += (*this)(r, c) * x.v[c]; return y; }
  Vec<R> col(int c) const { Vec<R> y; for (int r = 0; r < R; ++r) y.v[r] = (*this)(r, c); return y; }
};
struct Quat {
  double w_ = 1, x_ = 0, y_ = 0, z_ = 0;
  double& w() { return w_; } double& x() { return x_; } double& y() { return y_; } double& z() { return z_; }
};
}  // namespace stubeig

namespace Eigen {
using Vector3d = stubeig::Vec<3>;
using Vector4d = stubeig::Vec<4>;
using Matrix3d = stubeig::Mat<3, 3>;
using Quaterniond = stubeig::Quat;
using VectorXd = stubeig::Vec<13>;
}  // namespace Eigen

namespace legged {
struct LeggedFeedback {
  Eigen::Vector3d torso_pos_world, torso_lin_vel_world, torso_lin_vel_body, torso_ang_vel_body;
  Eigen::Vector3d torso_euler, torso_ang_vel_world;   // ConvexMpc only
  Eigen::Quaterniond torso_quat;
  Eigen::Matrix3d torso_rot_mat, torso_rot_mat_z;
  stubeig::Mat<3, 4> foot_pos_body, foot_pos_world, foot_pos_abs_com;
  Eigen::Vector4d foot_contact_flag;
  double mpc_time = 0;
};
struct LeggedCtrl {
  Eigen::Vector4d gait_counter;
  Eigen::Vector3d torso_pos_d_world, torso_pos_d_body, torso_lin_vel_d_body, torso_lin_vel_d_rel,
      torso_lin_vel_d_world, torso_ang_vel_d_body, torso_euler_d;
  Eigen::Quaterniond torso_quat_d;
  stubeig::Mat<3, 4> foot_pos_target_world;
  bool plan_contacts[NUM_LEG] = {true, true, true, true};
  int movement_mode = 0;
  stubeig::Vec<18> optimized_state;
  stubeig::Vec<36> optimized_input;
  stubeig::Vec<12> mpc_grf_world;
};
struct LeggedJoyCmd {
  double velx = 0, vely = 0, roll_rate = 0, pitch_rate = 0, yaw_rate = 0, body_height = 0.3, body_x = 0, body_y = 0;
  bool sin_ang_vel = false;   // LeggedState.h:157
};
struct LeggedParam {
  double mpc_update_period = 10.0;
  int mpc_horizon = 10;
  stubeig::Vec<13> q_weights;
  stubeig::Vec<12> r_weights;
  double w = 50.0, mu = 0.7, fz_max = 100.0, robot_mass = 12.84, gait_freq = 2.2;
  Eigen::Matrix3d trunk_inertia;
  int terrain_adpt_state = 0;
};
struct LeggedState {
  LeggedFeedback fbk;
  LeggedCtrl ctrl;
  LeggedJoyCmd joy;
  LeggedParam param;
  bool estimator_init = true;
};
enum LeggedContactState { SWING, STANCE };   // utils/LeggedContactFSM.h:12-15
struct LeggedContactFSM {
  Eigen::Vector3d FSM_foot_pos_target_world, FSM_foot_vel_target_world, FSM_foot_acc_target_world;
  bool contact = true;
  int leg_id = 0;
  double gait_phase = 0.0, gait_freq = 2.2;
  // stub of LeggedContactFSM::predict_contact_state for the default trot pattern (LeggedContactFSM.cpp:87-108, 272-286)
  LeggedContactState predict_contact_state(double dt) {
    double ph = gait_phase + gait_freq * dt;
    while (ph > 1.0) ph -= 1.0;
    const bool first_half = ph <= 0.5;
    const bool diag = leg_id == 0 || leg_id == 3;
    return (first_half == diag) ? STANCE : SWING;
  }
  void reset_params(LeggedState&, int id) { leg_id = id; }
  void reset() { contact = true; }
  double update(double, double, Eigen::Vector3d, Eigen::Vector3d, bool) { return 0.0; }
  bool get_contact_state() const { return contact; }
};
class LeggedMpc {
 public:
  LeggedMpc() {}
  virtual ~LeggedMpc() {}
  virtual bool update(LeggedState&) { return true; }
  virtual bool goal_update(LeggedState&) { return true; }
  virtual bool grf_update(LeggedState&) { return true; }
  virtual bool foot_update(LeggedState&) { return true; }
  virtual bool terrain_update(LeggedState&) { return true; }

 protected:
  LeggedContactFSM leg_FSM[NUM_LEG];
  int n = 0, m = 0;
  double h = 0;
  int horizon = 0, num_contacts = 0;
};
}  // namespace legged
