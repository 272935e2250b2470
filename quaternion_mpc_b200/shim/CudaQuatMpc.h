// CudaQuatMpc.h — drop-in replacement for legged::QuatMpc behind the reference's own boundary,
// the abstract class legged::LeggedMpc (legged_ctrl/include/mpc/LeggedMpc.h:21-49).
//
// Main.cpp:94-95 would construct this instead of QuatMpc; the 200 Hz
// mpc_thread keeps calling mpc_ptr->update(state) (Main.cpp:107) unchanged.  goal_update and
// foot_update stay host-side (stateful joystick filters and gait FSM, QuatMpc.cpp:68-107, 278-305);
// only grf_update — the solve — goes through the C-ABI (include/qmpc.h) to the B200.
//
// The class only uses the LeggedState fields QuatMpc itself touches, through accessors that both
// Eigen (real build) and tests/stubs (stand-alone test build) provide: operator[], operator()(i,j),
// w()/x()/y()/z().
#pragma once

#include <deque>

#include "mpc/LeggedMpc.h"
#include "qmpc.h"

namespace legged {

class CudaQuatMpc : public LeggedMpc {
 public:
  // device: CUDA device ordinal.  Throws std::runtime_error if the CUDA library cannot create a
  // handle (there is no CPU fallback).
  explicit CudaQuatMpc(LeggedState& state, int device = 0);
  ~CudaQuatMpc();
  CudaQuatMpc(const CudaQuatMpc&) = delete;
  CudaQuatMpc& operator=(const CudaQuatMpc&) = delete;

  bool update(LeggedState& state) override;       // QuatMpc.cpp:57-66
  bool goal_update(LeggedState& state) override;  // QuatMpc.cpp:68-107
  bool grf_update(LeggedState& state) override;   // QuatMpc.cpp:109-276  -> qmpc_solve_batch_host
  bool foot_update(LeggedState& state) override;  // QuatMpc.cpp:278-305
  bool terrain_update(LeggedState&) override { return true; }

  // Optional (off by default = reference behaviour): plan with the per-knot contact schedule that
  // the reference's own, unused, LeggedContactFSM::predict_contact_state yields at t + k h
  // (LeggedContactFSM.cpp:272-286; TODO at ConvexMpc.cpp:82) instead of one mask for the horizon.
  void enable_contact_schedule(bool on) { use_schedule_ = on; }
  const QmpcContactSchedule& last_schedule() const { return sched_; }
  LeggedContactFSM& leg_fsm(int leg) { return leg_FSM[leg]; }   // the inherited per-leg gait FSM

  // last solve's status / iteration count (the reference discards ALTRO's SolveStatus)
  int last_status() const { return last_.status; }
  int last_iterations() const { return last_.iterations; }
  const QmpcProblem& last_problem() const { return prob_; }

  // What grf_update writes when the solve FAILED - the C-ABI call returned an error (CUDA fault, lost device)
  // or the solver reported a non-finite / non-positive-definite problem.  update() keeps returning true (the
  // reference's contract, Main.cpp:107 ignores the value), but the failure is never silent: it is counted,
  // the last error text is kept, the first occurrence is logged to stderr, and the GRF outputs follow an
  // explicit policy instead of silently staying at the previous tick's values.
  enum class FailurePolicy {
    kHoldLast,     // keep the previous tick's GRFs (what "outputs left untouched" amounts to), default
    kWeightShare,  // u_ref: the robot's weight shared by the feet planned in contact (QuatMpc.cpp:118-125)
    kZero          // zero forces
  };
  void set_failure_policy(FailurePolicy p) { failure_policy_ = p; }
  long failure_count() const { return failure_count_; }        // ticks whose solve failed, since construction
  int last_return_code() const { return last_rc_; }            // QMPC_OK or the error of the last tick
  const char* last_error() const { return last_error_; }       // sticky: text of the most recent failure
  bool last_tick_failed() const { return last_tick_failed_; }

 private:
  // 100-sample moving average with Neumaier-compensated running sum, same arithmetic as
  // utils/MovingWindowFilter.hpp:26-62 (kept local so the shim has no dependency on that header)
  struct Avg100 {
    std::deque<double> q;
    double sum = 0.0, corr = 0.0;
    double push(double v);
  };
  Avg100 vel_filt_[3], pos_filt_[3];
  double vel_d_body_filt_[3] = {0, 0, 0}, pos_d_body_filt_[3] = {0, 0, 0};
  bool pos_d_world_init_ = false;
  QmpcHandle* handle_ = nullptr;
  QmpcConfig cfg_;
  QmpcProblem prob_;
  QmpcResult last_;
  bool use_schedule_ = false;
  QmpcContactSchedule sched_{};
  double attitude_traj_count_ = 0;   // QuatMpc.h:33, the sine attitude test trajectory (QuatMpc.cpp:139-146)
  FailurePolicy failure_policy_ = FailurePolicy::kHoldLast;
  long failure_count_ = 0;
  int last_rc_ = QMPC_OK;
  bool last_tick_failed_ = false;
  char last_error_[256] = {0};
  void on_failure(LeggedState& state, int rc, int status);
};

}  // namespace legged
