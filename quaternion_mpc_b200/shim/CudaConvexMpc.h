// CudaConvexMpc.h — drop-in replacement for legged::ConvexMpc (legged_ctrl/include/mpc/ConvexMpc.h) behind the
// same abstract boundary as CudaQuatMpc: legged::LeggedMpc (include/mpc/LeggedMpc.h:21-49).
//
// Main.cpp:89-93 would construct this instead of ConvexMpc when controller_type selects the convex MPC
// (INTEGRATION.md).  goal_update / foot_update stay on the host (joystick ramp and gait FSM,
// ConvexMpc.cpp:51-79, 200-222); grf_update — the Euler-angle SRB solve — goes through the C-ABI
// (qmpc_solve_batch_convex_host, include/qmpc.h) to the B200.  No CPU fallback: the constructor throws when
// the CUDA library cannot create a handle.
//
// Field access goes through operator[], operator()(i, j) only, which both Eigen (real build) and
// tests/stubs (stand-alone test build) provide.
#pragma once

#include "mpc/LeggedMpc.h"
#include "qmpc.h"

namespace legged {

class CudaConvexMpc : public LeggedMpc {
 public:
  explicit CudaConvexMpc(LeggedState& state, int device = 0);
  ~CudaConvexMpc();
  CudaConvexMpc(const CudaConvexMpc&) = delete;
  CudaConvexMpc& operator=(const CudaConvexMpc&) = delete;

  bool update(LeggedState& state) override;        // ConvexMpc.cpp:41-49
  bool goal_update(LeggedState& state) override;   // ConvexMpc.cpp:51-79
  bool grf_update(LeggedState& state) override;    // ConvexMpc.cpp:81-198  -> qmpc_solve_batch_convex_host
  bool foot_update(LeggedState& state) override;   // ConvexMpc.cpp:200-222
  bool terrain_update(LeggedState&) override { return true; }   // ConvexMpc.cpp:224-226

  // Optional (off by default = reference behaviour): per-knot contact masks from the inherited gait FSM's
  // predict_contact_state — the TODO the reference leaves at ConvexMpc.cpp:82.
  void enable_contact_schedule(bool on) { use_schedule_ = on; }
  const QmpcContactSchedule& last_schedule() const { return sched_; }
  LeggedContactFSM& leg_fsm(int leg) { return leg_FSM[leg]; }

  int last_status() const { return last_.status; }
  int last_iterations() const { return last_.iterations; }
  const QmpcConvexProblem& last_problem() const { return prob_; }

  // failed solves are never silent (see CudaQuatMpc.h): counted, last error kept, first one logged; on failure
  // the previous tick's forces are held unless zero_on_failure is set
  void set_zero_on_failure(bool z) { zero_on_failure_ = z; }
  long failure_count() const { return failure_count_; }
  int last_return_code() const { return last_rc_; }
  const char* last_error() const { return last_error_; }
  bool last_tick_failed() const { return last_tick_failed_; }

 private:
  bool zero_on_failure_ = false, last_tick_failed_ = false;
  long failure_count_ = 0;
  int last_rc_ = QMPC_OK;
  char last_error_[256] = {0};
  QmpcHandle* handle_ = nullptr;
  QmpcConfig cfg_;
  QmpcConvexProblem prob_;
  QmpcResult last_;
  bool use_schedule_ = false;
  QmpcContactSchedule sched_{};
};

}  // namespace legged
