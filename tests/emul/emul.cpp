// emul.cpp — TEST-ONLY host build of the solver bodies in quaternion_mpc_b200/csrc/*.cuh.
//
// The kernels' per-problem bodies are written as QMPC_HD (host+device) functions; this file
// compiles the very same source with g++ so the arithmetic can be checked against the oracle in
// a container without a GPU.  It is never linked into libqmpc_b200.so and the product never
// loads it: on the GPU box the parity tests go through the C-ABI and the real kernels.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../quaternion_mpc_b200/csrc/qmpc_dense.cuh"
#include "../../quaternion_mpc_b200/csrc/qmpc_periph.cuh"
#ifdef QMPC_EMUL_SRB
#include "../../quaternion_mpc_b200/csrc/qmpc_srb.cuh"
#include "../../quaternion_mpc_b200/csrc/qmpc_coop.cuh"
#include "../../quaternion_mpc_b200/csrc/qmpc_phased.cuh"
#endif

using namespace qmpc;

static SolverOpts make_opts(const QmpcConfig& cfg) {
  SolverOpts o;
  o.N = cfg.horizon; o.iterations_max = cfg.iterations_max; o.h = (float)cfg.dt;
  o.penalty_initial = cfg.penalty_initial; o.penalty_scaling = cfg.penalty_scaling; o.penalty_max = cfg.penalty_max;
  o.tol_cost_intermediate = cfg.tol_cost_intermediate; o.tol_primal_feasibility = cfg.tol_primal_feasibility;
  o.tol_stationarity = cfg.tol_stationarity;
  o.ls_c1 = 1e-4; o.ls_decrease = 0.5; o.ls_iters_max = 25;
  return o;
}

template <class M>
static int run_dense(const QmpcConfig& cfg, const void* in, const unsigned char* sched, QmpcWarmStart* warm, int batch, QmpcResult* out) {
  SolverOpts o = make_opts(cfg);
  size_t stride = batch;
  std::vector<double> ws(DenseLayout<M>::total(cfg.horizon) * stride);
  for (int i = 0; i < batch; ++i)
    dense_solve_one<M>(cfg, o, (const typename M::Problem*)in, sched, warm, out, ws.data(), i, stride);
  return 0;
}

// `sched`: null, or batch x QMPC_MAX_HORIZON contact-mask bytes (include/qmpc.h QmpcContactSchedule)
extern "C" int emul_solve_dense(const QmpcConfig* cfg, const void* in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                                QmpcResult* out) {
  switch (cfg->model) {
    case QMPC_MODEL_QUAT_4FOOT: return run_dense<QuatModel<4>>(*cfg, in, sched, warm, batch, out);
    case QMPC_MODEL_QUAT_2FOOT: return run_dense<QuatModel<2>>(*cfg, in, sched, warm, batch, out);
    default: return run_dense<ConvexModel>(*cfg, in, sched, warm, batch, out);
  }
}

#ifdef QMPC_EMUL_SRB
template <int NF>
static int run_srb(const QmpcConfig& cfg, const QmpcProblem* in, const unsigned char* sched, QmpcWarmStart* warm, int batch, QmpcResult* out) {
  SolverOpts o = make_opts(cfg);
  size_t stride = batch;
  std::vector<double> ws(SrbLayout<NF>::total(cfg.horizon) * stride);
  for (int i = 0; i < batch; ++i) srb_solve_one<NF>(cfg, o, in, sched, warm, out, ws.data(), i, stride);
  return 0;
}
extern "C" int emul_solve_srb(const QmpcConfig* cfg, const QmpcProblem* in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                              QmpcResult* out) {
  if (cfg->model == QMPC_MODEL_QUAT_4FOOT) return run_srb<4>(*cfg, in, sched, warm, batch, out);
  if (cfg->model == QMPC_MODEL_QUAT_2FOOT) return run_srb<2>(*cfg, in, sched, warm, batch, out);
  return -1;
}

template <class M, int G>
static int run_coop(const QmpcConfig& cfg, const typename M::Problem* in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                    QmpcResult* out) {
  SolverOpts o = make_opts(cfg);
  using L = CoopLayout<M, G>;
  const int wide = cfg.horizon <= 10 ? 3 : (cfg.horizon <= 16 ? 2 : 0);
  std::vector<double> sm(L::smem_doubles(cfg.horizon, wide)), gs(L::scratch_doubles(cfg.horizon));
  double wts[kCoopBlockShared];
  for (int i = 0; i < kCoopBlockShared; ++i) wts[i] = coop_block_const(cfg, o.h, i);
  for (int i = 0; i < batch; ++i) coop_solve_one<M, G>(cfg, o, in, sched, warm, out, i, sm.data(), gs.data(), 0, 0u, wide, wts);
  return 0;
}
// `in`: QmpcProblem (QUAT models) or QmpcConvexProblem (EULER_CONVEX) records
extern "C" int emul_solve_coop(const QmpcConfig* cfg, const void* in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                               QmpcResult* out) {
  if (cfg->model == QMPC_MODEL_QUAT_4FOOT) return run_coop<QuatModel<4>, 16>(*cfg, (const QmpcProblem*)in, sched, warm, batch, out);
  if (cfg->model == QMPC_MODEL_QUAT_2FOOT) return run_coop<QuatModel<2>, 16>(*cfg, (const QmpcProblem*)in, sched, warm, batch, out);
  return run_coop<ConvexModel, 16>(*cfg, (const QmpcConvexProblem*)in, sched, nullptr, batch, out);
}

// the phased path on the host: the same phase functions, one phase per "launch", the solver state stored to and
// reloaded from the per-problem block between them exactly as the split kernels do (qmpc_phased.cuh)
template <class M, int G>
static int run_phased(const QmpcConfig& cfg, const typename M::Problem* in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                      QmpcResult* out) {
  SolverOpts o = make_opts(cfg);
  using L = CoopLayout<M, G>;
  const int N = cfg.horizon;
  const size_t pstride = L::problem_doubles(N);
  std::vector<double> ws(pstride * batch), trial(L::trial_doubles(N));
  std::vector<double> smB(L::smem_doubles(N, 1)), smF(L::fwd_smem_doubles(N));
  double wts[kCoopBlockShared];
  for (int i = 0; i < kCoopBlockShared; ++i) wts[i] = coop_block_const(cfg, o.h, i);
  for (int i = 0; i < batch; ++i) phased_setup_one<M, G>(cfg, o, in, sched, warm, out, i, ws.data() + pstride * i, wts);
  for (int it = 0; it < o.iterations_max; ++it) {
    for (int i = 0; i < batch; ++i)
      phased_backward_one<M, G>(cfg, o, it, warm, out, i, smB.data(), ws.data() + pstride * i, 0, 0u, 1, wts);
    for (int i = 0; i < batch; ++i)
      phased_forward_one<M, G>(cfg, o, it, warm, out, i, smF.data(), ws.data() + pstride * i, trial.data(), 0, 0u, wts);
  }
  return 0;
}
extern "C" int emul_solve_phased(const QmpcConfig* cfg, const void* in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                                 QmpcResult* out) {
  if (cfg->model == QMPC_MODEL_QUAT_4FOOT) return run_phased<QuatModel<4>, 16>(*cfg, (const QmpcProblem*)in, sched, warm, batch, out);
  if (cfg->model == QMPC_MODEL_QUAT_2FOOT) return run_phased<QuatModel<2>, 16>(*cfg, (const QmpcProblem*)in, sched, warm, batch, out);
  return run_phased<ConvexModel, 16>(*cfg, (const QmpcConvexProblem*)in, sched, nullptr, batch, out);
}
extern "C" int emul_coop_smem_bytes(int nf, int horizon) {
  return 8 * (nf == 4 ? CoopLayout<QuatModel<4>, 16>::smem_doubles(horizon, 0) : CoopLayout<QuatModel<2>, 16>::smem_doubles(horizon, 0));
}
#endif

// ---- rows N1 / N2: the streaming kernels' bodies on the host
extern "C" int emul_predict_schedule(const QmpcConfig* cfg, const QmpcGaitState* g, int batch, QmpcContactSchedule* out) {
  for (int i = 0; i < batch; ++i) predict_schedule_one(g[i], cfg->horizon, cfg->dt, out[i]);
  return 0;
}
extern "C" int emul_leg_kinematics(const QmpcLegParams* lp, const double* q, int batch, double* foot, double* jac) {
  for (int t = 0; t < 4 * batch; ++t) leg_fk_jac(q + 3 * t, lp->rho_fix[t & 3], lp->rho_opt[t & 3], foot + 3 * t, jac + 9 * t);
  return 0;
}

// ---- row N3: goal_update / Raibert bodies on the host (state: element-major doubles, stride = capacity)
extern "C" int emul_goal_state_doubles(void) { return kGoalFields; }
extern "C" int emul_goal_update(double* state, int capacity, const QmpcGoalInput* in, int batch, QmpcProblem* out) {
  for (int i = 0; i < batch; ++i) goal_update_one(GoalStateRef{state + i, (size_t)capacity}, in[i], out[i]);
  return 0;
}
extern "C" int emul_raibert(const QmpcRaibertParams* rp, const QmpcGoalInput* in, int batch, double* tw, double* tr) {
  for (int i = 0; i < batch; ++i) raibert_one(*rp, in[i], tw + 12 * i, tr + 12 * i);
  return 0;
}

// ---- row N3, gait-FSM half: the foot_update body on the host (state: element-major doubles, stride = 4 * capacity)
extern "C" int emul_fsm_state_doubles(void) { return kFsmFields; }
extern "C" int emul_leg_fsm_init(double* state, int capacity, const int32_t* gait, int batch) {
  for (int t = 0; t < 4 * batch; ++t) leg_fsm_init_one(FsmRef{state + t, (size_t)4 * capacity}, t & 3, gait ? gait[t >> 2] : QMPC_GAIT_TROT);
  return 0;
}
extern "C" int emul_foot_update(double* state, int capacity, const QmpcFootUpdateInput* in, double dt, double gait_freq, int batch,
                                QmpcFootUpdateOutput* out) {
  QuinticInv ci;
  if (!quintic_C_inverse((float)(0.5 / gait_freq), ci.m)) return -1;
  for (int t = 0; t < 4 * batch; ++t) {
    const int b = t >> 2, leg = t & 3;
    int contact;
    leg_fsm_tick_one(FsmRef{state + t, (size_t)4 * capacity}, ci, leg, in[b].movement_mode, dt, gait_freq, in[b].foot_pos_world + 3 * leg,
                     in[b].foot_pos_target_world + 3 * leg, in[b].foot_contact_flag[leg] != 0, out[b].foot_pos_target + 3 * leg,
                     out[b].foot_vel_target + 3 * leg, out[b].foot_acc_target + 3 * leg, &out[b].gait_counter[leg], &contact);
    out[b].plan_contacts[leg] = contact;
  }
  return 0;
}
