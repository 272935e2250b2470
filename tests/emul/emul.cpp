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
#ifdef QMPC_EMUL_SRB
#include "../../quaternion_mpc_b200/csrc/qmpc_srb.cuh"
#include "../../quaternion_mpc_b200/csrc/qmpc_coop.cuh"
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
static int run_dense(const QmpcConfig& cfg, const void* in, int batch, QmpcResult* out) {
  SolverOpts o = make_opts(cfg);
  size_t stride = batch;
  std::vector<double> ws(DenseLayout<M>::total(cfg.horizon) * stride);
  for (int i = 0; i < batch; ++i)
    dense_solve_one<M>(cfg, o, (const typename M::Problem*)in, out, ws.data(), i, stride);
  return 0;
}

extern "C" int emul_solve_dense(const QmpcConfig* cfg, const void* in, int batch, QmpcResult* out) {
  switch (cfg->model) {
    case QMPC_MODEL_QUAT_4FOOT: return run_dense<QuatModel<4>>(*cfg, in, batch, out);
    case QMPC_MODEL_QUAT_2FOOT: return run_dense<QuatModel<2>>(*cfg, in, batch, out);
    default: return run_dense<ConvexModel>(*cfg, in, batch, out);
  }
}

#ifdef QMPC_EMUL_SRB
template <int NF>
static int run_srb(const QmpcConfig& cfg, const QmpcProblem* in, int batch, QmpcResult* out) {
  SolverOpts o = make_opts(cfg);
  size_t stride = batch;
  std::vector<double> ws(SrbLayout<NF>::total(cfg.horizon) * stride);
  for (int i = 0; i < batch; ++i) srb_solve_one<NF>(cfg, o, in, out, ws.data(), i, stride);
  return 0;
}
extern "C" int emul_solve_srb(const QmpcConfig* cfg, const QmpcProblem* in, int batch, QmpcResult* out) {
  if (cfg->model == QMPC_MODEL_QUAT_4FOOT) return run_srb<4>(*cfg, in, batch, out);
  if (cfg->model == QMPC_MODEL_QUAT_2FOOT) return run_srb<2>(*cfg, in, batch, out);
  return -1;
}

template <int NF, int G>
static int run_coop(const QmpcConfig& cfg, const QmpcProblem* in, int batch, QmpcResult* out) {
  SolverOpts o = make_opts(cfg);
  using L = CoopLayout<NF, G>;
  const bool wide = cfg.horizon <= 10;
  std::vector<double> sm(L::smem_doubles(cfg.horizon, wide)), gs(L::scratch_doubles(cfg.horizon));
  for (int i = 0; i < batch; ++i) coop_solve_one<NF, G>(cfg, o, in, out, i, sm.data(), gs.data(), 0, 0u, wide);
  return 0;
}
extern "C" int emul_solve_coop(const QmpcConfig* cfg, const QmpcProblem* in, int batch, QmpcResult* out) {
  if (cfg->model == QMPC_MODEL_QUAT_4FOOT) return run_coop<4, 16>(*cfg, in, batch, out);
  if (cfg->model == QMPC_MODEL_QUAT_2FOOT) return run_coop<2, 16>(*cfg, in, batch, out);
  return -1;
}
extern "C" int emul_coop_smem_bytes(int nf, int horizon) {
  return 8 * (nf == 4 ? CoopLayout<4, 16>::smem_doubles(horizon, false) : CoopLayout<2, 16>::smem_doubles(horizon, false));
}
#endif
