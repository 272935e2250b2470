// qmpc_phased.cuh — the cooperative solve split into one launch per phase (QMPC_KERNEL_PHASED).
//
// Why: the fused persistent kernel (qmpc_coop.cuh) carries ONE register budget (255, forced by the register
// Cholesky of the backward pass), ONE shared-memory footprint (13.5 KB per problem) and 95 KB of per-iteration
// SASS against a 32 KB instruction cache.  Here every AL-iLQR iteration is two launches
//     backward kernel : expansions, stationarity / dual update, Riccati recursion     (255 registers, full layout)
//     forward  kernel : speculative line search + accepted step                        (128 registers, 6 KB / problem)
// after one set-up launch (thread per problem: model assembly + nominal roll-out).  Every SM then runs one phase's
// code at a time (the phase alignment of the fused kernel taken to its end), the forward pass holds twice the
// resident warps, and the solver state that crosses a launch boundary (model, X, U, duals, gains, value functions,
// scalars: ~37 KB per problem) goes through L2 / HBM, which this path leaves idle (DRAM ~13 % busy fused).
// The phase bodies are the very functions the fused kernel calls, so the arithmetic is identical by construction.
#pragma once
#include "qmpc_coop.cuh"

namespace qmpc {

// solver scalars of one problem between launches (8 doubles at CoopLayout::pScal)
struct PhasedScal {
  double rho, phi, viol, cost_decrease, dphi0;
  int status, iters;
  double pad_[2];
};
static_assert(sizeof(PhasedScal) == 64, "PhasedScal is 8 doubles");

template <class M, int G>
QMPC_HD inline void phased_store_scal(const CoopCtx<M, G>& c, double* gp) {
  PhasedScal& s = *reinterpret_cast<PhasedScal*>(gp + CoopLayout<M, G>::pScal(c.N));
  s.rho = c.rho; s.phi = c.phi; s.viol = c.viol; s.cost_decrease = c.cost_decrease; s.dphi0 = c.dphi0;
  s.status = c.status; s.iters = c.iters;
}
template <class M, int G>
QMPC_HD inline void phased_load_scal(CoopCtx<M, G>& c, const double* gp) {
  const PhasedScal& s = *reinterpret_cast<const PhasedScal*>(gp + CoopLayout<M, G>::pScal(c.N));
  c.rho = s.rho; c.phi = s.phi; c.viol = s.viol; c.cost_decrease = s.cost_decrease; c.dphi0 = s.dphi0;
  c.status = s.status; c.iters = s.iters;
}

// ---- set-up: ONE THREAD per problem, everything in the problem's global block (model assembly and the nominal
//      roll-out are serial per problem: in the fused kernel 15 of a problem's 16 lanes idle through them)
template <class M, int G>
QMPC_HD void phased_setup_one(const QmpcConfig& cfg, const SolverOpts& o, const typename M::Problem* in,
                              const unsigned char* sched, QmpcWarmStart* warm, QmpcResult* out, int pid, double* gp,
                              const double* wts) {
  using L = CoopLayout<M, G>;
  constexpr int NU = M::NU, NC = M::NC;
  const int N = o.N;
  M& m = *reinterpret_cast<M*>(gp + L::pModel(N));
  double* X = gp + L::pX(N);
  double* U = gp + L::pU(N);
  double* gmu = gp + L::gmu(N);
  for (int i = 0; i < N * NC; ++i) gmu[i] = 0.0;
  {
    typename M::Problem prob = in[pid];
    m.setup(cfg, prob, sched ? sched + (size_t)pid * QMPC_MAX_HORIZON : nullptr, X);
  }
  CoopCtx<M, G> c;
  c.N = N;
  c.rho = o.penalty_initial;
  double J, vl;
  coop_rollout<M, false>(m, cfg, wts + 13, N, o.h, X, U, gp + L::gK(N), gmu, c.rho, 0.0, 0, &J, &vl, nullptr, nullptr, 0, G, nullptr, 0u,
                  (warm && warm[pid].valid) ? warm + pid : nullptr);
  c.phi = J; c.viol = vl;
  c.status = QMPC_STATUS_MAX_ITERATIONS; c.iters = 0;
  c.cost_decrease = INFINITY; c.dphi0 = 0;
  if (!isfinite(c.phi)) c.status = QMPC_STATUS_NONFINITE;
  phased_store_scal<M, G>(c, gp);
  if (c.status != QMPC_STATUS_MAX_ITERATIONS || o.iterations_max <= 0) {   // nothing more will run for this problem
    QmpcResult r;
    m.write_result(U, r);
    r.max_violation = c.viol;
    r.iterations = 0;
    r.status = c.status;
    out[pid] = r;
    if (warm) {
      warm[pid].valid = c.status != QMPC_STATUS_NONFINITE;
      for (int k = 0; k < N; ++k)
        for (int i = 0; i < 12; ++i) warm[pid].u[k][i] = i < NU ? U[k * NU + i] : 0.0;
    }
  }
}

// cooperative copy of the model, X and U between the problem block and shared memory
template <class M, int G>
QMPC_HD inline void phased_load_state(CoopCtx<M, G>& c, const double* gp, COOP_ARGS_DECL) {
  using L = CoopLayout<M, G>;
  (void)lane_id; (void)lane_mask;
  const int N = c.N;
  double* mdst = reinterpret_cast<double*>(c.m);
  COOP_PHASE {
#pragma unroll 1
    for (int e = lane; e < L::kModel; e += G) mdst[e] = gp[L::pModel(N) + e];
#pragma unroll 1
    for (int e = lane; e < (N + 1) * M::NX; e += G) c.X[e] = gp[L::pX(N) + e];
#pragma unroll 1
    for (int e = lane; e < N * M::NU; e += G) c.U[e] = gp[L::pU(N) + e];
  }
  COOP_SYNC();
}

// ---- backward launch of iteration `it`: pre + Riccati
template <class M, int G>
QMPC_HD void phased_backward_one(const QmpcConfig& cfg, const SolverOpts& o, int it, QmpcWarmStart* warm, QmpcResult* out, int pid,
                                 double* sm, double* gp, int lane_id, unsigned lane_mask, int flags, const double* wts) {
  CoopCtx<M, G> c;
  c.bind(sm, gp, nullptr, o.N, o.h, flags & 1, wts);   // duals stay in the problem block (they cross launches)
  phased_load_scal<M, G>(c, gp);
  if (c.status != QMPC_STATUS_MAX_ITERATIONS) return;
  phased_load_state<M, G>(c, gp, COOP_ARGS);
  coop_phase_pre<M, G>(c, cfg, o, it, COOP_ARGS);
  if (c.status == QMPC_STATUS_MAX_ITERATIONS) coop_phase_backward<M, G>(c, COOP_ARGS);
  COOP_PHASE {
    if (lane == 0) phased_store_scal<M, G>(c, gp);
  }
  if (c.status != QMPC_STATUS_MAX_ITERATIONS) coop_phase_epilogue<M, G>(c, out, warm, pid, c.P, COOP_ARGS);
}

// ---- forward launch of iteration `it`: line search + accepted step
template <class M, int G>
QMPC_HD void phased_forward_one(const QmpcConfig& cfg, const SolverOpts& o, int it, QmpcWarmStart* warm, QmpcResult* out, int pid,
                                double* sm, double* gp, double* trial, int lane_id, unsigned lane_mask, const double* wts) {
  using L = CoopLayout<M, G>;
  CoopCtx<M, G> c;
  c.bind_forward(sm, gp, trial, o.N, o.h, wts);
  phased_load_scal<M, G>(c, gp);
  if (c.status != QMPC_STATUS_MAX_ITERATIONS) return;
  phased_load_state<M, G>(c, gp, COOP_ARGS);
  coop_phase_forward<M, G>(c, cfg, o, it, COOP_ARGS);
  const int N = c.N;
  if (c.status == QMPC_STATUS_MAX_ITERATIONS) {   // accepted: the new trajectory goes back to the problem block
    COOP_PHASE {
#pragma unroll 1
      for (int e = lane; e < (N + 1) * M::NX; e += G) gp[L::pX(N) + e] = c.X[e];
#pragma unroll 1
      for (int e = lane; e < N * M::NU; e += G) gp[L::pU(N) + e] = c.U[e];
    }
  }
  COOP_PHASE {
    if (lane == 0) phased_store_scal<M, G>(c, gp);
  }
  if (c.status != QMPC_STATUS_MAX_ITERATIONS || it + 1 >= o.iterations_max)
    coop_phase_epilogue<M, G>(c, out, warm, pid, c.kstage, COOP_ARGS);
}

#ifdef __CUDACC__
template <class M, int G>
__global__ void __launch_bounds__(128)
qmpc_phased_setup_kernel(QmpcConfig cfg, SolverOpts o, const typename M::Problem* __restrict__ in,
                         const unsigned char* __restrict__ sched, QmpcWarmStart* __restrict__ warm,
                         QmpcResult* __restrict__ out, double* __restrict__ ws, int batch, size_t pstride) {
  const int pid = blockIdx.x * blockDim.x + threadIdx.x;
  if (pid >= batch) return;
  double wts[kCoopBlockShared];
#pragma unroll
  for (int i = 0; i < kCoopBlockShared; ++i) wts[i] = coop_block_const(cfg, o.h, i);
  phased_setup_one<M, G>(cfg, o, in, sched, warm, out, pid, ws + (size_t)pid * pstride, wts);
}

template <class M, int G>
__global__ void __launch_bounds__(QMPC_COOP_BLOCK, QMPC_COOP_MIN_BLOCKS)
qmpc_phased_backward_kernel(QmpcConfig cfg, SolverOpts o, int it, QmpcWarmStart* __restrict__ warm, QmpcResult* __restrict__ out,
                            double* __restrict__ ws, int batch, size_t pstride, int smem_per_problem, int flags) {
  extern __shared__ __align__(16) double smem_pool[];
  if (threadIdx.x < kCoopBlockShared) smem_pool[threadIdx.x] = coop_block_const(cfg, o.h, threadIdx.x);
  __syncthreads();
  const int groups_per_block = blockDim.x / G, group = threadIdx.x / G, lane_id = threadIdx.x % G;
  const unsigned lane_mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x % 32) / G * G));
  double* sm = smem_pool + kCoopBlockShared + (size_t)group * smem_per_problem;
  const int nslots = gridDim.x * groups_per_block;
  for (int pid = blockIdx.x * groups_per_block + group; pid < batch; pid += nslots)
    phased_backward_one<M, G>(cfg, o, it, warm, out, pid, sm, ws + (size_t)pid * pstride, lane_id, lane_mask, flags, smem_pool);
}

#ifndef QMPC_PHASED_FWD_MIN_BLOCKS
#define QMPC_PHASED_FWD_MIN_BLOCKS 4   // 16 warps per SM at 128 registers
#endif
template <class M, int G>
__global__ void __launch_bounds__(QMPC_COOP_BLOCK, QMPC_PHASED_FWD_MIN_BLOCKS)
qmpc_phased_forward_kernel(QmpcConfig cfg, SolverOpts o, int it, QmpcWarmStart* __restrict__ warm, QmpcResult* __restrict__ out,
                           double* __restrict__ ws, double* __restrict__ trial, int batch, size_t pstride, int smem_per_problem,
                           size_t trial_stride) {
  extern __shared__ __align__(16) double smem_pool[];
  if (threadIdx.x < kCoopBlockShared) smem_pool[threadIdx.x] = coop_block_const(cfg, o.h, threadIdx.x);
  __syncthreads();
  const int groups_per_block = blockDim.x / G, group = threadIdx.x / G, lane_id = threadIdx.x % G;
  const unsigned lane_mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x % 32) / G * G));
  double* sm = smem_pool + kCoopBlockShared + (size_t)group * smem_per_problem;
  const int slot = blockIdx.x * groups_per_block + group, nslots = gridDim.x * groups_per_block;
  double* tr = trial + (size_t)slot * trial_stride;
  for (int pid = slot; pid < batch; pid += nslots)
    phased_forward_one<M, G>(cfg, o, it, warm, out, pid, sm, ws + (size_t)pid * pstride, tr, lane_id, lane_mask, smem_pool);
}
#endif

}  // namespace qmpc
