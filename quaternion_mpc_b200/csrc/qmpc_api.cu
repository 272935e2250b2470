// qmpc_api.cu — the C-ABI of libqmpc_b200.so (declared in include/qmpc.h).
//
// Host side only: configuration defaults, handle life-cycle (device workspace allocated once),
// kernel selection and launch.  No solver arithmetic happens on the host and there is no CPU
// fallback: every solve entry point ends in a CUDA kernel launch or returns QMPC_ERR_CUDA.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>

#include "../../include/qmpc.h"
#include "qmpc_dense.cuh"
#include "qmpc_srb.cuh"
#include "qmpc_coop.cuh"
#include "qmpc_phased.cuh"
#include "qmpc_periph.cuh"

using namespace qmpc;

constexpr int kHostChunksMax = 4;   // a host batch of several problem waves is copied and solved in up to 4 chunks
struct QmpcHandle {
  QmpcConfig cfg;
  SolverOpts opts;
  int device;
  int max_batch;
  size_t stride;       // workspace problem stride (max_batch rounded up to 32)
  double* ws;          // device workspace
  size_t ws_bytes;
  void* d_in;          // staging for the *_host entry points
  void* h_stage;       // pinned host staging for small host batches from pageable memory (batch <= kStageBatch)
  int packed_launch;   // QmpcCreateOptions::packed_launch
  int host_chunks;     // QmpcCreateOptions::host_chunks
  QmpcContactSchedule* d_sched;
  QmpcResult* d_out;
  cudaStream_t stream; // stream used by the *_host entry points
  cudaStream_t copy_stream = nullptr;   // second stream of the chunked host pipeline (copies of the other chunks)
  cudaEvent_t ev_in[kHostChunksMax] = {}, ev_k[kHostChunksMax] = {};
  int64_t launches;
  int kernel;          // 0 = dense (generic), 1 = srb (structured, thread per problem), 2 = coop (structured,
                       //     16 lanes per problem, shared-memory resident; default for the QUAT models)
  int coop_grid, coop_smem_doubles, coop_wide, coop_blocks_per_sm, coop_sms, coop_last_grid = 0, coop_last_active = 0;   // persistent launch geometry of the coop kernel
  size_t coop_scratch_doubles;
  // phased kernels (QMPC_KERNEL_PHASED): backward launch uses coop_grid / coop_smem_doubles / coop_wide above
  int ph_fwd_grid, ph_fwd_smem_doubles, ph_fwd_blocks_per_sm, ph_chunk;
  size_t ph_problem_doubles, ph_trial_doubles;
  double* ph_trial;
  char err[256];
};

constexpr int kStageBatch = 64;   // host calls up to this batch are staged through pinned memory

static int set_err(QmpcHandle* h, cudaError_t e, const char* where) {
  if (h) snprintf(h->err, sizeof(h->err), "%s: %s", where, cudaGetErrorString(e));
  return QMPC_ERR_CUDA;
}
#define CU(call)                                       \
  do {                                                 \
    cudaError_t e_ = (call);                           \
    if (e_ != cudaSuccess) return set_err(h, e_, #call); \
  } while (0)

extern "C" int32_t qmpc_abi_version(void) { return QMPC_ABI_VERSION; }

extern "C" const char* qmpc_status_string(int32_t s) {
  switch (s) {
    case QMPC_STATUS_SUCCESS: return "success";
    case QMPC_STATUS_MAX_ITERATIONS: return "max_iterations";
    case QMPC_STATUS_LINESEARCH_FAILED: return "linesearch_failed";
    case QMPC_STATUS_BACKWARD_FAILED: return "backward_failed";
    case QMPC_STATUS_NONFINITE: return "nonfinite";
  }
  return "unknown";
}

extern "C" int qmpc_default_config(int32_t model, int32_t horizon, QmpcConfig* c) {
  if (!c) return QMPC_ERR_ARG;
  if (model != QMPC_MODEL_QUAT_4FOOT && model != QMPC_MODEL_QUAT_2FOOT && model != QMPC_MODEL_EULER_CONVEX)
    return QMPC_ERR_ARG;
  if (horizon < 1 || horizon > QMPC_MAX_HORIZON) return QMPC_ERR_ARG;
  memset(c, 0, sizeof(*c));
  c->model = model;
  c->horizon = horizon;
  c->robot_mass = 12.84;  // gazebo_go1_quat_mpc.yaml:115
  c->gravity = 9.81;
  c->quat_d_dt = 5.0 / 1000.0;  // QuatMpc.cpp:132
  c->com_offset[0] = 0.0223; c->com_offset[1] = 0.002; c->com_offset[2] = -0.0005;  // AltroUtils.cpp:373
  c->com_mass = 5.204;
  for (int i = 0; i < 12; ++i) c->r_weights[i] = 1e-6;  // yaml:58-72
  c->penalty_initial = 1.0;
  c->penalty_max = 1e8;
  c->tol_cost_intermediate = c->tol_primal_feasibility = c->tol_stationarity = 1e-4;
  c->drop_omega0 = 1;
  const double It[3] = {0.0168128557, 0.063009565, 0.0716547275};  // yaml:117-122
  double scale;
  if (model == QMPC_MODEL_EULER_CONVEX) {
    c->dt = 5.0 / 1000.0;  // gazebo_go1_convex_mpc.yaml:36
    const double q[13] = {3.0, 3.0, 3.0, 1.0, 1.0, 20.0, 0.0, 0.0, 3.0, 2.0, 3.0, 2.0, 0.0};
    memcpy(c->q_weights, q, sizeof(q));
    c->w = 0.0; c->mu = 0.6; c->fz_max = 200.0;
    scale = 1.0;
    c->iterations_max = 5;      // ConvexMpc.cpp:37
    c->penalty_scaling = 10.0;  // ALTRO default
  } else {
    c->dt = 10.0 / 1000.0;  // gazebo_go1_quat_mpc.yaml:36
    const double q[13] = {2.5, 2.5, 10.0, 0, 0, 0, 0, 0.1, 0.1, 0.1, 0.15, 0.15, 0.15};
    memcpy(c->q_weights, q, sizeof(q));
    c->w = 50.0; c->mu = 0.7; c->fz_max = 100.0;
    scale = 1.2;                // QuatMpc.cpp:182
    c->iterations_max = 10;     // QuatMpc.cpp:22
    c->penalty_scaling = 20.0;  // QuatMpc.cpp:26
  }
  for (int i = 0; i < 3; ++i) c->inertia[4 * i] = scale * It[i];
  return QMPC_OK;
}

static size_t ws_elems(const QmpcConfig& c, int kernel) {
  switch (c.model) {
    case QMPC_MODEL_QUAT_4FOOT:
      return kernel == 1 ? SrbLayout<4>::total(c.horizon) : DenseLayout<QuatModel<4>>::total(c.horizon);
    case QMPC_MODEL_QUAT_2FOOT:
      return kernel == 1 ? SrbLayout<2>::total(c.horizon) : DenseLayout<QuatModel<2>>::total(c.horizon);
    default: return DenseLayout<ConvexModel>::total(c.horizon);
  }
}

constexpr int kCoopG = 16, kCoopBlock = QMPC_COOP_BLOCK;

// persistent-kernel geometry: as many resident blocks as the device holds (or the batch needs)
template <class M>
static int coop_prepare_t(QmpcHandle* h, int smem_residents) {
  using L = CoopLayout<M, kCoopG>;
  const int N = h->cfg.horizon;
  const int groups = kCoopBlock / kCoopG;
  const bool phased = h->kernel == QMPC_KERNEL_PHASED;
  auto smem_bytes_of = [&](int flags) {
    return (size_t)(groups * L::smem_doubles(N, flags) + kCoopBlockShared) * sizeof(double);
  };
  auto blocks_per_sm = [&](int flags, int* per_sm) -> int {
    const size_t b = smem_bytes_of(flags);
    *per_sm = 0;
    if (b > 227 * 1024) return QMPC_OK;
    if (phased) {
      CU(cudaFuncSetAttribute(qmpc_phased_backward_kernel<M, kCoopG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b));
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, qmpc_phased_backward_kernel<M, kCoopG>, kCoopBlock, b));
    } else {
      CU(cudaFuncSetAttribute(qmpc_coop_kernel<M, kCoopG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b));
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, qmpc_coop_kernel<M, kCoopG>, kCoopBlock, b));
    }
    return QMPC_OK;
  };
  if (phased)
    CU(cudaFuncSetAttribute(qmpc_phased_backward_kernel<M, kCoopG>, cudaFuncAttributePreferredSharedMemoryCarveout,
                            (int)cudaSharedmemCarveoutMaxShared));
  else
    CU(cudaFuncSetAttribute(qmpc_coop_kernel<M, kCoopG>, cudaFuncAttributePreferredSharedMemoryCarveout,
                            (int)cudaSharedmemCarveoutMaxShared));
  // Residency is what this latency-bound kernel lives on: first find the block count the bare layout
  // reaches (registers cap it at QMPC_COOP_MIN_BLOCKS), then keep the duals / linearisation blocks in
  // shared memory too if that does not cost a block.  (Phased: the duals cross launches and stay in the
  // problem block; only the linearisation blocks can be residents.)
  int best = 0, rc;
  if ((rc = blocks_per_sm(0, &best))) return rc;
  if (best < 1) { snprintf(h->err, sizeof(h->err), "coop kernel does not fit on an SM"); return QMPC_ERR_CUDA; }
  int flags = 0;
  const int order[3] = {3, 2, 1};
  for (int c = 0; c < 3; ++c) {
    if (phased && order[c] != 1) continue;
    int per = 0;
    if ((rc = blocks_per_sm(order[c], &per))) return rc;
    if (per >= best) { flags = order[c]; break; }
  }
  if (smem_residents >= 0) flags = smem_residents & (phased ? 1 : 3);
  int per_sm = 0, sms = 0;
  if ((rc = blocks_per_sm(flags, &per_sm))) return rc;
  if (per_sm < 1) { snprintf(h->err, sizeof(h->err), "coop kernel does not fit on an SM"); return QMPC_ERR_CUDA; }
  h->coop_wide = flags;
  h->coop_smem_doubles = L::smem_doubles(N, flags);
  h->coop_scratch_doubles = L::scratch_doubles(N);
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  // every resident block may be launched whatever the batch (launch_coop spreads a partial wave over all of
  // them); the scratch holds one slot per problem in flight
  const int resident = per_sm * sms;
  h->coop_grid = resident;
  h->coop_sms = sms;
  h->coop_blocks_per_sm = per_sm;
  if (!phased) {
    const long long slots_cap = (long long)resident * groups;
    const long long ws_slots = (h->max_batch < slots_cap ? h->max_batch : slots_cap) + groups;
    h->ws_bytes = (size_t)ws_slots * h->coop_scratch_doubles * sizeof(double);
    return QMPC_OK;
  }
  // ---- phased: forward-kernel geometry, per-problem blocks (a batch is processed in chunks so that the
  //      workspace stays bounded: 37 KB per problem at N = 10), per-slot trial trajectories
  h->ph_fwd_smem_doubles = L::fwd_smem_doubles(N);
  const size_t fb = (size_t)(groups * h->ph_fwd_smem_doubles + kCoopBlockShared) * sizeof(double);
  if (fb > 227 * 1024) { snprintf(h->err, sizeof(h->err), "forward kernel does not fit on an SM"); return QMPC_ERR_CUDA; }
  CU(cudaFuncSetAttribute(qmpc_phased_forward_kernel<M, kCoopG>, cudaFuncAttributePreferredSharedMemoryCarveout,
                          (int)cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(qmpc_phased_forward_kernel<M, kCoopG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fb));
  int fper = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fper, qmpc_phased_forward_kernel<M, kCoopG>, kCoopBlock, fb));
  if (fper < 1) { snprintf(h->err, sizeof(h->err), "forward kernel does not fit on an SM"); return QMPC_ERR_CUDA; }
  h->ph_fwd_blocks_per_sm = fper;
  h->ph_fwd_grid = fper * sms;
  h->ph_problem_doubles = L::problem_doubles(N);
  h->ph_trial_doubles = L::trial_doubles(N);
  h->ph_chunk = h->max_batch < 65536 ? h->max_batch : 65536;
  h->ws_bytes = (size_t)h->ph_chunk * h->ph_problem_doubles * sizeof(double);
  return QMPC_OK;
}
static int coop_prepare(QmpcHandle* h, int smem_residents) {
  switch (h->cfg.model) {
    case QMPC_MODEL_QUAT_4FOOT: return coop_prepare_t<QuatModel<4>>(h, smem_residents);
    case QMPC_MODEL_QUAT_2FOOT: return coop_prepare_t<QuatModel<2>>(h, smem_residents);
    default: return coop_prepare_t<ConvexModel>(h, smem_residents);
  }
}

extern "C" int qmpc_create_ex(const QmpcConfig* cfg, int32_t max_batch, int32_t device, const QmpcCreateOptions* opt,
                              QmpcHandle** out) {
  if (!cfg || !out || max_batch < 1) return QMPC_ERR_ARG;
  QmpcCreateOptions op = {QMPC_KERNEL_AUTO, -1, 0, 0};
  if (opt) op = *opt;
  if (op.kernel < QMPC_KERNEL_AUTO || op.kernel > QMPC_KERNEL_PHASED) return QMPC_ERR_ARG;
  if (op.kernel == QMPC_KERNEL_SRB && cfg->model == QMPC_MODEL_EULER_CONVEX) return QMPC_ERR_ARG;
#ifdef QMPC_NO_XCHECK   // the product library: the cross-check kernels are compiled into the test-only sibling only
  if (op.kernel == QMPC_KERNEL_DENSE || op.kernel == QMPC_KERNEL_SRB) return QMPC_ERR_ARG;
#endif

  if (cfg->horizon < 1 || cfg->horizon > QMPC_MAX_HORIZON) return QMPC_ERR_ARG;
  if (cfg->model < 0 || cfg->model > QMPC_MODEL_EULER_CONVEX) return QMPC_ERR_ARG;
  if (cfg->iterations_max < 0 || !(cfg->penalty_initial > 0) || !(cfg->dt > 0)) return QMPC_ERR_ARG;
  QmpcHandle* h = new (std::nothrow) QmpcHandle();
  if (!h) return QMPC_ERR_ARG;
  memset(h, 0, sizeof(*h));
  *out = h;  // returned even on CUDA failure so the caller can read qmpc_last_error(); destroy is safe
  h->cfg = *cfg;
  h->device = device;
  h->max_batch = max_batch;
  h->stride = ((size_t)max_batch + 31) / 32 * 32;
  SolverOpts& o = h->opts;
  o.N = cfg->horizon;
  o.iterations_max = cfg->iterations_max;
  o.h = (float)cfg->dt;  // ALTRO takes the step as float (AltroUtils.cpp:10)
  o.penalty_initial = cfg->penalty_initial;
  o.penalty_scaling = cfg->penalty_scaling;
  o.penalty_max = cfg->penalty_max;
  o.tol_cost_intermediate = cfg->tol_cost_intermediate;
  o.tol_primal_feasibility = cfg->tol_primal_feasibility;
  o.tol_stationarity = cfg->tol_stationarity;
  o.ls_c1 = 1e-4;
  o.ls_decrease = 0.5;
  o.ls_iters_max = 25;
  // kernel selection, resolved once (no environment variables anywhere in the library): the cooperative
  // kernel for every model; the dense and srb kernels only on explicit request (the tests' on-device
  // cross-checks), the phased launches likewise
  h->kernel = op.kernel != QMPC_KERNEL_AUTO ? op.kernel : QMPC_KERNEL_COOP;
  h->packed_launch = op.packed_launch;
  h->host_chunks = op.host_chunks;
  CU(cudaSetDevice(device));
  if (h->kernel == QMPC_KERNEL_COOP || h->kernel == QMPC_KERNEL_PHASED) {
    int rc = coop_prepare(h, op.smem_residents);
    if (rc) return rc;
    if (h->kernel == QMPC_KERNEL_PHASED)
      CU(cudaMalloc(&h->ph_trial, (size_t)h->ph_fwd_grid * (kCoopBlock / kCoopG) * h->ph_trial_doubles * sizeof(double)));
  } else {
    h->ws_bytes = ws_elems(*cfg, h->kernel) * h->stride * sizeof(double);
  }
  CU(cudaMalloc(&h->ws, h->ws_bytes));
  // once, at creation: the 16-byte cp.async copies of the knot rows / linearisation blocks also move the padding double
  // at the end of an odd-length row, which no phase ever writes (compute-sanitizer initcheck reports exactly those reads)
  CU(cudaMemset(h->ws, 0, h->ws_bytes));
  size_t in_sz = cfg->model == QMPC_MODEL_EULER_CONVEX ? sizeof(QmpcConvexProblem) : sizeof(QmpcProblem);
  CU(cudaMalloc(&h->d_in, in_sz * (size_t)max_batch));
  CU(cudaMalloc(&h->d_out, sizeof(QmpcResult) * (size_t)max_batch));
  CU(cudaMalloc(&h->d_sched, sizeof(QmpcContactSchedule) * (size_t)max_batch));
  CU(cudaHostAlloc(&h->h_stage, (size_t)kStageBatch * (sizeof(QmpcConvexProblem) + sizeof(QmpcContactSchedule) + sizeof(QmpcResult)),
                   cudaHostAllocDefault));
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < kHostChunksMax; ++i) {
    CU(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming));
  }
  return QMPC_OK;
}

extern "C" int qmpc_create(const QmpcConfig* cfg, int32_t max_batch, int32_t device, QmpcHandle** out) {
  return qmpc_create_ex(cfg, max_batch, device, nullptr, out);
}

extern "C" void qmpc_destroy(QmpcHandle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; i < kHostChunksMax; ++i) {
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]);
  }
  if (h->ws) cudaFree(h->ws);
  if (h->ph_trial) cudaFree(h->ph_trial);
  if (h->d_in) cudaFree(h->d_in);
  if (h->d_out) cudaFree(h->d_out);
  if (h->d_sched) cudaFree(h->d_sched);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  delete h;
}

extern "C" int64_t qmpc_launch_count(const QmpcHandle* h) { return h ? h->launches : 0; }
extern "C" const char* qmpc_last_error(const QmpcHandle* h) { return h ? h->err : "null handle"; }

extern "C" int qmpc_describe(const QmpcHandle* h, char* buf, int32_t n) {
  if (!h || !buf || n < 1) return QMPC_ERR_ARG;
  const char* names[3] = {"dense", "srb", "coop"};
  if (h->kernel == QMPC_KERNEL_PHASED)
    snprintf(buf, n, "kernel=phased lanes_per_problem=%d block=%d launches_per_solve=%d backward:blocks_per_sm=%d,smem_per_problem=%dB%s "
                     "forward:blocks_per_sm=%d,smem_per_problem=%dB problem_block=%zuB chunk=%d",
             kCoopG, kCoopBlock, 1 + 2 * h->cfg.iterations_max, h->coop_blocks_per_sm, h->coop_smem_doubles * 8,
             (h->coop_wide & 1) ? ",lin_resident" : "", h->ph_fwd_blocks_per_sm, h->ph_fwd_smem_doubles * 8,
             h->ph_problem_doubles * 8, h->ph_chunk);
  else if (h->kernel == 2)
    snprintf(buf, n, "kernel=coop lanes_per_problem=%d block=%d blocks_per_sm=%d grid=%d last_launch=%dblocks_x_%dproblems "
                     "smem_per_problem=%dB smem_residents=%s%s scratch_per_slot=%zuB",
             kCoopG, kCoopBlock, h->coop_blocks_per_sm, h->coop_grid, h->coop_last_grid, h->coop_last_active,
             h->coop_smem_doubles * 8,
             (h->coop_wide & 1) ? "lin" : "", (h->coop_wide & 2) ? "+duals" : "", h->coop_scratch_doubles * 8);
  else
    snprintf(buf, n, "kernel=%s threads_per_problem=1 workspace=%zuB", names[h->kernel], h->ws_bytes);
  return QMPC_OK;
}

#ifndef QMPC_NO_XCHECK   // experiment builds (-DQMPC_NO_XCHECK) leave the cross-check kernels out: faster to compile
template <class M>
static int launch_dense(QmpcHandle* h, const typename M::Problem* d_in, const unsigned char* sched,
                        QmpcWarmStart* warm, int batch,
                        QmpcResult* d_out, cudaStream_t s) {
  const int block = 64;
  const int grid = (batch + block - 1) / block;
  qmpc_dense_kernel<M><<<grid, block, 0, s>>>(h->cfg, h->opts, d_in, sched, warm, d_out, h->ws, batch, h->stride);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

template <int NF>
static int launch_srb(QmpcHandle* h, const QmpcProblem* d_in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                      QmpcResult* d_out,
                      cudaStream_t s) {
  const int block = 64;
  const int grid = (batch + block - 1) / block;
  qmpc_srb_kernel<NF><<<grid, block, 0, s>>>(h->cfg, h->opts, d_in, sched, warm, d_out, h->ws, batch, h->stride);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

#endif

template <class M>
static int launch_coop(QmpcHandle* h, const typename M::Problem* d_in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                       QmpcResult* d_out,
                       cudaStream_t s) {
  const int groups = kCoopBlock / kCoopG;
  // Persistent slots stride over the batch.  Balance the waves: with S resident slots a batch needs
  // w = ceil(batch / S) passes, so launch only ceil(batch / w) slots - every slot then solves w (or
  // w - 1) problems and no SM idles through a mostly empty last pass at full-residency latency.
  // Spread them: the slots are dealt over the blocks (`active` groups per block, the others only pass
  // the barriers) rather than filling blocks one by one - a solve is faster the fewer problems share its
  // SM (1.29 ms alone, 1.77 ms among 16) - but over one block per SM as long as that holds the wave: a
  // second, unaligned block on the SM costs more (instruction cache) than a fuller first one (measured
  // at batch 1024: 1.48 ms in 128 blocks of 8 against 1.64 ms in 256 blocks of 4).
  const long long slots_max = (long long)h->coop_grid * groups;
  const long long waves = (batch + slots_max - 1) / slots_max;
  const long long slots = (batch + waves - 1) / waves;
  const int blocks_cap = slots <= (long long)h->coop_sms * groups ? h->coop_sms : h->coop_grid;
  int grid = slots < blocks_cap ? (int)slots : blocks_cap;
  int active = (int)((slots + grid - 1) / grid);
  if (h->packed_launch) active = groups;
  grid = (int)((slots + active - 1) / active);
  h->coop_last_grid = grid;
  h->coop_last_active = active;
  const size_t smem_bytes = (size_t)(groups * h->coop_smem_doubles + kCoopBlockShared) * sizeof(double);
  // the opt-in shared-memory limit is per function AND per device, not per handle: another handle (other
  // horizon) created or solved in between may have lowered it, so it is set before every launch
  CU(cudaFuncSetAttribute(qmpc_coop_kernel<M, kCoopG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  qmpc_coop_kernel<M, kCoopG><<<grid, kCoopBlock, smem_bytes, s>>>(h->cfg, h->opts, d_in, sched, warm, d_out, h->ws, batch,
                                                                  h->coop_smem_doubles, h->coop_scratch_doubles,
                                                                  h->coop_wide, active);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

// phased: set-up launch, then (backward, forward) per iteration; the batch in chunks of ph_chunk problems
template <class M>
static int launch_phased(QmpcHandle* h, const typename M::Problem* d_in, const unsigned char* sched, QmpcWarmStart* warm, int batch,
                         QmpcResult* d_out, cudaStream_t s) {
  const int groups = kCoopBlock / kCoopG;
  const size_t bsm = (size_t)(groups * h->coop_smem_doubles + kCoopBlockShared) * sizeof(double);
  const size_t fsm = (size_t)(groups * h->ph_fwd_smem_doubles + kCoopBlockShared) * sizeof(double);
  CU(cudaFuncSetAttribute(qmpc_phased_backward_kernel<M, kCoopG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
  CU(cudaFuncSetAttribute(qmpc_phased_forward_kernel<M, kCoopG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
  auto balanced_grid = [&](int n, int resident_blocks) {
    // w passes over the resident slots: launch ceil(n / w) slots so that every pass is equally full
    const long long slots_max = (long long)resident_blocks * groups;
    const long long waves = (n + slots_max - 1) / slots_max;
    const long long slots = (n + waves - 1) / waves;
    return (int)((slots + groups - 1) / groups);
  };
  for (int base = 0; base < batch; base += h->ph_chunk) {
    const int n = batch - base < h->ph_chunk ? batch - base : h->ph_chunk;
    const typename M::Problem* in = d_in + base;
    const unsigned char* sc = sched ? sched + (size_t)base * QMPC_MAX_HORIZON : nullptr;
    QmpcWarmStart* wm = warm ? warm + base : nullptr;
    QmpcResult* out = d_out + base;
    qmpc_phased_setup_kernel<M, kCoopG><<<(n + 127) / 128, 128, 0, s>>>(h->cfg, h->opts, in, sc, wm, out, h->ws, n,
                                                                         h->ph_problem_doubles);
    h->launches += 1;
    const int bgrid = balanced_grid(n, h->coop_grid), fgrid = balanced_grid(n, h->ph_fwd_grid);
    for (int it = 0; it < h->opts.iterations_max; ++it) {
      qmpc_phased_backward_kernel<M, kCoopG><<<bgrid, kCoopBlock, bsm, s>>>(h->cfg, h->opts, it, wm, out, h->ws, n,
                                                                            h->ph_problem_doubles, h->coop_smem_doubles, h->coop_wide);
      qmpc_phased_forward_kernel<M, kCoopG><<<fgrid, kCoopBlock, fsm, s>>>(h->cfg, h->opts, it, wm, out, h->ws, h->ph_trial, n,
                                                                           h->ph_problem_doubles, h->ph_fwd_smem_doubles,
                                                                           h->ph_trial_doubles);
      h->launches += 2;
    }
    h->coop_last_grid = bgrid;
    h->coop_last_active = fgrid;
  }
  CU(cudaGetLastError());
  return QMPC_OK;
}

static int solve_any(QmpcHandle* h, const void* d_in, const QmpcContactSchedule* d_sched, QmpcWarmStart* warm, int32_t batch,
                     QmpcResult* d_out, void* stream, bool convex) {
  if (!h || !h->ws) return h ? QMPC_ERR_CUDA : QMPC_ERR_ARG;
  if (!d_in || !d_out || batch < 0) return QMPC_ERR_ARG;
  if (convex != (h->cfg.model == QMPC_MODEL_EULER_CONVEX)) return QMPC_ERR_ARG;
  if (batch > h->max_batch) return QMPC_ERR_CAPACITY;
  if (batch == 0) return QMPC_OK;
  CU(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned char* sc = reinterpret_cast<const unsigned char*>(d_sched);
#ifdef QMPC_NO_XCHECK
  if (h->kernel < 2) return QMPC_ERR_ARG;
#define QMPC_XCHECK(call_) QMPC_ERR_ARG
#else
#define QMPC_XCHECK(call_) call_
#endif
  switch (h->cfg.model) {
    case QMPC_MODEL_QUAT_4FOOT:
      if (h->kernel == 3) return launch_phased<QuatModel<4>>(h, (const QmpcProblem*)d_in, sc, warm, batch, d_out, s);
      if (h->kernel == 2) return launch_coop<QuatModel<4>>(h, (const QmpcProblem*)d_in, sc, warm, batch, d_out, s);
      if (h->kernel == 1) return QMPC_XCHECK(launch_srb<4>(h, (const QmpcProblem*)d_in, sc, warm, batch, d_out, s));
      return QMPC_XCHECK(launch_dense<QuatModel<4>>(h, (const QmpcProblem*)d_in, sc, warm, batch, d_out, s));
    case QMPC_MODEL_QUAT_2FOOT:
      if (h->kernel == 3) return launch_phased<QuatModel<2>>(h, (const QmpcProblem*)d_in, sc, warm, batch, d_out, s);
      if (h->kernel == 2) return launch_coop<QuatModel<2>>(h, (const QmpcProblem*)d_in, sc, warm, batch, d_out, s);
      if (h->kernel == 1) return QMPC_XCHECK(launch_srb<2>(h, (const QmpcProblem*)d_in, sc, warm, batch, d_out, s));
      return QMPC_XCHECK(launch_dense<QuatModel<2>>(h, (const QmpcProblem*)d_in, sc, warm, batch, d_out, s));
    default:
      if (h->kernel == 3) return launch_phased<ConvexModel>(h, (const QmpcConvexProblem*)d_in, sc, nullptr, batch, d_out, s);
      if (h->kernel == 2) return launch_coop<ConvexModel>(h, (const QmpcConvexProblem*)d_in, sc, nullptr, batch, d_out, s);
      return QMPC_XCHECK(launch_dense<ConvexModel>(h, (const QmpcConvexProblem*)d_in, sc, warm, batch, d_out, s));
  }
#undef QMPC_XCHECK
}

extern "C" int qmpc_solve_batch(QmpcHandle* h, const QmpcProblem* d_in, int32_t batch, QmpcResult* d_out,
                                void* cuda_stream) {
  return solve_any(h, d_in, nullptr, nullptr, batch, d_out, cuda_stream, false);
}
extern "C" int qmpc_solve_batch_sched(QmpcHandle* h, const QmpcProblem* d_in, const QmpcContactSchedule* d_sched,
                                      int32_t batch, QmpcResult* d_out, void* cuda_stream) {
  return solve_any(h, d_in, d_sched, nullptr, batch, d_out, cuda_stream, false);
}
extern "C" int qmpc_solve_batch_warm(QmpcHandle* h, const QmpcProblem* d_in, const QmpcContactSchedule* d_sched,
                                     QmpcWarmStart* d_warm, int32_t batch, QmpcResult* d_out, void* cuda_stream) {
  if (!d_warm) return QMPC_ERR_ARG;
  return solve_any(h, d_in, d_sched, d_warm, batch, d_out, cuda_stream, false);
}
extern "C" int qmpc_solve_batch_convex_sched(QmpcHandle* h, const QmpcConvexProblem* d_in,
                                             const QmpcContactSchedule* d_sched, int32_t batch, QmpcResult* d_out,
                                             void* cuda_stream) {
  return solve_any(h, d_in, d_sched, nullptr, batch, d_out, cuda_stream, true);
}
extern "C" int qmpc_solve_batch_convex(QmpcHandle* h, const QmpcConvexProblem* d_in, int32_t batch,
                                       QmpcResult* d_out, void* cuda_stream) {
  return solve_any(h, d_in, nullptr, nullptr, batch, d_out, cuda_stream, true);
}

static bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

static int solve_host_any(QmpcHandle* h, const void* in, const QmpcContactSchedule* sched, int32_t batch,
                          QmpcResult* out, bool convex) {
  if (!h || !h->ws) return h ? QMPC_ERR_CUDA : QMPC_ERR_ARG;
  if (!in || !out || batch < 0) return QMPC_ERR_ARG;
  // the entry point must match the handle's model BEFORE anything is copied: the staging buffers are
  // sized for the handle's own problem struct
  if (convex != (h->cfg.model == QMPC_MODEL_EULER_CONVEX)) return QMPC_ERR_ARG;
  if (batch > h->max_batch) return QMPC_ERR_CAPACITY;
  if (batch == 0) return QMPC_OK;
  CU(cudaSetDevice(h->device));
  const size_t in_sz = convex ? sizeof(QmpcConvexProblem) : sizeof(QmpcProblem);
  // small batches from pageable memory (the batch-1 call of the 200 Hz mpc_thread passes stack objects)
  // go through the handle's pinned staging: both copies are then truly asynchronous DMA transfers
  const bool stage = batch <= kStageBatch && h->h_stage && !(host_ptr_is_pinned(in) && host_ptr_is_pinned(out));
  char* st_in = (char*)h->h_stage;
  char* st_sched = st_in + (size_t)kStageBatch * sizeof(QmpcConvexProblem);
  char* st_out = st_sched + (size_t)kStageBatch * sizeof(QmpcContactSchedule);
  const void* src_in = in;
  const void* src_sched = sched;
  void* dst_out = out;
  if (stage) {
    memcpy(st_in, in, in_sz * batch);
    src_in = st_in;
    if (sched) { memcpy(st_sched, sched, sizeof(QmpcContactSchedule) * batch); src_sched = st_sched; }
    dst_out = st_out;
  }
  // Opt-in (QmpcCreateOptions.host_chunks = 2..4): a batch of several problem waves (more problems than the persistent
  // kernel has slots) is copied and solved in chunks of whole waves - chunk i + 1's problems go up and chunk i - 1's
  // results come down on a second stream while chunk i is being solved, so that only the first copy in and the last
  // copy out are exposed.  Results are identical (a problem's solve does not depend on its neighbours).  NOT the
  // default: measured on the B200 (run 18) the hidden copies are worth less than the extra launch boundaries cost -
  // inside one launch a block starts its next wave the moment it finishes, across launches every block waits for the
  // slowest (batch 4096 = 2 waves: 1.60 M against 1.67 M solves/s end to end; batch 65 536 in 4 chunks: no difference).
  int nchunks = 1;
  if (h->kernel == QMPC_KERNEL_COOP && !stage && h->copy_stream && h->host_chunks > 1) {
    const long long slots = (long long)h->coop_grid * (kCoopBlock / kCoopG);
    const long long waves = (batch + slots - 1) / slots;
    const int cap = h->host_chunks < kHostChunksMax ? h->host_chunks : kHostChunksMax;
    nchunks = (int)(waves < cap ? waves : cap);
  }
  if (nchunks <= 1) {
    CU(cudaMemcpyAsync(h->d_in, src_in, in_sz * batch, cudaMemcpyHostToDevice, h->stream));
    if (sched)
      CU(cudaMemcpyAsync(h->d_sched, src_sched, sizeof(QmpcContactSchedule) * batch, cudaMemcpyHostToDevice, h->stream));
    int rc = solve_any(h, h->d_in, sched ? h->d_sched : nullptr, nullptr, batch, h->d_out, h->stream, convex);
    if (rc) return rc;
    CU(cudaMemcpyAsync(dst_out, h->d_out, sizeof(QmpcResult) * batch, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (stage) memcpy(out, st_out, sizeof(QmpcResult) * batch);
    return QMPC_OK;
  }
  const int32_t per = (batch + nchunks - 1) / nchunks;
  auto lo = [&](int i) { long long v = (long long)per * i; return (int32_t)(v < batch ? v : batch); };
  auto copy_in = [&](int i, cudaStream_t st) -> int {
    const int32_t a = lo(i), n = lo(i + 1) - a;
    if (n <= 0) return QMPC_OK;
    CU(cudaMemcpyAsync((char*)h->d_in + in_sz * a, (const char*)src_in + in_sz * a, in_sz * n, cudaMemcpyHostToDevice, st));
    if (sched)
      CU(cudaMemcpyAsync(h->d_sched + a, (const QmpcContactSchedule*)src_sched + a, sizeof(QmpcContactSchedule) * n,
                         cudaMemcpyHostToDevice, st));
    return QMPC_OK;
  };
  int rc = copy_in(0, h->stream);
  if (rc) return rc;
  for (int i = 0; i < nchunks; ++i) {
    const int32_t a = lo(i), n = lo(i + 1) - a;
    if (n <= 0) break;
    if (i + 1 < nchunks) {
      if ((rc = copy_in(i + 1, h->copy_stream))) return rc;
      CU(cudaEventRecord(h->ev_in[i + 1], h->copy_stream));
    }
    if (i > 0) CU(cudaStreamWaitEvent(h->stream, h->ev_in[i], 0));
    rc = solve_any(h, (char*)h->d_in + in_sz * a, sched ? h->d_sched + a : nullptr, nullptr, n, h->d_out + a, h->stream, convex);
    if (rc) return rc;
    CU(cudaEventRecord(h->ev_k[i], h->stream));
    CU(cudaStreamWaitEvent(h->copy_stream, h->ev_k[i], 0));
    CU(cudaMemcpyAsync((QmpcResult*)dst_out + a, h->d_out + a, sizeof(QmpcResult) * n, cudaMemcpyDeviceToHost, h->copy_stream));
  }
  CU(cudaStreamSynchronize(h->copy_stream));
  CU(cudaStreamSynchronize(h->stream));
  return QMPC_OK;
}

extern "C" int qmpc_solve_batch_host(QmpcHandle* h, const QmpcProblem* in, int32_t batch, QmpcResult* out) {
  return solve_host_any(h, in, nullptr, batch, out, false);
}
extern "C" int qmpc_solve_batch_sched_host(QmpcHandle* h, const QmpcProblem* in, const QmpcContactSchedule* sched,
                                           int32_t batch, QmpcResult* out) {
  return solve_host_any(h, in, sched, batch, out, false);
}
extern "C" int qmpc_solve_batch_convex_host(QmpcHandle* h, const QmpcConvexProblem* in, int32_t batch,
                                            QmpcResult* out) {
  return solve_host_any(h, in, nullptr, batch, out, true);
}
extern "C" int qmpc_solve_batch_convex_sched_host(QmpcHandle* h, const QmpcConvexProblem* in,
                                                  const QmpcContactSchedule* sched, int32_t batch, QmpcResult* out) {
  return solve_host_any(h, in, sched, batch, out, true);
}

// ------------------------------------------------------------------------------------------------
// Rows N1 / N2 of the scope table: streaming kernels either side of the solve (qmpc_periph.cuh)
static int periph_check(QmpcHandle* h, int32_t batch) {
  if (!h) return QMPC_ERR_ARG;
  if (batch < 0) return QMPC_ERR_ARG;
  if (batch > h->max_batch) return QMPC_ERR_CAPACITY;
  return QMPC_OK;
}

extern "C" int qmpc_predict_contact_schedule(QmpcHandle* h, const QmpcGaitState* d_gait, int32_t batch,
                                             QmpcContactSchedule* d_sched, void* cuda_stream) {
  int rc = periph_check(h, batch);
  if (rc) return rc;
  if (!d_gait || !d_sched) return QMPC_ERR_ARG;
  if (batch == 0) return QMPC_OK;
  CU(cudaSetDevice(h->device));
  qmpc_predict_schedule_kernel<<<(4 * batch + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(
      d_gait, batch, h->cfg.horizon, h->cfg.dt, d_sched);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

extern "C" int qmpc_default_leg_params(QmpcLegParams* lp) {
  if (!lp) return QMPC_ERR_ARG;
  memset(lp, 0, sizeof(*lp));
  for (int i = 0; i < 4; ++i) {
    lp->rho_fix[i][0] = i < 2 ? 0.1881 : -0.1881;        // leg_offset_x   BaseInterface.cpp:12-15
    lp->rho_fix[i][1] = (i & 1) ? -0.04675 : 0.04675;    // leg_offset_y   :16-19
    lp->rho_fix[i][2] = (i & 1) ? -0.0812 : 0.0812;      // motor_offset   :20-23
    lp->rho_fix[i][3] = 0.213;                           // UPPER_LEG_LENGTH  LeggedParams.h:14
    lp->rho_fix[i][4] = 0.213;                           // LOWER_LEG_LENGTH  LeggedParams.h:15
  }
  return QMPC_OK;
}

extern "C" int qmpc_leg_kinematics(QmpcHandle* h, const QmpcLegParams* lp, const double* d_joint_pos, int32_t batch,
                                   double* d_foot_pos_body, double* d_jac_foot, void* cuda_stream) {
  int rc = periph_check(h, batch);
  if (rc) return rc;
  if (!lp || !d_joint_pos || (!d_foot_pos_body && !d_jac_foot)) return QMPC_ERR_ARG;
  if (batch == 0) return QMPC_OK;
  CU(cudaSetDevice(h->device));
  const int n = batch * 4;
  qmpc_leg_kinematics_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(*lp, d_joint_pos, batch,
                                                                                    d_foot_pos_body, d_jac_foot);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

extern "C" int qmpc_joint_torques(QmpcHandle* h, const QmpcResult* d_results, const double* d_jac_foot,
                                  const int32_t* d_plan_contacts, int32_t movement_mode, int32_t batch, double* d_tau,
                                  void* cuda_stream) {
  int rc = periph_check(h, batch);
  if (rc) return rc;
  if (!d_results || !d_jac_foot || !d_tau) return QMPC_ERR_ARG;
  if (batch == 0) return QMPC_OK;
  CU(cudaSetDevice(h->device));
  const int n = batch * 4;
  qmpc_joint_torque_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(d_results, d_jac_foot, d_plan_contacts,
                                                                                  movement_mode, batch, d_tau);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

// ---- row N3: reference generation
extern "C" int64_t qmpc_goal_state_bytes(const QmpcHandle* h) {
  return h ? (int64_t)kGoalFields * (int64_t)sizeof(double) * h->max_batch : 0;
}

extern "C" int qmpc_goal_update(QmpcHandle* h, void* d_goal_state, const QmpcGoalInput* d_in, int32_t batch,
                                QmpcProblem* d_problems, void* cuda_stream) {
  int rc = periph_check(h, batch);
  if (rc) return rc;
  if (!d_goal_state || !d_in || !d_problems) return QMPC_ERR_ARG;
  if (batch == 0) return QMPC_OK;
  CU(cudaSetDevice(h->device));
  qmpc_goal_update_kernel<<<(batch + kGoalBlock - 1) / kGoalBlock, kGoalBlock, 0, (cudaStream_t)cuda_stream>>>(
      (double*)d_goal_state, (size_t)h->max_batch, d_in, batch, d_problems);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

// ---- row N3, gait-FSM half
extern "C" int64_t qmpc_leg_fsm_state_bytes(const QmpcHandle* h) {
  return h ? (int64_t)kFsmFields * (int64_t)sizeof(double) * 4 * h->max_batch : 0;
}

extern "C" int qmpc_leg_fsm_init(QmpcHandle* h, void* d_fsm_state, const int32_t* d_gait, int32_t batch, void* cuda_stream) {
  int rc = periph_check(h, batch);
  if (rc) return rc;
  if (!d_fsm_state) return QMPC_ERR_ARG;
  if (batch == 0) return QMPC_OK;
  CU(cudaSetDevice(h->device));
  qmpc_leg_fsm_init_kernel<<<(4 * batch + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>((double*)d_fsm_state,
                                                                                         (size_t)4 * h->max_batch, d_gait, batch);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

extern "C" int qmpc_foot_update(QmpcHandle* h, void* d_fsm_state, const QmpcFootUpdateInput* d_in, double dt, double gait_freq,
                                int32_t batch, QmpcFootUpdateOutput* d_out, QmpcProblem* d_problems, QmpcGaitState* d_gait_out,
                                void* cuda_stream) {
  int rc = periph_check(h, batch);
  if (rc) return rc;
  if (!d_fsm_state || !d_in || !d_out || !(gait_freq > 0) || !(dt > 0)) return QMPC_ERR_ARG;
  if (batch == 0) return QMPC_OK;
  // C^-1 of the quintic swing curve depends on the gait frequency only: once per launch, on the host
  QuinticInv ci;
  if (!quintic_C_inverse((float)(0.5 / gait_freq), ci.m)) return QMPC_ERR_ARG;
  CU(cudaSetDevice(h->device));
  qmpc_foot_update_kernel<<<(4 * batch + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(
      (double*)d_fsm_state, (size_t)4 * h->max_batch, ci, d_in, dt, gait_freq, batch, d_out, d_problems, d_gait_out);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

extern "C" int qmpc_default_raibert_params(QmpcRaibertParams* rp) {
  if (!rp) return QMPC_ERR_ARG;
  memset(rp, 0, sizeof(*rp));
  rp->gait_freq = 2.2;                                  // gazebo_go1_quat_mpc.yaml:33
  const double feet[12] = {0.20, 0.14, -0.30, 0.20, -0.14, -0.30, -0.20, 0.14, -0.30, -0.20, -0.14, -0.30};  // yaml:16-30
  memcpy(rp->default_foot_pos_rel, feet, sizeof(feet));
  rp->delta_x_limit = 0.5;                              // FOOT_DELTA_X_LIMIT  LeggedParams.h:21
  rp->delta_y_limit = 0.3;                              // FOOT_DELTA_Y_LIMIT  LeggedParams.h:22
  return QMPC_OK;
}

extern "C" int qmpc_raibert_targets(QmpcHandle* h, const QmpcRaibertParams* rp, const QmpcGoalInput* d_in, int32_t batch,
                                    double* d_foot_pos_target_world, double* d_foot_pos_target_rel, void* cuda_stream) {
  int rc = periph_check(h, batch);
  if (rc) return rc;
  if (!rp || !d_in || (!d_foot_pos_target_world && !d_foot_pos_target_rel)) return QMPC_ERR_ARG;
  if (batch == 0) return QMPC_OK;
  CU(cudaSetDevice(h->device));
  qmpc_raibert_kernel<<<(batch + kRaibertBlock - 1) / kRaibertBlock, kRaibertBlock, 0, (cudaStream_t)cuda_stream>>>(*rp, d_in, batch, d_foot_pos_target_world,
                                                                                 d_foot_pos_target_rel);
  h->launches += 1;
  CU(cudaGetLastError());
  return QMPC_OK;
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU host entry point: contiguous balanced shards, one stream + pinned staging + ONE HOST
// WORKER THREAD per device (created once in qmpc_create_multi), so that the copies and launches of all devices are
// issued concurrently - a single thread issuing to 8 devices one after another delays the last launch by ~150 us
// of a 3 ms step.  Results land in the caller's one array.  No collective.
constexpr int kMaxDevices = 16;
struct QmpcMultiHandle {
  int n;
  int max_batch;
  bool convex;
  size_t in_sz;
  QmpcHandle* h[kMaxDevices];
  int shard_cap[kMaxDevices];
  void* pin_in[kMaxDevices];     // pinned staging, used when the caller's buffers are pageable
  QmpcResult* pin_out[kMaxDevices];
  char err[256];
  // worker threads: a call publishes (in, out, batch, direct) and bumps `generation`; worker g solves shard g
  std::thread workers[kMaxDevices];
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  long generation = 0;
  int pending = 0;
  bool stop = false;
  const void* job_in = nullptr;
  QmpcResult* job_out = nullptr;
  int job_batch = 0;
  bool job_direct = false;
  int job_rc[kMaxDevices];
  char job_err[kMaxDevices][200];
};

static void multi_shard(int batch, int g, int n, int* lo, int* hi) {
  *lo = (int)((long long)batch * g / n);
  *hi = (int)((long long)batch * (g + 1) / n);
}

// the whole per-device sequence of one call: stage, H2D, solve, D2H, synchronise, un-stage
static int multi_run_shard(QmpcMultiHandle* mh, int g, const void* in, QmpcResult* out, int batch, bool direct, char* err, size_t errn) {
  int lo, hi;
  multi_shard(batch, g, mh->n, &lo, &hi);
  const int cnt = hi - lo;
  if (cnt == 0) return QMPC_OK;
  QmpcHandle* h = mh->h[g];
  const char* src = (const char*)in + mh->in_sz * (size_t)lo;
  if (!direct) { memcpy(mh->pin_in[g], src, mh->in_sz * (size_t)cnt); src = (const char*)mh->pin_in[g]; }
  // the single-device host path (chunked copy / solve pipeline included) on this device's shard
  const int rc = solve_host_any(h, src, nullptr, cnt, direct ? out + lo : (QmpcResult*)mh->pin_out[g], mh->convex);
  if (rc) { snprintf(err, errn, "device %d: %.150s", h->device, h->err); return rc; }
  if (!direct) memcpy(out + lo, mh->pin_out[g], sizeof(QmpcResult) * (size_t)cnt);
  return QMPC_OK;
}

static void multi_worker(QmpcMultiHandle* mh, int g) {
  long seen = 0;
  for (;;) {
    const void* in; QmpcResult* out; int batch; bool direct;
    {
      std::unique_lock<std::mutex> lk(mh->mu);
      mh->cv_go.wait(lk, [&] { return mh->stop || mh->generation != seen; });
      if (mh->stop) return;
      seen = mh->generation;
      in = mh->job_in; out = mh->job_out; batch = mh->job_batch; direct = mh->job_direct;
    }
    mh->job_err[g][0] = 0;
    const int rc = multi_run_shard(mh, g, in, out, batch, direct, mh->job_err[g], sizeof(mh->job_err[g]));
    {
      std::lock_guard<std::mutex> lk(mh->mu);
      mh->job_rc[g] = rc;
      if (--mh->pending == 0) mh->cv_done.notify_one();
    }
  }
}

extern "C" void qmpc_destroy_multi(QmpcMultiHandle* mh) {
  if (!mh) return;
  {
    std::lock_guard<std::mutex> lk(mh->mu);
    mh->stop = true;
  }
  mh->cv_go.notify_all();
  for (int g = 0; g < mh->n; ++g)
    if (mh->workers[g].joinable()) mh->workers[g].join();
  for (int g = 0; g < mh->n; ++g) {
    if (mh->h[g]) cudaSetDevice(mh->h[g]->device);
    if (mh->pin_in[g]) cudaFreeHost(mh->pin_in[g]);
    if (mh->pin_out[g]) cudaFreeHost(mh->pin_out[g]);
    qmpc_destroy(mh->h[g]);
  }
  delete mh;
}

extern "C" int qmpc_create_multi(const QmpcConfig* cfg, int32_t max_batch, const int32_t* devices, int32_t n_devices,
                                 QmpcMultiHandle** out) {
  if (!cfg || !out || !devices || max_batch < 1 || n_devices < 1 || n_devices > kMaxDevices) return QMPC_ERR_ARG;
  for (int a = 0; a < n_devices; ++a)
    for (int b = 0; b < a; ++b)
      if (devices[a] == devices[b]) return QMPC_ERR_ARG;
  QmpcMultiHandle* mh = new (std::nothrow) QmpcMultiHandle();
  if (!mh) return QMPC_ERR_ARG;
  *out = mh;   // returned even on failure: qmpc_multi_last_error / qmpc_destroy_multi stay usable
  mh->n = n_devices;
  mh->max_batch = max_batch;
  mh->convex = cfg->model == QMPC_MODEL_EULER_CONVEX;
  mh->in_sz = mh->convex ? sizeof(QmpcConvexProblem) : sizeof(QmpcProblem);
  mh->err[0] = 0;
  for (int g = 0; g < kMaxDevices; ++g) { mh->h[g] = nullptr; mh->pin_in[g] = nullptr; mh->pin_out[g] = nullptr; mh->job_rc[g] = 0; }
  for (int g = 0; g < n_devices; ++g) {
    const int cap = (max_batch + n_devices - 1) / n_devices;   // the largest shard of any batch <= max_batch
    mh->shard_cap[g] = cap;
    int rc = qmpc_create(cfg, cap, devices[g], &mh->h[g]);
    if (rc) {
      snprintf(mh->err, sizeof(mh->err), "device %d: %.200s", devices[g], mh->h[g] ? mh->h[g]->err : "qmpc_create failed");
      return rc;
    }
    if (cudaHostAlloc(&mh->pin_in[g], mh->in_sz * (size_t)cap, cudaHostAllocPortable) != cudaSuccess ||
        cudaHostAlloc((void**)&mh->pin_out[g], sizeof(QmpcResult) * (size_t)cap, cudaHostAllocPortable) != cudaSuccess) {
      snprintf(mh->err, sizeof(mh->err), "device %d: pinned staging allocation failed", devices[g]);
      return QMPC_ERR_CUDA;
    }
  }
  if (n_devices > 1)
    for (int g = 0; g < n_devices; ++g) mh->workers[g] = std::thread(multi_worker, mh, g);
  return QMPC_OK;
}

extern "C" int qmpc_solve_batch_host_multi(QmpcMultiHandle* mh, const void* in, int32_t batch, QmpcResult* out) {
  if (!mh) return QMPC_ERR_ARG;
  if (!in || !out || batch < 0) return QMPC_ERR_ARG;
  if (batch > mh->max_batch) return QMPC_ERR_CAPACITY;
  if (batch == 0) return QMPC_OK;
  for (int g = 0; g < mh->n; ++g)
    if (!mh->h[g] || !mh->h[g]->ws) return QMPC_ERR_CUDA;
  const bool direct = host_ptr_is_pinned(in) && host_ptr_is_pinned(out);
  if (mh->n == 1) return multi_run_shard(mh, 0, in, out, batch, direct, mh->err, sizeof(mh->err));
  {
    std::unique_lock<std::mutex> lk(mh->mu);
    mh->job_in = in; mh->job_out = out; mh->job_batch = batch; mh->job_direct = direct;
    mh->pending = mh->n;
    ++mh->generation;
    mh->cv_go.notify_all();
    mh->cv_done.wait(lk, [&] { return mh->pending == 0; });
  }
  int rc_all = QMPC_OK;
  for (int g = 0; g < mh->n; ++g)
    if (mh->job_rc[g]) { rc_all = mh->job_rc[g]; snprintf(mh->err, sizeof(mh->err), "%.200s", mh->job_err[g]); }
  return rc_all;
}

extern "C" int32_t qmpc_multi_device_count(const QmpcMultiHandle* mh) { return mh ? mh->n : 0; }
extern "C" int64_t qmpc_multi_launch_count(const QmpcMultiHandle* mh) {
  int64_t n = 0;
  if (mh) for (int g = 0; g < mh->n; ++g) n += mh->h[g] ? mh->h[g]->launches : 0;
  return n;
}
extern "C" const char* qmpc_multi_last_error(const QmpcMultiHandle* mh) { return mh ? mh->err : "null handle"; }
