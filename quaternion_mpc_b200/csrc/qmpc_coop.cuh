// qmpc_coop.cuh — kernel "coop": G (=16) lanes of a warp own one MPC problem; every per-problem
// matrix lives in shared memory, the gains / value functions / duals in an L2-resident per-slot
// scratch.  Same algorithm, decisions and (where it matters) operation order as the structured
// one-thread-per-problem kernel (qmpc_srb.cuh) and the CPU oracle.
//
// Why (profiles/r01_ncu_srb_thread_per_problem_B16384.txt): with one thread per problem the fp64
// working set (~13 KB of 12x12 temporaries per thread) spills to local memory, misses L1/L2 and the
// kernel runs at 3 % fp64-pipe utilisation on DRAM latency.  Here the working set per problem is
// ~10 KB of shared memory, shared by 16 lanes:
//   * the Riccati step is a sequence of "phases"; in each phase the lanes split the output
//     elements of one small block product (operands read from shared memory, mostly broadcast or
//     conflict-free), separated by __syncwarp on the half-warp;
//   * the back-tracking line search is evaluated speculatively: lane l rolls out step length
//     2^-l (all 16 trial steps at once, K_k/d_k read as broadcasts), the first lane that passes the
//     Armijo test wins — identical result to the sequential search, ~1 roll-out of latency instead
//     of ~5;
//   * linearisation, stationarity residuals, dual update, Riccati duals are parallel over knots.
// The kernel is persistent: grid = resident slots, each slot strides over the batch.
//
// The body is written with COOP_PHASE / COOP_SYNC so that the very same source runs on the host
// (tests/emul, lanes executed one after another) for GPU-less debugging.
#pragma once
#include "qmpc_srb.cuh"

#ifdef __CUDA_ARCH__
#define COOP_PHASE for (int lane = lane_id, once_ = 1; once_; once_ = 0)
#define COOP_SYNC() __syncwarp(lane_mask)
#else
#define COOP_PHASE for (int lane = 0; lane < G; ++lane)
#define COOP_SYNC() ((void)0)
#endif

// Block-level phase alignment (default; -DQMPC_COOP_NO_BLOCK_SYNC disables): every thread of the block
// passes one barrier per AL-iLQR iteration - a uniform iterations_max times per problem wave, finished
// or idle slots included - so that the block's warps walk through the same code region together and
// share instruction-cache lines instead of thrashing the 32 KB L1.5 with eight different phases
// (measured +10 % with 128-thread blocks; a second barrier per iteration or one per knot adds nothing).
#if !defined(QMPC_COOP_NO_BLOCK_SYNC) && !defined(QMPC_COOP_BLOCK_SYNC)
#define QMPC_COOP_BLOCK_SYNC
#endif
#if defined(__CUDA_ARCH__) && defined(QMPC_COOP_BLOCK_SYNC)
// non-aligned barrier: the two problems of a warp may arrive from different code paths (one finished,
// one still iterating), which __syncthreads() / barrier.sync.aligned does not allow
#define COOP_BLOCK_SYNC() asm volatile("barrier.sync 1, %0;" ::"r"(blockDim.x) : "memory")
#define COOP_BLOCK_SYNC_MID() ((void)0)
#define COOP_KNOT_SYNC() ((void)0)
#define COOP_KNOT_SYNC_ALL(N_) ((void)0)
#else
#define COOP_BLOCK_SYNC() ((void)0)
#define COOP_BLOCK_SYNC_MID() ((void)0)
#define COOP_KNOT_SYNC() ((void)0)
#define COOP_KNOT_SYNC_ALL(N_) ((void)0)
#endif

#ifdef __CUDACC__
#define QMPC_NOINLINE __noinline__
#else
#define QMPC_NOINLINE
#endif

namespace qmpc {

// access points of the two classes of scratch traffic: trial trajectories (written 16x per iteration, read
// once) and gains / value functions (written once per backward pass, read once per forward pass). Cache
// hints were measured on both (evict-first: -6 %, evict_last: +0.7 %, profiles/r01_session3_experiments.md)
// and dropped; the accessors stay so that the two streams remain distinguishable in the source.
QMPC_HD inline void st_stream(double* p, double v) { *p = v; }
QMPC_HD inline double ld_stream(const double* p) { return *p; }
QMPC_HD inline void st_keep(double* p, double v) { *p = v; }
QMPC_HD inline double ld_keep(const double* p) { return *p; }

QMPC_HD inline double qmpc_rsqrt(double x) {
#ifdef __CUDA_ARCH__
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

// ---- 3x3 block kernels used by the block-per-lane phases (lane = (block row, block col)) --------
// The row loop is deliberately NOT unrolled: the kernel is instruction-fetch bound, compact code wins.
// Operands that are reused across the rows are read into registers ONCE, before the first store:
// source and destination are plain pointers into the same shared-memory pool, so the compiler must
// otherwise assume every store clobbers them and reload (measured: 7 LDS per 4 flops; now ~3).
// dst = X(:, 3:6) * Mt + beta * X(:, 9:12)   X: 3 rows of a row-major matrix with leading dim ld
QMPC_HD inline void blk_right(const double* X, int ld, const double* Mt, double beta, double* dst, int ldd) {
  double m[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = Mt[i];
#pragma unroll 1
  for (int a = 0; a < 3; ++a) {
    const double* Xa = X + ld * a;
    const double x3 = Xa[3], x4 = Xa[4], x5 = Xa[5], y0 = Xa[9], y1 = Xa[10], y2 = Xa[11];
    const double r0 = x3 * m[0] + x4 * m[3] + x5 * m[6] + beta * y0;
    const double r1 = x3 * m[1] + x4 * m[4] + x5 * m[7] + beta * y1;
    const double r2 = x3 * m[2] + x4 * m[5] + x5 * m[8] + beta * y2;
    dst[ldd * a] = r0; dst[ldd * a + 1] = r1; dst[ldd * a + 2] = r2;
  }
}
// dst = alpha * X(:, 0:3) + beta * X(:, 6:9)
QMPC_HD inline void blk_even(const double* X, int ld, double alpha, double beta, double* dst, int ldd) {
#pragma unroll 1
  for (int a = 0; a < 3; ++a) {
    const double* Xa = X + ld * a;
    const double x0 = Xa[0], x1 = Xa[1], x2 = Xa[2], y0 = Xa[6], y1 = Xa[7], y2 = Xa[8];
    dst[ldd * a] = alpha * x0 + beta * y0;
    dst[ldd * a + 1] = alpha * x1 + beta * y1;
    dst[ldd * a + 2] = alpha * x2 + beta * y2;
  }
}
// dst = Mt^T * Y(3:6, :) + beta * Y(9:12, :)   Y: 3 columns (starting at Y) of a row-major matrix
QMPC_HD inline void blk_left(const double* Y, int ld, const double* Mt, double beta, double* dst, int ldd) {
  double m[9], y[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = Mt[i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int b = 0; b < 3; ++b) y[3 * i + b] = Y[ld * (3 + i) + b];
#pragma unroll   // unrolled: m[a] must stay a compile-time register index
  for (int a = 0; a < 3; ++a) {
    const double* Ya = Y + ld * (9 + a);
    const double z0 = Ya[0], z1 = Ya[1], z2 = Ya[2];
    const double m0 = m[a], m1 = m[3 + a], m2 = m[6 + a];
    const double r0 = m0 * y[0] + m1 * y[3] + m2 * y[6] + beta * z0;
    const double r1 = m0 * y[1] + m1 * y[4] + m2 * y[7] + beta * z1;
    const double r2 = m0 * y[2] + m1 * y[5] + m2 * y[8] + beta * z2;
    dst[ldd * a] = r0; dst[ldd * a + 1] = r1; dst[ldd * a + 2] = r2;
  }
}
// dst = alpha * Y(0:3, :) + beta * Y(6:9, :)
QMPC_HD inline void blk_evenT(const double* Y, int ld, double alpha, double beta, double* dst, int ldd) {
#pragma unroll 1
  for (int a = 0; a < 3; ++a) {
    const double x0 = Y[ld * a], x1 = Y[ld * a + 1], x2 = Y[ld * a + 2];
    const double y0 = Y[ld * (6 + a)], y1 = Y[ld * (6 + a) + 1], y2 = Y[ld * (6 + a) + 2];
    dst[ldd * a] = alpha * x0 + beta * y0;
    dst[ldd * a + 1] = alpha * x1 + beta * y1;
    dst[ldd * a + 2] = alpha * x2 + beta * y2;
  }
}
// dst = s * T(0:3, :) + Mt^T * T(3:6, :)   (rows of T have leading dim ld); W^T-type products of phases D / E
QMPC_HD inline void blk_wt(const double* T, int ld, double s, const double* Mt, double* dst, int ldd,
                           const double* add /* nullable 3x3 row-major, added to the result */) {
  double m[9], y[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = Mt[i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int b = 0; b < 3; ++b) y[3 * i + b] = T[ld * (3 + i) + b];
#pragma unroll   // unrolled: m[a] must stay a compile-time register index
  for (int a = 0; a < 3; ++a) {
    const double t0 = T[ld * a], t1 = T[ld * a + 1], t2 = T[ld * a + 2];
    const double m0 = m[a], m1 = m[3 + a], m2 = m[6 + a];
    double r0 = s * t0 + m0 * y[0] + m1 * y[3] + m2 * y[6];
    double r1 = s * t1 + m0 * y[1] + m1 * y[4] + m2 * y[7];
    double r2 = s * t2 + m0 * y[2] + m1 * y[5] + m2 * y[8];
    if (add) { r0 += add[3 * a]; r1 += add[3 * a + 1]; r2 += add[3 * a + 2]; }
    dst[ldd * a] = r0; dst[ldd * a + 1] = r1; dst[ldd * a + 2] = r2;
  }
}
// dst = s * S(:, 0:3) + S(:, 3:6) * Mt   (3 rows of S with leading dim ld): the S W product of phase D
QMPC_HD inline void blk_w(const double* S, int ld, double s, const double* Mt, double* dst, int ldd) {
  double m[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = Mt[i];
#pragma unroll 1
  for (int a = 0; a < 3; ++a) {
    const double* Sa = S + ld * a;
    const double t0 = Sa[0], t1 = Sa[1], t2 = Sa[2], x3 = Sa[3], x4 = Sa[4], x5 = Sa[5];
    dst[ldd * a] = s * t0 + x3 * m[0] + x4 * m[3] + x5 * m[6];
    dst[ldd * a + 1] = s * t1 + x3 * m[1] + x4 * m[4] + x5 * m[7];
    dst[ldd * a + 2] = s * t2 + x3 * m[2] + x4 * m[5] + x5 * m[8];
  }
}
QMPC_HD inline void blk_store_keep(double* dst, int ld, const double* v) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) st_keep(dst + ld * a + b, v[3 * a + b]);
}
QMPC_HD inline void blk_store(double* dst, int ld, const double* v) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) dst[ld * a + b] = v[3 * a + b];
}

constexpr int kCoopBlockShared = 26;   // doubles at the head of the block's shared memory: q[13], r[12], pad

template <int V>
struct IntTag { static constexpr int value = V; };

template <int NF, int G>
struct CoopLayout {
  static constexpr int NU = 3 * NF, NC = 6 * NF;
  static constexpr int kModel = (int)((sizeof(QuatModel<NF>) + 7) / 8);
  // ---- shared memory (doubles) per problem
  static constexpr int kVec = 156;   // the register/shuffle Cholesky needs no column-exchange buffer (cv::tcol)
  QMPC_HD static int sX(int N) { return kModel; }
  QMPC_HD static int sU(int N) { return sX(N) + (N + 1) * 13; }
  QMPC_HD static int sP(int N) { return (sU(N) + N * NU + 1) / 2 * 2; }   // 16-byte aligned: cp.async target
  QMPC_HD static int sPA(int N) { return sP(N) + 144; }    // PA, later Quu / its Cholesky factor
  QMPC_HD static int sT(int N) { return sPA(N) + 144; }
  QMPC_HD static int sPM(int N) { return sT(N) + 72; }     // PM, later SW
  QMPC_HD static int sS(int N) { return sPM(N) + 72; }
  QMPC_HD static int sQux(int N) { return sS(N) + 36; }    // Qux, later V = L^-1 Qux
  QMPC_HD static int sVec(int N) { return sQux(N) + NU * 12; }
  QMPC_HD static int sLin(int N) { return sVec(N) + kVec; }
  // Optional residents (`flags`): bit 0 = the per-knot linearisation blocks (27 N), bit 1 = the duals
  // (NC N) also live in shared memory - used when they do not cost residency (short horizons);
  // otherwise they stay in the L2-resident scratch.  Same code either way, only the pointers differ.
  QMPC_HD static int sLinAll(int N) { return (sLin(N) + 27 + 1) / 2 * 2; }
  QMPC_HD static int sMu(int N, int flags) { return sLinAll(N) + ((flags & 1) ? 27 * N : 0); }
  QMPC_HD static int smem_doubles(int N, int flags) { return (sMu(N, flags) + ((flags & 2) ? NC * N : 0) + 1) / 2 * 2; }
  // ---- global scratch (doubles) per slot
  QMPC_HD static size_t gK(int N) { return 0; }
  QMPC_HD static size_t gd(int N) { return gK(N) + (size_t)N * NU * 12; }
  QMPC_HD static size_t gP(int N) { return gd(N) + (size_t)N * NU; }
  QMPC_HD static size_t gpv(int N) { return gP(N) + (size_t)(N + 1) * 144; }
  QMPC_HD static size_t gmu(int N) { return gpv(N) + (size_t)(N + 1) * 12; }
  QMPC_HD static size_t glin(int N) { return gmu(N) + (size_t)N * NC; }
  QMPC_HD static size_t gDX(int N) { return glin(N) + (size_t)N * 27; }
  // trial trajectories of the speculative line search, [element][lane] so the 16 lanes store coalesced
  // cost expansion of every knot (gradient 12 + attitude Hessian block 9), written once per iteration
  QMPC_HD static size_t gLX(int N) { return gDX(N) + (size_t)(N + 1) * 12; }
  QMPC_HD static size_t gTX(int N) { return (gLX(N) + (size_t)(N + 1) * 21 + 15) / 16 * 16; }
  QMPC_HD static size_t gTU(int N) { return gTX(N) + (size_t)(N + 1) * 13 * G; }
  QMPC_HD static size_t scratch_doubles(int N) { return (gTU(N) + (size_t)N * NU * G + 15) / 16 * 16; }
};

// offsets inside the shared "vec" block
namespace cv {
constexpr int lx = 0, Qx = 12, Qu = 24, s = 36, Atp = 42, g = 54, Dblk = 66, Hphi = 102, vu = 111, pv = 123,
              scal = 135, rdiag = 144, tcol = 156;  // scal[0]=dphi0 [1]=hphi [2]=phi [3]=viol
}

// stage cost + AL terms of one knot: same accumulation order as stage_cost() / merit() in
// qmpc_dense.cuh, but rolled per foot - it runs once per dual update, compact code matters more
template <class M>
QMPC_HD inline void knot_merit(const M& m, const QmpcConfig& cfg, const double* wr, int k, int N, const double* x,
                               const double* u, const double* mu_k, double rho, double& J, double& viol) {
  constexpr int NF = M::NU / 3;
  double xr[M::NX], Jl = 0;
  m.xref(k, xr);
#pragma unroll
  for (int i = 0; i < M::NX; ++i) { const double dxi = x[i] - xr[i]; Jl += 0.5 * cfg.q_weights[i] * dxi * dxi; }
  if (k < N) {
#pragma unroll 1
    for (int f = 0; f < NF; ++f) {
      const double d0 = u[3 * f], d1 = u[3 * f + 1], d2 = u[3 * f + 2] - m.urefz(k, f);
      Jl += 0.5 * wr[3 * f] * d0 * d0;
      Jl += 0.5 * wr[3 * f + 1] * d1 * d1;
      Jl += 0.5 * wr[3 * f + 2] * d2 * d2;
    }
  }
  if (cfg.w != 0.0) {
    const double s = xr[3] * x[3] + xr[4] * x[4] + xr[5] * x[5] + xr[6] * x[6];
    Jl += cfg.w * (1.0 - fabs(s));
  }
  J += Jl;
  if (k < N) {
    double acc = 0;
#pragma unroll 1
    for (int f = 0; f < NF; ++f) {
      const double u0 = u[3 * f], u1 = u[3 * f + 1], u2 = u[3 * f + 2];
      const double fzc_f = m.fzc(k, f);
#pragma unroll 1
      for (int r = 0; r < 6; ++r) {
        double c = m.CR[3 * r] * u0 + m.CR[3 * r + 1] * u1 + m.CR[3 * r + 2] * u2;
        if (r == 4) c += -fzc_f;
        const double mui = mu_k[6 * f + r];
        const double est = mui + rho * c;
        const double lh = est > 0 ? est : 0;
        if (c > viol) viol = c;
        acc += lh * lh - mui * mui;
      }
    }
    J += acc / (2 * rho);
  }
}

// attitude block of the cost Hessian: G^T diag(Qq) G + hphi I
QMPC_HD inline void hphi_block(const QmpcConfig& cfg, const double* x, double hphi, double* H) {
  double Gq[12];
  quat_G(x + 3, Gq);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double s = 0;
      for (int i = 0; i < 4; ++i) s += Gq[3 * i + a] * cfg.q_weights[3 + i] * Gq[3 * i + b];
      H[3 * a + b] = s + (a == b ? hphi : 0.0);
    }
}

// 3x3 block (br, bc) of the cost Hessian in error coordinates
QMPC_HD inline void lxx_block(const double* wq, const double* Hphi, int br, int bc, double* out) {
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = 0.0;
  if (br != bc) return;
  if (br == 1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) out[i] = Hphi[i];
  } else {
    const int q0 = br == 0 ? 0 : (br == 2 ? 7 : 10);
    out[0] = wq[q0]; out[4] = wq[q0 + 1]; out[8] = wq[q0 + 2];
  }
}

// One roll-out of the whole horizon by ONE lane (kept out of line: it is used by the nominal
// roll-out, by the 16 speculative line-search lanes and by the accepted step, and inlining it three
// times is what pushed the SASS far past the instruction cache).
//   mode 0: open loop, u = u_ref: writes the nominal X, U; returns merit / violation
//   mode 1: trial step `alpha` around (X, U) with gains (gK, gd): X, U untouched; the trial
//           trajectory is recorded in the scratch (gTX/gTU, element-major, lane `tl` of `tstride`) so
//           that the accepted one is simply copied back - no second roll-out; returns merit / violation
//   mode 2: (unused by the kernel, kept for the host emulation tests) accepted step in place
template <int NF>
QMPC_HD QMPC_NOINLINE void coop_rollout(const QuatModel<NF>& m, const QmpcConfig& cfg, const double* wr, int N, float h, double* X,
                                        double* U, double* DX, const double* gK, const double* gd,
                                        const double* gmu, double rho, double alpha, int mode, double* Jout,
                                        double* violout, double* gTX, double* gTU, int tl, int tstride,
                                        double* kstage, unsigned lane_mask, const QmpcWarmStart* winit) {
  // Compact by construction (instruction-fetch bound otherwise): the input never
  // exists as an array - each foot's force is formed, costed, cone-checked and folded into the net
  // wrench inside one 4-trip loop; the wrench drives both midpoint evaluations.  Accumulation
  // orders are exactly those of stage_cost() / knot_merit() / ct_dyn() / mid_dyn().
  using M = QuatModel<NF>;
  constexpr int NX = 13, NE = 12, NU = M::NU, NC = M::NC;
  const double hd = (double)h, hh = (double)(h / 2);
  double x[NX], J = 0, vl = 0;
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = X[i];
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_KSTAGE)
  // Mode 1 runs on all `tstride` lanes of the problem in lock-step: the gain matrix of knot k+1 is
  // copied (cp.async, 16 bytes per request, the lanes split the rows) from the L2-resident scratch
  // into a double buffer in shared memory while knot k is being computed, so the 72 broadcast reads
  // of K_k per lane are shared-memory reads instead of L2 round trips.  The feed-forward d_k rides along.
  constexpr int kChunksK = NU * 12 / 2, kChunks = kChunksK + NU / 2, kStage = NU * 12 + NU;
  auto stage_gain = [&](int k) {
    const double* srcK = gK + (size_t)k * NU * 12;
    const double* srcd = gd + (size_t)k * NU - 2 * kChunksK;
    double* dst = kstage + (k & 1) * kStage;
    for (int c = tl; c < kChunks; c += tstride) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + 2 * c);
      const double* src = (c < kChunksK ? srcK : srcd) + 2 * c;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (mode == 1) stage_gain(0);
#else
  (void)kstage; (void)lane_mask;
#endif
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    double dx[NE];
    if (mode != 0) state_diff<M>(x, X + k * NX, dx);
    if (mode == 1) {
#pragma unroll
      for (int i = 0; i < NX; ++i) st_stream(gTX + (size_t)(k * NX + i) * tstride + tl, x[i]);
    }
    if (mode == 2) {
#pragma unroll
      for (int i = 0; i < NE; ++i) DX[k * NE + i] = dx[i];
#pragma unroll
      for (int i = 0; i < NX; ++i) X[k * NX + i] = x[i];
    }
    // ---- state part of the stage cost
    double Jl = 0, sq = 0;
    if (mode != 2) {
      double xr[NX];
      m.xref(k, xr);
#pragma unroll
      for (int i = 0; i < NX; ++i) { const double dxi = x[i] - xr[i]; Jl += 0.5 * cfg.q_weights[i] * dxi * dxi; }
      sq = xr[3] * x[3] + xr[4] * x[4] + xr[5] * x[5] + xr[6] * x[6];
    }
    if (k == N) {
      if (mode != 2) {
        if (cfg.w != 0.0) Jl += cfg.w * (1.0 - fabs(sq));
        J += Jl;
      }
      break;
    }
    // ---- per foot: force, input cost, cone rows / AL merit, wrench
    double mom0 = 0, mom1 = 0, mom2 = 0, fs0 = 0, fs1 = 0, fs2 = 0, acc = 0;
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_KSTAGE)
    const double* Kk = mode == 1 ? kstage + (k & 1) * kStage : gK;
    const double* dk = mode == 1 ? Kk + NU * 12 : gd;
    if (mode == 1) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp(lane_mask);              // K_k visible to all lanes; everyone is done with K_{k-1}
      if (k + 1 < N) stage_gain(k + 1);
    }
#else
    const double* Kk = gK + (size_t)k * NU * 12;
    const double* dk = gd + k * NU;
#endif
#ifndef QMPC_COOP_FOOT_UNROLL
#define QMPC_COOP_FOOT_UNROLL 1
#endif
    constexpr int kFootUnroll = QMPC_COOP_FOOT_UNROLL;   // 1 = compact (instruction cache), NF = all feet's chains interleaved
#pragma unroll(kFootUnroll)
    for (int f = 0; f < NF; ++f) {
      double u0, u1, u2;
      if (mode == 0) {
        if (winit) {   // warm start: previous solution shifted by one knot
          const double* wrow = warm_row(winit, k, N) + 3 * f;
          u0 = wrow[0]; u1 = wrow[1]; u2 = wrow[2];
        } else {
          u0 = 0.0; u1 = 0.0; u2 = m.urefz(0, f);   // SetInput(u_traj_ref.at(0)), QuatMpc.cpp:253
        }
      } else {
        double t0 = 0, t1 = 0, t2 = 0;
        const double* K0 = Kk + (3 * f) * 12;
#if !defined(QMPC_COOP_NO_K128) && defined(__CUDA_ARCH__)
        {   // three 96-byte gain rows as 18 x 16-byte loads (rows are 16-byte aligned in the scratch)
#ifndef QMPC_COOP_NO_KSTAGE
          const unsigned ks = (unsigned)__cvta_generic_to_shared(K0);
          auto ldk = [&](int l) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ks + 16u * l)); return v; };
#else
          const double2* K2 = reinterpret_cast<const double2*>(K0);
          auto ldk = [&](int l) { return K2[l]; };
#endif
#pragma unroll
          for (int l = 0; l < 6; ++l) { const double2 v = ldk(l); t0 += v.x * dx[2 * l]; t0 += v.y * dx[2 * l + 1]; }
#pragma unroll
          for (int l = 0; l < 6; ++l) { const double2 v = ldk(6 + l); t1 += v.x * dx[2 * l]; t1 += v.y * dx[2 * l + 1]; }
#pragma unroll
          for (int l = 0; l < 6; ++l) { const double2 v = ldk(12 + l); t2 += v.x * dx[2 * l]; t2 += v.y * dx[2 * l + 1]; }
        }
#else
#pragma unroll
        for (int l = 0; l < NE; ++l) t0 += K0[l] * dx[l];
#pragma unroll
        for (int l = 0; l < NE; ++l) t1 += K0[12 + l] * dx[l];
#pragma unroll
        for (int l = 0; l < NE; ++l) t2 += K0[24 + l] * dx[l];
#endif
        u0 = U[k * NU + 3 * f] + alpha * dk[3 * f] + t0;
        u1 = U[k * NU + 3 * f + 1] + alpha * dk[3 * f + 1] + t1;
        u2 = U[k * NU + 3 * f + 2] + alpha * dk[3 * f + 2] + t2;
      }
      if (mode != 1) { U[k * NU + 3 * f] = u0; U[k * NU + 3 * f + 1] = u1; U[k * NU + 3 * f + 2] = u2; }
      else {
        double* tu = gTU + (size_t)(k * NU + 3 * f) * tstride + tl;
        st_stream(tu, u0); st_stream(tu + tstride, u1); st_stream(tu + 2 * tstride, u2);
      }
      if (mode != 2) {
        const double d0 = u0, d1 = u1, d2 = u2 - m.urefz(k, f);   // u_ref = (0, 0, weight share)
        Jl += 0.5 * wr[3 * f] * d0 * d0;
        Jl += 0.5 * wr[3 * f + 1] * d1 * d1;
        Jl += 0.5 * wr[3 * f + 2] * d2 * d2;
        const double* mu_f = gmu + k * NC + 6 * f;
        const double fzc_f = m.fzc(k, f);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          double c = m.CR[3 * r] * u0 + m.CR[3 * r + 1] * u1 + m.CR[3 * r + 2] * u2;
          if (r == 4) c += -fzc_f;
          const double mui = mu_f[r];
          const double est = mui + rho * c;
          const double lh = est > 0 ? est : 0;
          if (c > vl) vl = c;
          acc += lh * lh - mui * mui;
        }
      }
      const double* rf = m.foot + 3 * f;
      const double c0 = rf[1] * u2 - rf[2] * u1, c1 = rf[2] * u0 - rf[0] * u2, c2 = rf[0] * u1 - rf[1] * u0;
      mom0 += c0; fs0 += u0;
      mom1 += c1; fs1 += u1;
      mom2 += c2; fs2 += u2;
    }
    if (mode != 2) {
      if (cfg.w != 0.0) Jl += cfg.w * (1.0 - fabs(sq));
      J += Jl;
      J += acc / (2 * rho);
    }
    // ---- explicit midpoint step driven by the net wrench (AltroUtils.cpp:9-22, 383-391)
    mom0 += m.tau_g[0]; mom1 += m.tau_g[1]; mom2 += m.tau_g[2];
    const double al0 = fs0 * m.inv_mass + m.g[0], al1 = fs1 * m.inv_mass + m.g[1], al2 = fs2 * m.inv_mass + m.g[2];
    const double aw0 = m.Iinv[0] * mom0 + m.Iinv[1] * mom1 + m.Iinv[2] * mom2;
    const double aw1 = m.Iinv[3] * mom0 + m.Iinv[4] * mom1 + m.Iinv[5] * mom2;
    const double aw2 = m.Iinv[6] * mom0 + m.Iinv[7] * mom1 + m.Iinv[8] * mom2;
    double xm[NX];
    {
      const double *q = x + 3, *w = x + 10;
      xm[0] = x[7] * hh + x[0]; xm[1] = x[8] * hh + x[1]; xm[2] = x[9] * hh + x[2];
      xm[3] = (0.5 * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2])) * hh + q[0];
      xm[4] = (0.5 * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2])) * hh + q[1];
      xm[5] = (0.5 * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2])) * hh + q[2];
      xm[6] = (0.5 * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2])) * hh + q[3];
      xm[7] = al0 * hh + x[7]; xm[8] = al1 * hh + x[8]; xm[9] = al2 * hh + x[9];
      xm[10] = aw0 * hh + x[10]; xm[11] = aw1 * hh + x[11]; xm[12] = aw2 * hh + x[12];
    }
    {
      const double *q = xm + 3, *w = xm + 10;
      const double qd0 = 0.5 * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2]);
      const double qd1 = 0.5 * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2]);
      const double qd2 = 0.5 * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2]);
      const double qd3 = 0.5 * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2]);
      x[0] = x[0] + hd * xm[7]; x[1] = x[1] + hd * xm[8]; x[2] = x[2] + hd * xm[9];
      x[3] = x[3] + hd * qd0; x[4] = x[4] + hd * qd1; x[5] = x[5] + hd * qd2; x[6] = x[6] + hd * qd3;
      x[7] = x[7] + hd * al0; x[8] = x[8] + hd * al1; x[9] = x[9] + hd * al2;
      x[10] = x[10] + hd * aw0; x[11] = x[11] + hd * aw1; x[12] = x[12] + hd * aw2;
    }
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) X[(k + 1) * NX + i] = x[i];
    }
  }
  *Jout = J;
  *violout = vl;
}

template <int NF, int G>
QMPC_HD void coop_solve_one(const QmpcConfig& cfg, const SolverOpts& o, const QmpcProblem* in,
                            const unsigned char* sched, QmpcWarmStart* warm, QmpcResult* out,
                            int pid, double* sm, double* gs, int lane_id, unsigned lane_mask, int flags,
                            const double* wts) {
  using M = QuatModel<NF>;
  using L = CoopLayout<NF, G>;
  constexpr int NX = 13, NE = 12, NU = M::NU, NC = M::NC;
  (void)lane_id; (void)lane_mask;
  // weights with run-time indices come from `wts` (q[13], r[12]): a block-shared copy in shared memory on
  // the device - an indexed read of the kernel parameter bank is an LDC that stalls like a global load
  const double* wq = wts;
  const double* wr = wts + 13;
  const int N = o.N;
  const float h = o.h;
  const double hd = (double)h, hh = (double)(h / 2), c1 = hd * hh;

  M& m = *reinterpret_cast<M*>(sm);
  double* X = sm + L::sX(N);
  double* U = sm + L::sU(N);
  double* DX = gs + L::gDX(N);
  double* gLX = gs + L::gLX(N);
  double* gTX = gs + L::gTX(N);
  double* gTU = gs + L::gTU(N);
  double* P = sm + L::sP(N);
  double* PA = sm + L::sPA(N);
    double* T = sm + L::sT(N);
  double* PM = sm + L::sPM(N);
  double* SW = PM;
  double* S = sm + L::sS(N);
  double* Qux = sm + L::sQux(N);
  double* vec = sm + L::sVec(N);
  // 2 G reduction slots at the tail of T (only used outside the backward pass, where T is dead); the
  // roll-out's gain stage (2 x (NU * 12 + NU) doubles) occupies P, PA and the head of T meanwhile
  double* red = T + 72 - 2 * G;
  static_assert(2 * (NU * 12 + NU) <= 288 + 72 - 2 * G, "gain stage overlaps the reduction slots");
  double* lin = sm + L::sLin(N);
  double* gK = gs + L::gK(N);
  double* gd = gs + L::gd(N);
  double* gP = gs + L::gP(N);
  double* gpv = gs + L::gpv(N);
  double* gmu = (flags & 2) ? sm + L::sMu(N, flags) : gs + L::gmu(N);
  double* glin = (flags & 1) ? sm + L::sLinAll(N) : gs + L::glin(N);
  double* scal = vec + cv::scal;

  // ------------------------------------------------------------------ set-up + nominal roll-out
  COOP_PHASE {
#pragma unroll 1
    for (int i = lane; i < N * NC; i += G) gmu[i] = 0.0;
    if (lane == 0) {
      QmpcProblem prob = in[pid];
      m.setup(cfg, prob, sched ? sched + (size_t)pid * QMPC_MAX_HORIZON : nullptr, X);
    }
  }
  COOP_SYNC();
  double rho = o.penalty_initial;
  constexpr int NCAND = G;
  COOP_PHASE {
    if (lane == 0) coop_rollout<NF>(m, cfg, wr, N, h, X, U, DX, gK, gd, gmu, rho, 0.0, 0, &scal[2], &scal[3], gTX, gTU, 0, G, P, lane_mask,
                                     (warm && warm[pid].valid) ? warm + pid : nullptr);
  }
  COOP_SYNC();
  double phi = scal[2], viol = scal[3];
  int status = QMPC_STATUS_MAX_ITERATIONS, iters = 0;
  double cost_decrease = INFINITY;
  if (!isfinite(phi)) status = QMPC_STATUS_NONFINITE;

#if defined(__CUDA_ARCH__) && defined(QMPC_COOP_BLOCK_SYNC)
#define COOP_ITER_COND(it_) ((it_) < o.iterations_max)
#define COOP_ITER_LEAVE continue
#else
#define COOP_ITER_COND(it_) ((it_) < o.iterations_max && status == QMPC_STATUS_MAX_ITERATIONS)
#define COOP_ITER_LEAVE break
#endif
#pragma unroll 1
  for (int it = 0; COOP_ITER_COND(it); ++it) {
    COOP_BLOCK_SYNC();
#if defined(__CUDA_ARCH__) && defined(QMPC_COOP_BLOCK_SYNC)
    if (status != QMPC_STATUS_MAX_ITERATIONS) { COOP_KNOT_SYNC_ALL(N); COOP_BLOCK_SYNC_MID(); continue; }   // finished: keep passing the barriers
#endif
    // ---------------- expansions, lane k <- knot k: cost gradient + attitude Hessian block (21 doubles) and
    // the dynamics blocks (27 doubles).  The cost expansion used to be recomputed by single lanes inside
    // the backward pass (on its critical path, and 4 KB of code in its loop) and again for the
    // stationarity test; it is the same X throughout the iteration.
    COOP_PHASE {
#pragma unroll 1
      for (int k = lane; k <= N; k += G) {
        double lx[NE], Hk[9], hphi;
        cost_expand(m, cfg, k, X + k * NX, lx, &hphi);
        hphi_block(cfg, X + k * NX, hphi, Hk);
        for (int i = 0; i < NE; ++i) gLX[k * 21 + i] = lx[i];
        for (int i = 0; i < 9; ++i) gLX[k * 21 + 12 + i] = Hk[i];
        if (k < N) {
          KnotLin Lk;
          srb_linearize(m, X + k * NX, U + k * NU, X + (k + 1) * NX, hd, hh, Lk);
          for (int i = 0; i < 9; ++i) {
            glin[k * 27 + i] = Lk.Aff[i];
            glin[k * 27 + 9 + i] = Lk.Afw[i];
            glin[k * 27 + 18 + i] = Lk.Cf[i];
          }
        }
      }
    }
    COOP_SYNC();

    if (it > 0) {
      // ---------------- stationarity with the Riccati duals of the accepted step (DX holds y_k)
      COOP_PHASE {
        double rx = 0, ru = 0;
#pragma unroll 1
        for (int k = lane; k <= N; k += G) {
          double lx[NE];
          for (int a = 0; a < NE; ++a) lx[a] = gLX[k * 21 + a];
          if (k == N) {
            for (int a = 0; a < NE; ++a) {
              double v = fabs(lx[a] - DX[N * NE + a]);
              if (v > rx) rx = v;
            }
          } else {
            KnotLin Lk;
            for (int i = 0; i < 9; ++i) {
              Lk.Aff[i] = glin[k * 27 + i];
              Lk.Afw[i] = glin[k * 27 + 9 + i];
              Lk.Cf[i] = glin[k * 27 + 18 + i];
            }
            const double* u = U + k * NU;
            const double* yn = DX + (k + 1) * NE;
            double Aty[NE], t6[6];
            srb_At_vec(Lk, hd, yn, Aty);
            srb_Mt_vec(Lk, hd, hh, yn, t6);
            for (int a = 0; a < NE; ++a) {
              double v = fabs(lx[a] + Aty[a] - DX[k * NE + a]);
              if (v > rx) rx = v;
            }
            // input residual per foot, rolled (once per iteration: compact code beats unrolled speed):
            // R (u - u_ref) + J^T max(0, mu + rho c) + W^T M^T y   (same accumulation order as al_terms)
#pragma unroll 1
            for (int f = 0; f < NF; ++f) {
              const double* uf = u + 3 * f;
              const double* IS = m.IS + 9 * f;
              const double fzc_f = m.fzc(k, f);
              double g0 = 0, g1 = 0, g2 = 0;
#pragma unroll 1
              for (int r = 0; r < 6; ++r) {
                const double j0 = m.CR[3 * r], j1 = m.CR[3 * r + 1], j2 = m.CR[3 * r + 2];
                double c = j0 * uf[0] + j1 * uf[1] + j2 * uf[2];
                if (r == 4) c += -fzc_f;
                const double est = gmu[k * NC + 6 * f + r] + rho * c;
                if (est > 0) { g0 += j0 * est; g1 += j1 * est; g2 += j2 * est; }
              }
              const double b0 = m.inv_mass * t6[0] + IS[0] * t6[3] + IS[3] * t6[4] + IS[6] * t6[5];
              const double b1 = m.inv_mass * t6[1] + IS[1] * t6[3] + IS[4] * t6[4] + IS[7] * t6[5];
              const double b2 = m.inv_mass * t6[2] + IS[2] * t6[3] + IS[5] * t6[4] + IS[8] * t6[5];
              const double v0 = fabs(wr[3 * f] * uf[0] + g0 + b0);
              const double v1 = fabs(wr[3 * f + 1] * uf[1] + g1 + b1);
              const double v2 = fabs(wr[3 * f + 2] * (uf[2] - m.urefz(k, f)) + g2 + b2);
              if (v0 > ru) ru = v0;
              if (v1 > ru) ru = v1;
              if (v2 > ru) ru = v2;
            }
          }
        }
        red[lane] = rx > ru ? rx : ru;
      }
      COOP_SYNC();
      double stat = 0;
      for (int l = 0; l < G; ++l) stat = red[l] > stat ? red[l] : stat;
      COOP_SYNC();
      if (stat < o.tol_stationarity && viol < o.tol_primal_feasibility) {
        status = QMPC_STATUS_SUCCESS;
        COOP_KNOT_SYNC_ALL(N);
        COOP_BLOCK_SYNC_MID();
        COOP_ITER_LEAVE;
      }
      if (fabs(cost_decrease) < o.tol_cost_intermediate || stat < o.tol_stationarity) {
        // dual update (row-parallel), penalty update, merit refresh (knot-parallel)
        COOP_PHASE {
#pragma unroll 1
          for (int idx = lane; idx < N * NC; idx += G) {
            const int k = idx / NC, r = idx % NC, f = r / 6, rr = r % 6;
            const double* u = U + k * NU + 3 * f;
            double c = m.CR[3 * rr] * u[0] + m.CR[3 * rr + 1] * u[1] + m.CR[3 * rr + 2] * u[2];
            if (rr == 4) c += -m.fzc(k, f);
            const double est = gmu[idx] + rho * c;
            gmu[idx] = est > 0 ? est : 0;
          }
        }
        COOP_SYNC();
        {
          const double r = rho * o.penalty_scaling;
          rho = r < o.penalty_max ? r : o.penalty_max;
        }
        COOP_PHASE {
          double J = 0, vl = 0;
#pragma unroll 1
          for (int k = lane; k <= N; k += G) knot_merit(m, cfg, wr, k, N, X + k * NX, U + k * NU, gmu + k * NC, rho, J, vl);
          red[lane] = J;
          red[G + lane] = vl;
        }
        COOP_SYNC();
        phi = 0;
        viol = 0;
        for (int l = 0; l < G; ++l) {
          phi += red[l];
          viol = red[G + l] > viol ? red[G + l] : viol;
        }
        COOP_SYNC();
      }
    }

    // ---------------- Riccati backward pass.  Lane (br, bc) = (lane / 4, lane % 4) owns the 3x3 block
    // (br, bc) of every 12x12 quantity; all inner indices are compile-time.
    static_assert(G == 16, "block-per-lane mapping assumes 16 lanes per problem");
    bool bp_ok = true;
    double* Pc = P;   // value-function Hessian of knot k+1 (then the not-yet-corrected one of knot k)
    double* Pw = PA;  // work buffer: P A, then Quu and its Cholesky factor, then the new P (ping-pong)
    COOP_PHASE {
#pragma unroll 1
      for (int e = lane; e < 21; e += G) vec[(e < 12 ? cv::pv : cv::Hphi - 12) + e] = gLX[N * 21 + e];
      if (lane == 0) scal[0] = 0.0;
    }
    COOP_SYNC();
    COOP_PHASE {
      const int br = lane >> 2, bc = lane & 3;
      double o[9];
      lxx_block(wq, vec + cv::Hphi, br, bc, o);
      blk_store(Pc + 36 * br + 3 * bc, 12, o);
      blk_store_keep(gP + (size_t)N * 144 + 36 * br + 3 * bc, 12, o);
      if (lane < 12) st_keep(gpv + N * 12 + lane, vec[cv::pv + lane]);
    }
    COOP_SYNC();

#pragma unroll 1
#define COOP_KNOT_LEAVE break
    for (int k = N - 1; k >= 0 && bp_ok; --k) {
      // ---- phase A: stage the knot's 3x3 blocks; per-foot AL terms; cost expansion
      COOP_PHASE {
#pragma unroll 1
        for (int e = lane; e < 27 + 21; e += G) {   // the knot's dynamics blocks and cost expansion
          if (e < 27) lin[e] = glin[k * 27 + e];
          else vec[(e < 27 + 12 ? cv::lx - 27 : cv::Hphi - 39) + e] = gLX[k * 21 + e - 27];
        }
        if (lane < NF) {
          const int f = lane;
          const double* u = U + k * NU + 3 * f;
          double g0 = 0, g1 = 0, g2 = 0, hb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
          for (int r = 0; r < 6; ++r) {
            double c = m.CR[3 * r] * u[0] + m.CR[3 * r + 1] * u[1] + m.CR[3 * r + 2] * u[2];
            if (r == 4) c += -m.fzc(k, f);
            const double est = gmu[k * NC + 6 * f + r] + rho * c;
            if (est > 0) {
              const double j0 = m.CR[3 * r], j1 = m.CR[3 * r + 1], j2 = m.CR[3 * r + 2];
              g0 += j0 * est; g1 += j1 * est; g2 += j2 * est;
              hb[0] += rho * j0 * j0; hb[1] += rho * j0 * j1; hb[2] += rho * j0 * j2;
              hb[3] += rho * j1 * j0; hb[4] += rho * j1 * j1; hb[5] += rho * j1 * j2;
              hb[6] += rho * j2 * j0; hb[7] += rho * j2 * j1; hb[8] += rho * j2 * j2;
            }
          }
          hb[0] += wr[3 * f]; hb[4] += wr[3 * f + 1]; hb[8] += wr[3 * f + 2];
#pragma unroll
          for (int a = 0; a < 9; ++a) vec[cv::Dblk + 9 * f + a] = hb[a];
          vec[cv::g + 3 * f] = wr[3 * f] * u[0] + g0;
          vec[cv::g + 3 * f + 1] = wr[3 * f + 1] * u[1] + g1;
          vec[cv::g + 3 * f + 2] = wr[3 * f + 2] * (u[2] - m.urefz(k, f)) + g2;
        }
      }
      COOP_SYNC();
      const double* Aff = lin;
      const double* Afw = lin + 9;
      const double* Cf = lin + 18;
      const double* pv = vec + cv::pv;
      // ---- phase B: PA = P A (16 blocks), PM = P M (8 blocks), s = M^T p, Atp = A^T p
      COOP_PHASE {
        const int br = lane >> 2, bc = lane & 3;
        const double* Pr = Pc + 36 * br;
        if (bc & 1) blk_right(Pr, 12, bc == 1 ? Aff : Afw, bc == 1 ? 0.0 : 1.0, Pw + 36 * br + 3 * bc, 12);
        else blk_even(Pr, 12, bc == 0 ? 1.0 : hd, bc == 0 ? 0.0 : 1.0, Pw + 36 * br + 3 * bc, 12);
        if (bc < 2) {
          if (bc == 1) blk_right(Pr, 12, Cf, hd, PM + 18 * br + 3 * bc, 6);
          else blk_even(Pr, 12, c1, hd, PM + 18 * br + 3 * bc, 6);
        }
        if (lane < 6) {
          const int e = lane;
          vec[cv::s + e] = e < 3 ? hd * hh * pv[e] + hd * pv[6 + e]
                                 : Cf[e - 3] * pv[3] + Cf[e] * pv[4] + Cf[3 + e] * pv[5] + hd * pv[6 + e];
        }
        if (lane < 12) {
          const int a = lane, ab = a / 3, aa = a % 3;
          double v;
          if (ab == 0) v = pv[a];
          else if (ab == 1) v = Aff[aa] * pv[3] + Aff[3 + aa] * pv[4] + Aff[6 + aa] * pv[5];
          else if (ab == 2) v = hd * pv[aa] + pv[6 + aa];
          else v = Afw[aa] * pv[3] + Afw[3 + aa] * pv[4] + Afw[6 + aa] * pv[5] + pv[9 + aa];
          vec[cv::Atp + a] = v;
        }
      }
      COOP_SYNC();
      // ---- phase C: P <- A^T PA + lxx (16 blocks), T = M^T PA (8 blocks), S = M^T PM (4 blocks), Qx
      COOP_PHASE {
        const int br = lane >> 2, bc = lane & 3;
        const double* Yc = Pw + 3 * bc;
        double* Pd = Pc + 36 * br + 3 * bc;
        if (br & 1) blk_left(Yc, 12, br == 1 ? Aff : Afw, br == 1 ? 0.0 : 1.0, Pd, 12);
        else blk_evenT(Yc, 12, br == 0 ? 1.0 : hd, br == 0 ? 0.0 : 1.0, Pd, 12);
        if (br == bc) {   // + lxx on the diagonal blocks (the same lane just wrote the block)
          if (br == 1) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) Pd[12 * i + j] += vec[cv::Hphi + 3 * i + j];
          } else {
            const int q0 = br == 0 ? 0 : (br == 2 ? 7 : 10);
            Pd[0] += wq[q0]; Pd[13] += wq[q0 + 1]; Pd[26] += wq[q0 + 2];
          }
        }
        if (br < 2) {
          if (br == 1) blk_left(Yc, 12, Cf, hd, T + 36 * br + 3 * bc, 12);
          else blk_evenT(Yc, 12, c1, hd, T + 36 * br + 3 * bc, 12);
        }
        if (lane >= 8 && lane < 12) {
          const int r = (lane >> 1) & 1, c = lane & 1;
          const double* Ym = PM + 3 * c;
          if (r == 1) blk_left(Ym, 6, Cf, hd, S + 18 * r + 3 * c, 6);
          else blk_evenT(Ym, 6, c1, hd, S + 18 * r + 3 * c, 6);
        }
        if (lane < 12) vec[cv::Qx + lane] = vec[cv::Atp + lane] + vec[cv::lx + lane];
      }
      COOP_SYNC();
      // ---- phase D: Qux = W^T T (NF x 4 blocks), SW = S W (2 x NF blocks), Qu = g + W^T s
      COOP_PHASE {
        const int br = lane >> 2, bc = lane & 3;
        if (br < NF) blk_wt(T + 3 * bc, 12, m.inv_mass, m.IS + 9 * br, Qux + 36 * br + 3 * bc, 12, nullptr);
        if (br < 2 && bc < NF) blk_w(S + 18 * br, 6, m.inv_mass, m.IS + 9 * bc, SW + 3 * NU * br + 3 * bc, NU);
        if (lane < NU) {
          const int f = lane / 3, a = lane % 3;
          const double* IS = m.IS + 9 * f;
          const double* s = vec + cv::s;
          vec[cv::Qu + lane] = (m.inv_mass * s[a] + IS[a] * s[3] + IS[3 + a] * s[4] + IS[6 + a] * s[5]) + vec[cv::g + lane];
        }
      }
      COOP_SYNC();
      // ---- phase E: Quu = D + W^T (S W)  (NF x NF blocks, into the work buffer)
      double* Quu = Pw;
      COOP_PHASE {
        const int br = lane >> 2, bc = lane & 3;
        if (br < NF && bc < NF)
          blk_wt(SW + 3 * bc, NU, m.inv_mass, m.IS + 9 * br, Quu + 3 * NU * br + 3 * bc, NU,
                 br == bc ? vec + cv::Dblk + 9 * br : nullptr);
      }
      COOP_SYNC();
      // ---- Cholesky + both triangular solves, fused, per lane, entirely in registers.  Every lane
      //      factors the 12x12 Quu redundantly (the kernel is latency- and shared-memory-bound, not
      //      FLOP-bound: 16 lanes doing the same 364 flops cost the same issue slots as one) and then
      //      solves for its own right-hand side (column `lane` of Qux; lane 12: Qu) with L still in
      //      registers: no column exchange, no barrier, no shared-memory traffic for L (this replaced
      //      ~300 LDS/STS and 24 barriers per knot).  Straight-line code; a non-positive pivot poisons
      //      the lane's result and is reported through `ok`.  Same operation order per entry as the
      //      variants below, so the results are bit-identical.
      COOP_PHASE {
        const int c = lane < 12 ? lane : 12;   // lanes 13..15 shadow the Qu column and store nothing
        constexpr int NT = NU * (NU + 1) / 2;
#define QMPC_TRI(i_, l_) ((i_) * ((i_) + 1) / 2 + (l_))
        double Lr[NT], rd[NU], rhs[NU];
#define QMPC_DIVD(x_, i_) { (x_) = (x_) * rd[i_]; }
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
          for (int l = 0; l <= i; ++l) Lr[QMPC_TRI(i, l)] = Quu[NU * i + l];
#pragma unroll
        for (int i = 0; i < NU; ++i) rhs[i] = c < 12 ? Qux[12 * i + c] : vec[cv::Qu + i];
        bool ok = true;
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          const double sjj = Lr[QMPC_TRI(j, j)];
          ok = ok && (sjj > 0.0);
#ifdef QMPC_COOP_FAST_RECIP
          const double rdg = qmpc_rsqrt(sjj);   // L(i,j) = a * rsqrt: 1-2 ulp from a / sqrt(sjj)
          rd[j] = rdg;
#pragma unroll
          for (int i = j + 1; i < NU; ++i) Lr[QMPC_TRI(i, j)] *= rdg;
#else
          // sqrt and the divisions of the reference Cholesky, built from rsqrt + FMA corrections (Markstein): the
          // quotients come out correctly rounded, i.e. equal to a / sqrt(sjj), at 2 extra FMAs each instead of a
          // ~25-instruction IEEE division.  The plain reciprocal form was 1-2 ulp off and, with cond(Quu) up to
          // 1e14, moved 1 solve in 65 536 by 3e-4 N against the oracle (division-based kernels: 4e-5 N there).
          // Only the factor's quotients matter; the substitutions below keep the plain reciprocal (measured).
          const double r0 = qmpc_rsqrt(sjj);
          double dg = sjj * r0;
          dg = fma(0.5 * fma(-dg, dg, sjj), r0, dg);          // sqrt(sjj), correctly rounded
          const double rdg = fma(fma(-dg, r0, 1.0), r0, r0);  // 1 / dg
          rd[j] = rdg;
#pragma unroll
          for (int i = j + 1; i < NU; ++i) {
            const double a = Lr[QMPC_TRI(i, j)], q = a * rdg;
            Lr[QMPC_TRI(i, j)] = fma(fma(-q, dg, a), rdg, q);   // a / dg
          }
#endif
#pragma unroll
          for (int i = j + 1; i < NU; ++i)
#pragma unroll
            for (int l = j + 1; l <= i; ++l) Lr[QMPC_TRI(i, l)] -= Lr[QMPC_TRI(i, j)] * Lr[QMPC_TRI(l, j)];
        }
        if (!ok) bp_ok = false;
        // forward substitution, column oriented: after y_i is final every remaining entry updates independently
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          QMPC_DIVD(rhs[i], i);
#pragma unroll
          for (int l = i + 1; l < NU; ++l) rhs[l] -= Lr[QMPC_TRI(l, i)] * rhs[i];
        }
        if (ok && lane <= 12) {
          if (c < 12) {
#pragma unroll
            for (int i = 0; i < NU; ++i) Qux[12 * i + c] = rhs[i];   // V = L^-1 Qux
          } else {
#pragma unroll
            for (int i = 0; i < NU; ++i) vec[cv::vu + i] = rhs[i];
          }
        }
#pragma unroll
        for (int i = NU - 1; i >= 0; --i) {
          QMPC_DIVD(rhs[i], i);
#pragma unroll
          for (int l = 0; l < i; ++l) rhs[l] -= Lr[QMPC_TRI(i, l)] * rhs[i];
        }
#undef QMPC_TRI
#undef QMPC_DIVD
        if (ok && lane <= 12) {
          if (c < 12) {
#pragma unroll
            for (int i = 0; i < NU; ++i) st_keep(gK + ((size_t)k * NU + i) * 12 + c, -rhs[i]);
          } else {
            double t = 0;
#pragma unroll
            for (int i = 0; i < NU; ++i) {
              st_keep(gd + k * NU + i, -rhs[i]);
              t += vec[cv::Qu + i] * (-rhs[i]);
            }
            scal[0] += t;
          }
        }
      }
      COOP_SYNC();
      if (!bp_ok) COOP_KNOT_LEAVE;
      // ---- phase F: new P = sym(P) - V^T V (16 blocks, written to the work buffer: no race with the
      //      transposed reads of Pc), pv <- Qx - V^T vu ; then swap the two buffers
      COOP_PHASE {
        const int br = lane >> 2, bc = lane & 3;
        double o[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#ifndef QMPC_COOP_F_UNROLL
#define QMPC_COOP_F_UNROLL 1
#endif
        constexpr int kFUnroll = QMPC_COOP_F_UNROLL;
#pragma unroll(kFUnroll)
        for (int l = 0; l < NU; ++l) {
          const double* Vr = Qux + 12 * l + 3 * br;
          const double* Vc = Qux + 12 * l + 3 * bc;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) o[3 * a + b] += Vr[a] * Vc[b];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b)
            o[3 * a + b] = 0.5 * (Pc[12 * (3 * br + a) + 3 * bc + b] + Pc[12 * (3 * bc + b) + 3 * br + a]) - o[3 * a + b];
        blk_store(Pw + 36 * br + 3 * bc, 12, o);
        blk_store_keep(gP + (size_t)k * 144 + 36 * br + 3 * bc, 12, o);
        if (lane < 12) {
          const int a = lane;
          double t = 0;
#pragma unroll 4
          for (int l = 0; l < NU; ++l) t += Qux[12 * l + a] * vec[cv::vu + l];
          const double v = vec[cv::Qx + a] - t;
          vec[cv::pv + a] = v;
          st_keep(gpv + k * 12 + a, v);
        }
      }
      COOP_SYNC();
      { double* t = Pc; Pc = Pw; Pw = t; }
    }
    if (!bp_ok) { status = QMPC_STATUS_BACKWARD_FAILED; COOP_BLOCK_SYNC_MID(); COOP_ITER_LEAVE; }
    COOP_BLOCK_SYNC_MID();
    const double dphi0 = scal[0];

    // ---------------- forward pass: speculative back-tracking line search.  Each round evaluates NCAND
    // consecutive step lengths alpha = decrease^j at once (feet roll-out: NF lanes per step length, 4 per
    // round for the quadruped; lane roll-out: one lane each, 16 per round); the first one passing the
    // Armijo test wins - the result is identical to the sequential search.
    int acc_j = -1;
    double phin = 0, violn = 0, alpha_acc = 0;
#pragma unroll 1
    for (int round = 0; round * G < o.ls_iters_max && acc_j < 0; ++round) {
      COOP_PHASE {
        const int j = round * G + lane;
        // every lane rolls out (lanes past ls_iters_max too: the roll-out stages the gains
        // cooperatively); their result is discarded below
        double J = NAN, vl = 0, alpha = 1.0;
        for (int q = 0; q < j; ++q) alpha *= o.ls_decrease;
        coop_rollout<NF>(m, cfg, wr, N, h, X, U, DX, gK, gd, gmu, rho, alpha, 1, &J, &vl, gTX, gTU, lane, G, P, lane_mask, nullptr);
        red[lane] = j < o.ls_iters_max ? J : NAN;
        red[G + lane] = vl;
      }
      COOP_SYNC();
      {
        double alpha = 1.0;
        for (int q = 0; q < round * G; ++q) alpha *= o.ls_decrease;
        for (int l = 0; l < G && acc_j < 0; ++l) {
          const double pl = red[l];
          if (round * G + l < o.ls_iters_max && isfinite(pl) && pl <= phi + o.ls_c1 * alpha * dphi0) {
            acc_j = round * G + l;
            phin = pl;
            violn = red[G + l];
            alpha_acc = alpha;
          }
          alpha *= o.ls_decrease;
        }
      }
      COOP_SYNC();
    }
    iters = it + 1;
    if (acc_j < 0) { status = QMPC_STATUS_LINESEARCH_FAILED; COOP_ITER_LEAVE; }
    // ---------------- accepted step: the winning lane's trial trajectory is already in the scratch.
    // lane k <- knot k: dx_k = x_new (-) x_old, Riccati dual y_k = P_k dx_k + p_k (stored in DX) ...
    const int acc_lane = acc_j % NCAND;
    COOP_PHASE {
#pragma unroll 1
      for (int k = lane; k <= N; k += G) {
        double xn[NX], dx[NE];
#pragma unroll
        for (int i = 0; i < NX; ++i) xn[i] = ld_stream(gTX + (size_t)(k * NX + i) * NCAND + acc_lane);
        state_diff<M>(xn, X + k * NX, dx);
        const double* Pk = gP + (size_t)k * 144;
#ifndef QMPC_COOP_ACCEPT_UNROLL
#define QMPC_COOP_ACCEPT_UNROLL 3
#endif
        // nearly rolled (once per iteration; the kernel is instruction-cache sensitive) yet several rows of
        // P_k - 12 L2 loads each - are in flight per trip
        constexpr int kAcceptUnroll = QMPC_COOP_ACCEPT_UNROLL;
#pragma unroll(kAcceptUnroll)
        for (int a = 0; a < NE; ++a) {
          double t = ld_keep(gpv + k * 12 + a);
#pragma unroll
          for (int b = 0; b < NE; ++b) t += ld_keep(Pk + 12 * a + b) * dx[b];
          DX[k * NE + a] = t;
        }
      }
    }
    COOP_SYNC();
    // ... then X, U <- accepted trajectory (cooperative strided copy)
    COOP_PHASE {
#pragma unroll 1
      for (int e = lane; e < (N + 1) * NX; e += G) X[e] = ld_stream(gTX + (size_t)e * NCAND + acc_lane);
#pragma unroll 1
      for (int e = lane; e < N * NU; e += G) U[e] = ld_stream(gTU + (size_t)e * NCAND + acc_lane);
    }
    COOP_SYNC();
    cost_decrease = phi - phin;
    phi = phin;
    viol = violn;
  }

  COOP_PHASE {
    if (lane == 0) {
      QmpcResult r;
      m.write_result(U, r);
      r.max_violation = viol;
      r.iterations = iters;
      r.status = status;
      out[pid] = r;
      if (warm) warm[pid].valid = status != QMPC_STATUS_NONFINITE;
    }
    if (warm) {
#pragma unroll 1
      for (int e = lane; e < N * 12; e += G) {
        const int k = e / 12, i = e % 12;
        warm[pid].u[k][i] = i < NU ? U[k * NU + i] : 0.0;
      }
    }
  }
  COOP_SYNC();
}

#ifdef __CUDACC__
// persistent launch: every group of G lanes is a "slot" that strides over the batch
#ifndef QMPC_COOP_BLOCK
#define QMPC_COOP_BLOCK 128        // 4 warps = 8 problems per block
#endif
#ifndef QMPC_COOP_MIN_BLOCKS
#define QMPC_COOP_MIN_BLOCKS (256 / QMPC_COOP_BLOCK)   // 8 warps per SM at 255 registers
#endif
template <int NF, int G>
__global__ void __launch_bounds__(QMPC_COOP_BLOCK, QMPC_COOP_MIN_BLOCKS)
qmpc_coop_kernel(QmpcConfig cfg, SolverOpts o, const QmpcProblem* __restrict__ in,
                 const unsigned char* __restrict__ sched, QmpcWarmStart* __restrict__ warm,
                 QmpcResult* __restrict__ out,
                 double* __restrict__ scratch, int batch, int smem_per_problem, size_t scratch_per_slot, int wide,
                 int active_groups) {
  extern __shared__ __align__(16) double smem_pool[];
  if (threadIdx.x < 13) smem_pool[threadIdx.x] = cfg.q_weights[threadIdx.x];
  else if (threadIdx.x < 25) smem_pool[threadIdx.x] = cfg.r_weights[threadIdx.x - 13];
  __syncthreads();
  const int groups_per_block = blockDim.x / G;
  const int group = threadIdx.x / G;
  const int lane_id = threadIdx.x % G;
  // `active_groups` of the block's groups own a slot (a partial wave is spread over all SMs instead of filling
  // some of them): even groups first, so that up to half occupancy every problem has a warp of its own
  const int rank = (group & 1) * ((groups_per_block + 1) / 2) + (group >> 1);
  const bool idle = rank >= active_groups;
  const int slot = blockIdx.x * active_groups + rank;
  const int nslots = gridDim.x * active_groups;
  const unsigned lane_mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x % 32) / G * G));
  double* sm = smem_pool + kCoopBlockShared + (size_t)group * smem_per_problem;
  double* gs = scratch + (size_t)(idle ? 0 : slot) * scratch_per_slot;
#ifdef QMPC_COOP_BLOCK_SYNC
  // problem waves: all slots of the block run the same number of waves and barriers; a slot without a
  // problem in the last wave only passes the barriers
  for (int base = 0; base < batch; base += nslots) {
    const int pid = base + slot;
    if (!idle && pid < batch) {
      coop_solve_one<NF, G>(cfg, o, in, sched, warm, out, pid, sm, gs, lane_id, lane_mask, wide, smem_pool);
    } else {
      for (int it = 0; it < o.iterations_max; ++it) { COOP_BLOCK_SYNC(); COOP_KNOT_SYNC_ALL(o.N); COOP_BLOCK_SYNC_MID(); }
    }
  }
#else
  for (int pid = slot; !idle && pid < batch; pid += nslots) {
    coop_solve_one<NF, G>(cfg, o, in, sched, warm, out, pid, sm, gs, lane_id, lane_mask, wide, smem_pool);
  }
#endif
}
#endif

}  // namespace qmpc
