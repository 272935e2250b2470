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
// Round 2: the solve is written as PHASE FUNCTIONS over a per-problem context (CoopCtx): set-up, pre
// (expansions, stationarity, dual update), backward (Riccati), forward (line search + accepted step),
// epilogue.  The fused persistent kernel calls them in sequence; the phased kernels (qmpc_phased.cuh) run
// one phase per launch with their own register / shared-memory budgets, state through L2/HBM.  The bodies
// are generic over the model: QuatModel<NF> (QuatMpc, 2-contact model) and ConvexModel (ConvexMpc's Euler
// SRB: state blocks swapped pairwise, a fourth knot-dependent block Dw) - see the traits in qmpc_models.cuh.
//
// The body is written with COOP_PHASE / COOP_SYNC so that the very same source runs on the host
// (tests/emul, lanes executed one after another) for GPU-less debugging.
#pragma once
#include <cstddef>
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

// The roll-out is inlined at its two call sites (nominal roll-out: mode 0; line search: mode 1): each copy is
// specialised for its mode and, above all, the call no longer spills the caller's live registers around it
// (31 LDL / 61 STL static in the out-of-line build).  Measured +6.7 % at B = 4096, +6.5 % at B = 65 536
// (profiles/r02_experiments.md); round 1 had three call sites and measured the opposite.
#if defined(__CUDACC__) && defined(QMPC_COOP_ROLLOUT_NOINLINE)
#define QMPC_NOINLINE __noinline__
#elif defined(__CUDACC__)
#define QMPC_NOINLINE __forceinline__
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
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_LIBM_RSQRT)
  // rsqrt() of the CUDA math library without its range checks (zero / negative / denormal / non-finite arguments branch
  // to a slow path there; here a non-positive pivot is reported through `ok` and poisons the lane's result anyway): the
  // same seed (MUFU.RSQ64H, low word 0) and the same refinement y0 + (0.5 + 0.375 e) (y0 e), e = 1 - x y0^2 - the same
  // value for every normal positive x, 6 instructions for 9 + a branch.
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(-x, y0 * y0, 1.0);
  return fma(fma(e, 0.375, 0.5), y0 * e, y0);
#elif defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

// max(0, x) of the AL terms.  Written as `x > 0 ? x : 0` the compiler emits max.f64, which sm_100a has no instruction
// for: ptxas expands it to DSETP.MAX + SEL + FSEL + LOP3 (NaN quieting) + register moves, 6-7 instructions, 24 times
// per knot of every roll-out.  setp + selp is 3, and returns the same value for every input (NaN -> 0, -0 -> +0).
QMPC_HD inline double pos_part(double x) {
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_POS_PART_MAX)
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, 0d0000000000000000;\n\tselp.f64 %0, %1, 0d0000000000000000, p;\n\t}"
      : "=d"(r) : "d"(x));
  return r;
#else
  return x > 0 ? x : 0;
#endif
}

// m0 p0 + m1 p1 + m2 p2 + beta pb in the order nvcc contracts that expression - written out with fma on the device so
// that operands which are selects of constants (beta = cond ? h : 0) cannot be folded into the products differently
QMPC_HD inline double coop_dot4(double m0, double p0, double m1, double p1, double m2, double p2, double beta, double pb) {
#ifdef __CUDA_ARCH__
  return fma(beta, pb, fma(m2, p2, fma(m0, p0, __dmul_rn(m1, p1))));
#else
  return m0 * p0 + m1 * p1 + m2 * p2 + beta * pb;
#endif
}

// ---- 3x3 block kernels used by the block-per-lane phases (lane = (block row, block col)) --------
// The row loop is deliberately NOT unrolled: the kernel is instruction-fetch bound, compact code wins.
// Operands that are reused across the rows are read into registers ONCE, before the first store:
// source and destination are plain pointers into the same shared-memory pool, so the compiler must
// otherwise assume every store clobbers them and reload (measured: 7 LDS per 4 flops; now ~3).
// Offsets: oa = first column (row) of the attitude block, ob = of the angular-velocity block for the "right" /
// "left" helpers; of the position and linear-velocity blocks for the "even" ones (memory order depends on the
// model: QuatModel 3, 9 / 0, 6; ConvexModel 0, 6 / 3, 9).
// dst = X(:, oa:oa+3) * Mt + beta * X(:, ob:ob+3)   X: 3 rows of a row-major matrix with leading dim ld
#ifndef QMPC_COOP_BLK_ROW_UNROLL
#define QMPC_COOP_BLK_ROW_UNROLL 3   // rows of the 3x3 block helpers: 3 = the three rows' FMA chains overlap (measured +1.0 ... +1.8 %
                                     // over 1 = rolled, once the roll-out was inlined; round 1 measured the opposite)
#endif
constexpr int kBlkRowUnroll = QMPC_COOP_BLK_ROW_UNROLL;
// Phases B / C give every lane of a problem one 3x3 block of P A, A^T (P A), P M, M^T (P A), M^T (P M).  The blocks of
// the "odd" roles (attitude, angular velocity) are genuine 3x3 products, those of the "even" roles (position, linear
// velocity) are alpha X_a + beta X_b.  Written as two helpers the lanes of a warp DIVERGE and every phase issues both
// instruction streams (phase C: six helpers in sequence; that form is in the history, commit 943f4b3).  Uniform form: the
// even lanes call the odd lanes' helper with the constant blocks I / h I from the block's shared memory and the
// operands exchanged - fma(beta, x_b, 1 * x_a + 0 + 0) is the very fma(alpha, x, beta y) the even helper evaluates,
// rounding for rounding, so the results are bit-identical (+7 %) - and T and S share one call with per-lane operands.
// The Euler model's moment blocks need the two-matrix helpers (blk_right2 / blk_left2) and keep blk_even / blk_evenT
// for the force blocks next to them.
QMPC_HD inline void blk_right(const double* X, int ld, int oa, int ob, const double* Mt, double beta, double* dst, int ldd) {
  double m[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = Mt[i];
#pragma unroll(kBlkRowUnroll)
  for (int a = 0; a < 3; ++a) {
    const double* Xa = X + ld * a;
    const double x3 = Xa[oa], x4 = Xa[oa + 1], x5 = Xa[oa + 2], y0 = Xa[ob], y1 = Xa[ob + 1], y2 = Xa[ob + 2];
    const double r0 = x3 * m[0] + x4 * m[3] + x5 * m[6] + beta * y0;
    const double r1 = x3 * m[1] + x4 * m[4] + x5 * m[7] + beta * y1;
    const double r2 = x3 * m[2] + x4 * m[5] + x5 * m[8] + beta * y2;
    dst[ldd * a] = r0; dst[ldd * a + 1] = r1; dst[ldd * a + 2] = r2;
  }
}
// dst = X(:, oa:oa+3) * Mt + X(:, ob:ob+3) * Nt   (ConvexModel: the moment column of P M, Nt = Dw)
QMPC_HD inline void blk_right2(const double* X, int ld, int oa, int ob, const double* Mt, const double* Nt, double* dst, int ldd) {
  double m[9], n[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { m[i] = Mt[i]; n[i] = Nt[i]; }
#pragma unroll(kBlkRowUnroll)
  for (int a = 0; a < 3; ++a) {
    const double* Xa = X + ld * a;
    const double x3 = Xa[oa], x4 = Xa[oa + 1], x5 = Xa[oa + 2], y0 = Xa[ob], y1 = Xa[ob + 1], y2 = Xa[ob + 2];
    const double r0 = x3 * m[0] + x4 * m[3] + x5 * m[6] + (y0 * n[0] + y1 * n[3] + y2 * n[6]);
    const double r1 = x3 * m[1] + x4 * m[4] + x5 * m[7] + (y0 * n[1] + y1 * n[4] + y2 * n[7]);
    const double r2 = x3 * m[2] + x4 * m[5] + x5 * m[8] + (y0 * n[2] + y1 * n[5] + y2 * n[8]);
    dst[ldd * a] = r0; dst[ldd * a + 1] = r1; dst[ldd * a + 2] = r2;
  }
}
// dst = alpha * X(:, oa:oa+3) + beta * X(:, ob:ob+3)
QMPC_HD inline void blk_even(const double* X, int ld, int oa, int ob, double alpha, double beta, double* dst, int ldd) {
#pragma unroll(kBlkRowUnroll)
  for (int a = 0; a < 3; ++a) {
    const double* Xa = X + ld * a;
    const double x0 = Xa[oa], x1 = Xa[oa + 1], x2 = Xa[oa + 2], y0 = Xa[ob], y1 = Xa[ob + 1], y2 = Xa[ob + 2];
    dst[ldd * a] = alpha * x0 + beta * y0;
    dst[ldd * a + 1] = alpha * x1 + beta * y1;
    dst[ldd * a + 2] = alpha * x2 + beta * y2;
  }
}
// dst = Mt^T * Y(oa:oa+3, :) + beta * Y(ob:ob+3, :)   Y: 3 columns (starting at Y) of a row-major matrix
QMPC_HD inline void blk_left(const double* Y, int ld, int oa, int ob, const double* Mt, double beta, double* dst, int ldd) {
  double m[9], y[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = Mt[i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int b = 0; b < 3; ++b) y[3 * i + b] = Y[ld * (oa + i) + b];
#pragma unroll   // unrolled: m[a] must stay a compile-time register index
  for (int a = 0; a < 3; ++a) {
    const double* Ya = Y + ld * (ob + a);
    const double z0 = Ya[0], z1 = Ya[1], z2 = Ya[2];
    const double m0 = m[a], m1 = m[3 + a], m2 = m[6 + a];
    const double r0 = m0 * y[0] + m1 * y[3] + m2 * y[6] + beta * z0;
    const double r1 = m0 * y[1] + m1 * y[4] + m2 * y[7] + beta * z1;
    const double r2 = m0 * y[2] + m1 * y[5] + m2 * y[8] + beta * z2;
    dst[ldd * a] = r0; dst[ldd * a + 1] = r1; dst[ldd * a + 2] = r2;
  }
}
// blk_left + a 3x3 block held in registers, added to the rounded result before the store (phase C: the cost
// Hessian block; it used to be added by a load-add-store pass over the block the lane had just stored)
QMPC_HD inline void blk_left_add(const double* Y, int ld, int oa, int ob, const double* Mt, double beta, const double* add,
                                 double* dst, int ldd) {
  double m[9], y[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = Mt[i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int b = 0; b < 3; ++b) y[3 * i + b] = Y[ld * (oa + i) + b];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double* Ya = Y + ld * (ob + a);
    const double z0 = Ya[0], z1 = Ya[1], z2 = Ya[2];
    const double m0 = m[a], m1 = m[3 + a], m2 = m[6 + a];
    const double r0 = m0 * y[0] + m1 * y[3] + m2 * y[6] + beta * z0;
    const double r1 = m0 * y[1] + m1 * y[4] + m2 * y[7] + beta * z1;
    const double r2 = m0 * y[2] + m1 * y[5] + m2 * y[8] + beta * z2;
    dst[ldd * a] = r0 + add[3 * a]; dst[ldd * a + 1] = r1 + add[3 * a + 1]; dst[ldd * a + 2] = r2 + add[3 * a + 2];
  }
}
// dst = Mt^T * Y(oa:oa+3, :) + Nt^T * Y(ob:ob+3, :)   (ConvexModel: the moment rows of M^T Y, Nt = Dw)
QMPC_HD inline void blk_left2(const double* Y, int ld, int oa, int ob, const double* Mt, const double* Nt, double* dst, int ldd) {
  double m[9], n[9], y[9], z[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { m[i] = Mt[i]; n[i] = Nt[i]; }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int b = 0; b < 3; ++b) { y[3 * i + b] = Y[ld * (oa + i) + b]; z[3 * i + b] = Y[ld * (ob + i) + b]; }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double m0 = m[a], m1 = m[3 + a], m2 = m[6 + a], n0 = n[a], n1 = n[3 + a], n2 = n[6 + a];
    const double r0 = m0 * y[0] + m1 * y[3] + m2 * y[6] + (n0 * z[0] + n1 * z[3] + n2 * z[6]);
    const double r1 = m0 * y[1] + m1 * y[4] + m2 * y[7] + (n0 * z[1] + n1 * z[4] + n2 * z[7]);
    const double r2 = m0 * y[2] + m1 * y[5] + m2 * y[8] + (n0 * z[2] + n1 * z[5] + n2 * z[8]);
    dst[ldd * a] = r0; dst[ldd * a + 1] = r1; dst[ldd * a + 2] = r2;
  }
}
// dst = alpha * Y(oa:oa+3, :) + beta * Y(ob:ob+3, :)
QMPC_HD inline void blk_evenT(const double* Y, int ld, int oa, int ob, double alpha, double beta, double* dst, int ldd) {
#pragma unroll(kBlkRowUnroll)
  for (int a = 0; a < 3; ++a) {
    const double x0 = Y[ld * (oa + a)], x1 = Y[ld * (oa + a) + 1], x2 = Y[ld * (oa + a) + 2];
    const double y0 = Y[ld * (ob + a)], y1 = Y[ld * (ob + a) + 1], y2 = Y[ld * (ob + a) + 2];
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
#pragma unroll(kBlkRowUnroll)
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

// doubles at the head of the block's shared memory: q[13], r[12], pad, then two constant 3x3 blocks - I at 26 and
// h I at 35 - the operands that let the "even" lanes of phases B / C run the very block product of the odd lanes
constexpr int kCoopBlockShared = 44;
constexpr int kCoopI3 = 26, kCoopHI3 = 35, kCoopC1 = 25;   // slot 25: c1 = h (h / 2), the (position, force) entry of M
QMPC_HD inline double coop_block_const(const QmpcConfig& cfg, float h, int i) {
  if (i < 13) return cfg.q_weights[i];
  if (i < 25) return cfg.r_weights[i - 13];
  if (i < kCoopI3) return (double)h * (double)(h / 2);
  const int e = (i - kCoopI3) % 9;
  return (e % 4 == 0) ? (i < kCoopHI3 ? 1.0 : (double)h) : 0.0;
}

template <int V>
struct IntTag { static constexpr int value = V; };

QMPC_HD constexpr int coop_even(int v) { return (v + 1) / 2 * 2; }

// Per-knot row of the expansions, contiguous and 16-byte aligned so that ONE bulk copy (cp.async / TMA 1-D)
// brings a knot's data into the shared "vec" block: cost gradient lx (12), attitude Hessian block (9), input
// gradient g = R (u - u_ref) + J^T max(0, mu + rho c) (NU), D blocks = R + rho J_a^T J_a per foot (9 NF).
template <class M>
struct CoopRow {
  static constexpr int NF = M::kFeet, NU = M::NU;
  static constexpr int lx = 0, Hphi = 12, g = 21, Dblk = 21 + NU, kLen = 21 + NU + 9 * NF, kStride = coop_even(kLen);
};

template <class M, int G>
struct CoopLayout {
  static constexpr int NF = M::kFeet, NU = M::NU, NC = M::NC, NX = M::NX, NLIN = M::NLIN;
  static constexpr int kLinStride = coop_even(NLIN);
  static constexpr int kRow = CoopRow<M>::kStride;
  static constexpr int kKD = NU * 12 + NU;   // one knot's gain matrix K_k followed by its feed-forward d_k
  static constexpr int kModel = (int)((sizeof(M) + 7) / 8);
  // ---- shared memory (doubles) per problem
  // vec block: the knot row (kRow), then Qx 12, Qu 12, s 6, Atp 12, vu 12, pv 12, scal 10
  static constexpr int vQx = kRow, vQu = vQx + 12, vs = vQu + 12, vAtp = vs + 6, vvu = vAtp + 12, vpv = vvu + 12,
                       vscal = vpv + 12, kVec = coop_even(vscal + 10);
  QMPC_HD static int sX(int N) { return coop_even(kModel); }
  QMPC_HD static int sU(int N) { return sX(N) + (N + 1) * NX; }
  QMPC_HD static int sP(int N) { return coop_even(sU(N) + N * NU); }   // 16-byte aligned: cp.async target
  QMPC_HD static int sPA(int N) { return sP(N) + 144; }    // PA, later Quu / its Cholesky factor
  QMPC_HD static int sT(int N) { return sPA(N) + 144; }
  QMPC_HD static int sPM(int N) { return sT(N) + 72; }     // PM, later SW
  QMPC_HD static int sS(int N) { return sPM(N) + 72; }
  QMPC_HD static int sQux(int N) { return sS(N) + 36; }    // Qux, later V = L^-1 Qux
  QMPC_HD static int sVec(int N) { return coop_even(sQux(N) + NU * 12); }
  QMPC_HD static int sLin(int N) { return sVec(N) + kVec; }
  // Optional residents (`flags`): bit 0 = the per-knot linearisation blocks (kLinStride N), bit 1 = the duals
  // (NC N) also live in shared memory - used when they do not cost residency (short horizons);
  // otherwise they stay in the L2-resident scratch.  Same code either way, only the pointers differ.
  QMPC_HD static int sLinAll(int N) { return sLin(N) + kLinStride; }
  QMPC_HD static int sMu(int N, int flags) { return sLinAll(N) + ((flags & 1) ? kLinStride * N : 0); }
  QMPC_HD static int smem_doubles(int N, int flags) { return coop_even(sMu(N, flags) + ((flags & 2) ? NC * N : 0)); }
  // shared memory of the forward-only kernel (qmpc_phased.cuh): model, X, U, the gain stage (also the dx buffer
  // of the accepted step) and the reduction slots
  QMPC_HD static int fStage(int N) { return coop_even(sU(N) + N * NU); }
  QMPC_HD static int fStageLen(int N) { int a = 2 * kKD + 2, b = (N + 1) * 12; return coop_even(a > b ? a : b); }   // + 2 mbarriers
  QMPC_HD static int fRed(int N) { return fStage(N) + fStageLen(N); }
  QMPC_HD static int fwd_smem_doubles(int N) { return fRed(N) + 2 * G; }
  // ---- global scratch (doubles) per slot (fused kernel) / per problem (phased kernels); every region starts
  //      16-byte aligned
  QMPC_HD static size_t gK(int N) { return 0; }                                          // N x [K_k | d_k]
  QMPC_HD static size_t gP(int N) { return gK(N) + (size_t)N * kKD; }
  QMPC_HD static size_t gpv(int N) { return gP(N) + (size_t)(N + 1) * 144; }
  QMPC_HD static size_t gmu(int N) { return gpv(N) + (size_t)(N + 1) * 12; }
  QMPC_HD static size_t glin(int N) { return coop_even((int)(gmu(N) + (size_t)N * NC)); }
  QMPC_HD static size_t gDX(int N) { return glin(N) + (size_t)N * kLinStride; }
  QMPC_HD static size_t gLX(int N) { return gDX(N) + (size_t)(N + 1) * 12; }            // (N + 1) knot rows
  QMPC_HD static size_t gEnd(int N) { return gLX(N) + (size_t)(N + 1) * kRow; }
  // trial trajectories of the speculative line search, [element][lane] so the 16 lanes store coalesced
  QMPC_HD static size_t gTX(int N) { return (gEnd(N) + 15) / 16 * 16; }
  QMPC_HD static size_t gTU(int N) { return gTX(N) + (size_t)(N + 1) * NX * G; }
  QMPC_HD static size_t scratch_doubles(int N) { return (gTU(N) + (size_t)N * NU * G + 15) / 16 * 16; }
  // phased kernels: persistent per-problem block = the regions above up to gEnd, then model, X, U, scalars
  QMPC_HD static size_t pModel(int N) { return coop_even((int)gEnd(N)); }
  QMPC_HD static size_t pX(int N) { return pModel(N) + coop_even(kModel); }
  QMPC_HD static size_t pU(int N) { return pX(N) + (size_t)(N + 1) * NX; }
  QMPC_HD static size_t pScal(int N) { return coop_even((int)(pU(N) + (size_t)N * NU)); }
  QMPC_HD static size_t problem_doubles(int N) { return (pScal(N) + 8 + 15) / 16 * 16; }
  // phased forward kernel: per-slot trial trajectories only
  QMPC_HD static size_t trial_doubles(int N) { return ((size_t)(N + 1) * NX * G + (size_t)N * NU * G + 15) / 16 * 16; }
};

// the three (four) knot-dependent 3x3 blocks of the error-state linearisation
struct KnotLin4 : KnotLin {
  double Dw[9];
};

// ---- model dispatch: linearisation blocks of one knot
template <int NF>
QMPC_HD inline void coop_linearize(const QuatModel<NF>& m, const double* x, const double* u, const double* xn, double hd,
                                   double hh, KnotLin4& L) {
  srb_linearize(m, x, u, xn, hd, hh, L);
}
// Euler SRB (AltroUtils.cpp:224-359 through the midpoint chain rule :78-110): with the state blocks
// [theta, p, omega, v], T = d(theta_dot)/d(yaw) (only column 2 non-zero), Rt = the yaw-only map omega -> rpy rates
//   Aff = I + h T_m                      Afw = h (Rt_m + (h/2) T_m Rt)
//   Cf  = h (h/2) Rt_m Iw(yaw)^-1        Dw  = h Iw(yaw_m)^-1
// entry by entry as v = h (h/2 * s + Jm) like the dense chain rule (terms that are exactly zero dropped).
QMPC_HD inline void coop_linearize(const ConvexModel& m, const double* x, const double* u, const double* /*xn*/, double hd,
                                   double hh, KnotLin4& L) {
  double fs0 = 0, fs1 = 0, fs2 = 0, mom0 = 0, mom1 = 0, mom2 = 0;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const double* rf = m.foot + 3 * f;
    const double u0 = u[3 * f], u1 = u[3 * f + 1], u2 = u[3 * f + 2];
    mom0 += rf[1] * u2 - rf[2] * u1; mom1 += rf[2] * u0 - rf[0] * u2; mom2 += rf[0] * u1 - rf[1] * u0;
    fs0 += u0; fs1 += u1; fs2 += u2;
  }
  double xd[12];
  m.wrench_dyn(x, fs0, fs1, fs2, mom0, mom1, mom2, xd);
  const double yaw_m = xd[2] * hh + x[2], w0m = xd[6] * hh + x[6], w1m = xd[7] * hh + x[7];
  double sy, cy, sym, cym;
  qmpc_sincos(x[2], &sy, &cy);
  qmpc_sincos(yaw_m, &sym, &cym);
  double iw[4], iwm[4];
  ConvexModel::Iw_inv(sy, cy, iw);
  ConvexModel::Iw_inv(sym, cym, iwm);
  const double am = w1m * cym - w0m * sym, bm = -w0m * cym - w1m * sym;   // AltroUtils.cpp:355-356 at the midpoint
#pragma unroll
  for (int i = 0; i < 9; ++i) { L.Aff[i] = 0; L.Afw[i] = 0; L.Cf[i] = 0; L.Dw[i] = 0; }
  L.Aff[0] = 1.0; L.Aff[4] = 1.0; L.Aff[8] = 1.0;
  L.Aff[2] = hd * am; L.Aff[5] = hd * bm;
  const double Rt[9] = {cy, sy, 0, -sy, cy, 0, 0, 0, 1}, Rtm[9] = {cym, sym, 0, -sym, cym, 0, 0, 0, 1};
  (void)Rt;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double s = (j == 2) ? (i == 0 ? am : (i == 1 ? bm : 0.0)) : 0.0;   // T_m(i,2) * Rt(2,j), Rt(2,:) = (0,0,1)
      L.Afw[3 * i + j] = hd * (hh * s + Rtm[3 * i + j]);
    }
  const double Iw[9] = {iw[0], iw[1], 0, iw[1], iw[2], 0, 0, 0, iw[3]};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0;
#pragma unroll
      for (int l = 0; l < 3; ++l) s += Rtm[3 * i + l] * Iw[3 * l + j];
      L.Cf[3 * i + j] = hd * (hh * s);
    }
  L.Dw[0] = hd * iwm[0]; L.Dw[1] = hd * iwm[1]; L.Dw[3] = hd * iwm[1]; L.Dw[4] = hd * iwm[2]; L.Dw[8] = hd * iwm[3];
}

// y = A^T v (12) and t = M^T v (6) in the model's memory order of the state blocks
template <class M>
QMPC_HD inline void coop_At_vec(const KnotLin4& L, double hd, const double* v, double* y) {
  constexpr int P = 3 * (0 ^ M::kSwap), A = 3 * (1 ^ M::kSwap), V = 3 * (2 ^ M::kSwap), W = 3 * (3 ^ M::kSwap);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    y[P + i] = v[P + i];
    y[A + i] = L.Aff[i] * v[A] + L.Aff[3 + i] * v[A + 1] + L.Aff[6 + i] * v[A + 2];
    y[V + i] = hd * v[P + i] + v[V + i];
    y[W + i] = L.Afw[i] * v[A] + L.Afw[3 + i] * v[A + 1] + L.Afw[6 + i] * v[A + 2] + v[W + i];
  }
}
template <class M>
QMPC_HD inline void coop_Mt_vec(const KnotLin4& L, double hd, double hh, const double* v, double* t) {
  constexpr int P = 3 * (0 ^ M::kSwap), A = 3 * (1 ^ M::kSwap), V = 3 * (2 ^ M::kSwap), W = 3 * (3 ^ M::kSwap);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    t[i] = hd * hh * v[P + i] + hd * v[V + i];
    if (M::kDw)
      t[3 + i] = L.Cf[i] * v[A] + L.Cf[3 + i] * v[A + 1] + L.Cf[6 + i] * v[A + 2] +
                 (L.Dw[i] * v[W] + L.Dw[3 + i] * v[W + 1] + L.Dw[6 + i] * v[W + 2]);
    else
      t[3 + i] = L.Cf[i] * v[A] + L.Cf[3 + i] * v[A + 1] + L.Cf[6 + i] * v[A + 2] + hd * v[W + i];
  }
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
  if (M::kQuat && cfg.w != 0.0) {
    constexpr int qi = M::kQuat ? 3 : 0;
    const double s = xr[qi] * x[qi] + xr[qi + 1] * x[qi + 1] + xr[qi + 2] * x[qi + 2] + xr[qi + 3] * x[qi + 3];
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
        const double lh = pos_part(est);
        if (c > viol) viol = c;
        acc += lh * lh - mui * mui;
      }
    }
    J += qmpc_div(acc, 2 * rho);
  }
}

// attitude block of the cost Hessian: G^T diag(Qq) G + hphi I (quaternion models) / diag(Q_rpy) (Euler model)
template <class M>
QMPC_HD inline void hphi_block(const QmpcConfig& cfg, const double* x, double hphi, double* H) {
  if (M::kQuat) {
    double Gq[12];
    quat_G(x + 3, Gq);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double s = 0;
        for (int i = 0; i < 4; ++i) s += Gq[3 * i + a] * cfg.q_weights[3 + i] * Gq[3 * i + b];
        H[3 * a + b] = s + (a == b ? hphi : 0.0);
      }
  } else {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) H[3 * a + b] = (a == b) ? cfg.q_weights[a] : 0.0;
  }
}

// 3x3 block (br, bc) of the cost Hessian in error coordinates (memory order of the blocks)
template <class M>
QMPC_HD inline void lxx_block(const double* wq, const double* Hphi, int br, int bc, double* out) {
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = 0.0;
  if (br != bc) return;
  if (br == (1 ^ M::kSwap)) {
#pragma unroll
    for (int i = 0; i < 9; ++i) out[i] = Hphi[i];
  } else {
    const int q0 = M::qoff(br);
    out[0] = wq[q0]; out[4] = wq[q0 + 1]; out[8] = wq[q0 + 2];
  }
}

#if defined(__CUDA_ARCH__)
// 16-byte cp.async requests dealt over the `nl` lanes of the problem (SASS LDGSTS)
__device__ __forceinline__ void coop_cp_async_row(double* dst, const double* src, int doubles, int tl, int nl) {
  for (int c = tl; 2 * c < doubles; c += nl) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + 2 * c);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + 2 * c) : "memory");
  }
}
#endif

// One roll-out of the whole horizon by ONE lane (kept out of line: it is used by the nominal
// roll-out and by the 16 speculative line-search lanes, and inlining it twice is what pushed the SASS
// far past the instruction cache).
//   mode 0: open loop, u = u_ref: writes the nominal X, U; returns merit / violation
//   mode 1: trial step `alpha` around (X, U) with gains (gK: N x [K_k | d_k]): X, U untouched; the trial
//           trajectory is recorded in the scratch (gTX/gTU, element-major, lane `tl` of `tstride`) so
//           that the accepted one is simply copied back - no second roll-out; returns merit / violation
template <class M, bool kSmem = true>
QMPC_HD QMPC_NOINLINE void coop_rollout(const M& m, const QmpcConfig& cfg, const double* wr, int N, float h, double* X,
                                        double* U, const double* gK,
                                        const double* gmu, double rho, double alpha, int mode, double* Jout,
                                        double* violout, double* gTX, double* gTU, int tl, int tstride,
                                        double* kstage, unsigned lane_mask, const QmpcWarmStart* winit) {
  // Compact by construction (instruction-fetch bound otherwise): the input never
  // exists as an array - each foot's force is formed, costed, cone-checked and folded into the net
  // wrench inside one NF-trip loop; the wrench drives both midpoint evaluations.  Accumulation
  // orders are exactly those of stage_cost() / knot_merit() / ct_dyn() / mid_dyn().
  constexpr int NX = M::NX, NE = 12, NU = M::NU, NC = M::NC, NF = M::kFeet;
  constexpr int kKD = NU * 12 + NU;
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_ASSUME_SHARED)
  // this function is kept out of line, so its pointers are generic: tell the compiler which ones are shared memory
  // (model, trajectories, weights - in the fused and forward kernels) so that it emits LDS / STS instead of generic
  // LD / ST (149 M generic loads per 16 384 solves in the roll-out otherwise, profiles/r02_ncu_coop_B16384.txt)
  if (kSmem) {
    __builtin_assume(__isShared(&m));
    __builtin_assume(__isShared(X));
    __builtin_assume(__isShared(U));
    __builtin_assume(__isShared(wr));
  }
#endif
  const double hd = (double)h, hh = (double)(h / 2);
  double x[NX], J = 0, vl = 0;
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = X[i];
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_KSTAGE)
  // Mode 1 runs on all `tstride` lanes of the problem in lock-step: the gain matrix (and feed-forward) of
  // knot k+1 is copied from the L2-resident scratch into a double buffer in shared memory while knot k is
  // being computed, so the 72 broadcast reads of K_k per lane are shared-memory reads instead of L2 round
  // trips.  16-byte cp.async requests dealt over the lanes (LDGSTS).  One cp.async.bulk per knot (TMA 1-D, mbarrier
  // completion, issued by lane 0) was measured in tools/probes/bulk_probe.cu: the same staging pattern runs 2.3x
  // SLOWER (1.47 ms against 0.63 ms), and in the kernel the gain stores of the backward pass would additionally
  // need a generic -> async proxy fence per knot (profiles/r02_experiments.md) - not adopted.
  auto stage_gain = [&](int k) {
    coop_cp_async_row(kstage + (k & 1) * kKD, gK + (size_t)k * kKD, kKD, tl, tstride);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto stage_wait = [&](int) { asm volatile("cp.async.wait_all;" ::: "memory"); };
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
    // ---- state part of the stage cost
    double Jl = 0, sq = 0;
    {
      double xr[NX];
      m.xref(k, xr);
#pragma unroll
      for (int i = 0; i < NX; ++i) { const double dxi = x[i] - xr[i]; Jl += 0.5 * cfg.q_weights[i] * dxi * dxi; }
      if (M::kQuat) sq = xr[3] * x[3] + xr[4] * x[4] + xr[5] * x[5] + xr[6] * x[6];
    }
    if (k == N) {
      if (M::kQuat && cfg.w != 0.0) Jl += cfg.w * (1.0 - fabs(sq));
      J += Jl;
      break;
    }
    // ---- per foot: force, input cost, cone rows / AL merit, wrench
    double mom0 = 0, mom1 = 0, mom2 = 0, fs0 = 0, fs1 = 0, fs2 = 0, acc = 0;
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_KSTAGE)
    const double* Kk = mode == 1 ? kstage + (k & 1) * kKD : gK;
    if (mode == 1) {
      stage_wait(k);
      __syncwarp(lane_mask);              // K_k visible to all lanes; everyone is done with K_{k-1}
      if (k + 1 < N) stage_gain(k + 1);
    }
#else
    const double* Kk = gK + (size_t)k * kKD;
#endif
    const double* dk = Kk + NU * 12;
#ifndef QMPC_COOP_CR_LDS
    double cr[18];   // cone rows in registers across the foot loop (the loop re-read them per foot: 72 LDS per knot; +1.6 %)
#if defined(__CUDA_ARCH__)
    if (kSmem && offsetof(M, CR) % 16 == 0) {   // 9 LDS.128 (the model sits 16-byte aligned in shared memory)
      const double2* cr2 = reinterpret_cast<const double2*>(m.CR);
#pragma unroll
      for (int i = 0; i < 9; ++i) { const double2 v = cr2[i]; cr[2 * i] = v.x; cr[2 * i + 1] = v.y; }
    } else
#endif
    {
#pragma unroll
      for (int i = 0; i < 18; ++i) cr[i] = m.CR[i];
    }
#else
    const double* cr = m.CR;
#endif
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_KSTAGE) && !defined(QMPC_COOP_NO_K128)
    const unsigned ks0 = (unsigned)__cvta_generic_to_shared(Kk);   // shared-window address of K_k, once per knot
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
        (void)K0;
#if !defined(QMPC_COOP_NO_K128) && defined(__CUDA_ARCH__)
        {   // three 96-byte gain rows as 18 x 16-byte loads (rows are 16-byte aligned)
#ifndef QMPC_COOP_NO_KSTAGE
          const unsigned ks = ks0 + (unsigned)(3 * f * 12 * sizeof(double));
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
      {
        const double d0 = u0, d1 = u1, d2 = u2 - m.urefz(k, f);   // u_ref = (0, 0, weight share)
        Jl += 0.5 * wr[3 * f] * d0 * d0;
        Jl += 0.5 * wr[3 * f + 1] * d1 * d1;
        Jl += 0.5 * wr[3 * f + 2] * d2 * d2;
        const double* mu_f = gmu + k * NC + 6 * f;
        const double fzc_f = m.fzc(k, f);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          double c = cr[3 * r] * u0 + cr[3 * r + 1] * u1 + cr[3 * r + 2] * u2;
          if (r == 4) c += -fzc_f;
          const double mui = mu_f[r];
          const double est = mui + rho * c;
          const double lh = pos_part(est);   // (a predicated fma instead of the two selects: ptxas if-converts it back)
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
    if (M::kQuat && cfg.w != 0.0) Jl += cfg.w * (1.0 - fabs(sq));
    J += Jl;
    J += qmpc_div(acc, 2 * rho);
    // ---- explicit midpoint step driven by the net wrench (AltroUtils.cpp:9-22, 383-391 / 224-294)
    m.wrench_step(x, fs0, fs1, fs2, mom0, mom1, mom2, hd, hh);
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) X[(k + 1) * NX + i] = x[i];
    }
  }
  *Jout = J;
  *violout = vl;
}

// ------------------------------------------------------------------------------------------------
// Per-problem context of the phase functions: where every array of the problem lives (shared memory or
// the L2-resident scratch - the phases do not care) and the solver scalars, uniform over the problem's lanes.
template <class M, int G>
struct CoopCtx {
  using L = CoopLayout<M, G>;
  M* m;
  double *X, *U, *P, *PA, *T, *PM, *S, *Qux, *vec, *lin, *red, *kstage, *dxs;
  double *gK, *gP, *gpv, *gmu, *glin, *DX, *gLX, *gTX, *gTU;
  int lin_stride_is_resident;   // 1: glin rows are read in place (shared memory); 0: staged into `lin` per knot
  const double *wq, *wr, *I3, *hI3;
  int N;
  float h;
  double hd, hh, c1;
  // solver state
  double rho, phi, viol, cost_decrease, dphi0;
  int status, iters;

  // fused kernel / backward kernel: the full shared-memory layout `sm`, scratch `gs`
  QMPC_HD void bind(double* sm, double* gs, double* trial, int N_, float h_, int flags, const double* wts) {
    N = N_; h = h_; hd = (double)h_; hh = (double)(h_ / 2); c1 = hd * hh;
    wq = wts; wr = wts + 13; I3 = wts + kCoopI3; hI3 = wts + kCoopHI3;
    m = reinterpret_cast<M*>(sm);
    X = sm + L::sX(N); U = sm + L::sU(N);
    P = sm + L::sP(N); PA = sm + L::sPA(N); T = sm + L::sT(N); PM = sm + L::sPM(N); S = sm + L::sS(N);
    Qux = sm + L::sQux(N); vec = sm + L::sVec(N); lin = sm + L::sLin(N);
    // 2 G reduction slots at the tail of T (only used outside the backward pass, where T is dead); the
    // roll-out's gain stage (2 kKD doubles + 2 mbarriers) occupies P, PA and the head of T meanwhile; the
    // accepted step's dx buffer ((N + 1) x 12 <= 396) overlays P .. Qux (612 doubles, all dead by then)
    red = T + 72 - 2 * G;
    kstage = P; dxs = P;
    static_assert(2 * L::kKD + 2 <= 288 + 72 - 2 * G, "gain stage overlaps the reduction slots");
    gK = gs + L::gK(N); gP = gs + L::gP(N); gpv = gs + L::gpv(N);
    gmu = (flags & 2) ? sm + L::sMu(N, flags) : gs + L::gmu(N);
    glin = (flags & 1) ? sm + L::sLinAll(N) : gs + L::glin(N);
    lin_stride_is_resident = flags & 1;
    DX = gs + L::gDX(N); gLX = gs + L::gLX(N);
    gTX = trial; gTU = trial + (size_t)(N + 1) * M::NX * G;
  }
  // forward kernel: model, X, U, stage, reduction slots in shared memory; everything else in the problem block
  QMPC_HD void bind_forward(double* sm, double* gs, double* trial, int N_, float h_, const double* wts) {
    N = N_; h = h_; hd = (double)h_; hh = (double)(h_ / 2); c1 = hd * hh;
    wq = wts; wr = wts + 13; I3 = wts + kCoopI3; hI3 = wts + kCoopHI3;
    m = reinterpret_cast<M*>(sm);
    X = sm + L::sX(N); U = sm + L::sU(N);
    P = PA = T = PM = S = Qux = vec = lin = nullptr;
    kstage = sm + L::fStage(N); dxs = kstage; red = sm + L::fRed(N);
    gK = gs + L::gK(N); gP = gs + L::gP(N); gpv = gs + L::gpv(N);
    gmu = gs + L::gmu(N); glin = gs + L::glin(N); lin_stride_is_resident = 0;
    DX = gs + L::gDX(N); gLX = gs + L::gLX(N);
    gTX = trial; gTU = trial + (size_t)(N + 1) * M::NX * G;
  }
};

#define COOP_ARGS_DECL int lane_id, unsigned lane_mask
#define COOP_ARGS lane_id, lane_mask

// ------------------------------------------------------------------ set-up + nominal roll-out
template <class M, int G>
QMPC_HD inline void coop_phase_setup(CoopCtx<M, G>& c, const QmpcConfig& cfg, const SolverOpts& o,
                                     const typename M::Problem* in, const unsigned char* sched, QmpcWarmStart* warm,
                                     int pid, COOP_ARGS_DECL) {
  constexpr int NC = M::NC;
  (void)lane_id; (void)lane_mask;
  const int N = c.N;
  COOP_PHASE {
#pragma unroll 1
    for (int i = lane; i < N * NC; i += G) c.gmu[i] = 0.0;
    if (lane == 0) {
      typename M::Problem prob = in[pid];
      c.m->setup(cfg, prob, sched ? sched + (size_t)pid * QMPC_MAX_HORIZON : nullptr, c.X);
    }
  }
  COOP_SYNC();
  c.rho = o.penalty_initial;
  double* scal = c.vec + CoopLayout<M, G>::vscal;
  COOP_PHASE {
    if (lane == 0) coop_rollout<M>(*c.m, cfg, c.wr, N, c.h, c.X, c.U, c.gK, c.gmu, c.rho, 0.0, 0, &scal[2], &scal[3], c.gTX, c.gTU,
                                   0, G, c.kstage, lane_mask, (warm && warm[pid].valid) ? warm + pid : nullptr);
  }
  COOP_SYNC();
  c.phi = scal[2]; c.viol = scal[3];
  c.status = QMPC_STATUS_MAX_ITERATIONS; c.iters = 0;
  c.cost_decrease = INFINITY; c.dphi0 = 0;
  if (!isfinite(c.phi)) c.status = QMPC_STATUS_NONFINITE;
}

// stationarity residuals of knot k with the Riccati duals y (DX): |lx + A^T y_{k+1} - y_k| and, per foot (rolled: once
// per iteration, compact code beats unrolled speed), |R (u - u_ref) + J^T max(0, mu + rho c) + W^T M^T y_{k+1}|
// (accumulation order of the AL terms)
template <class M>
QMPC_HD inline void coop_knot_residuals(const M& m, double rho, int N, int k, double hd, double hh, const double* U,
                                               const double* DX, const double* gmu, const double* wr, const double* lx,
                                               const KnotLin4& Lk, double& rx, double& ru) {
  constexpr int NE = 12, NU = M::NU, NC = M::NC, NF = M::kFeet;
  if (k == N) {
#pragma unroll
    for (int a = 0; a < NE; ++a) {
      double v = fabs(lx[a] - DX[N * NE + a]);
      if (v > rx) rx = v;
    }
    return;
  }
  const double* u = U + k * NU;
  const double* yn = DX + (k + 1) * NE;
  double Aty[NE], t6[6];
  coop_At_vec<M>(Lk, hd, yn, Aty);
  coop_Mt_vec<M>(Lk, hd, hh, yn, t6);
#pragma unroll
  for (int a = 0; a < NE; ++a) {
    double v = fabs(lx[a] + Aty[a] - DX[k * NE + a]);
    if (v > rx) rx = v;
  }
#pragma unroll 1
  for (int f = 0; f < NF; ++f) {
    const double* uf = u + 3 * f;
    const double* IS = m.IS + 9 * f;
    const double fzc_f = m.fzc(k, f);
    double g0 = 0, g1 = 0, g2 = 0;
#pragma unroll 1
    for (int r = 0; r < 6; ++r) {
      const double j0 = m.CR[3 * r], j1 = m.CR[3 * r + 1], j2 = m.CR[3 * r + 2];
      double cc = j0 * uf[0] + j1 * uf[1] + j2 * uf[2];
      if (r == 4) cc += -fzc_f;
      const double est = gmu[k * NC + 6 * f + r] + rho * cc;
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

// ------------------------------------------------------------------ per-iteration "pre": expansions,
// stationarity + convergence test, dual / penalty update, AL terms of every knot
template <class M, int G>
QMPC_HD inline void coop_phase_pre(CoopCtx<M, G>& c, const QmpcConfig& cfg, const SolverOpts& o, int it, COOP_ARGS_DECL) {
  using L = CoopLayout<M, G>;
  using Row = CoopRow<M>;
  constexpr int NX = M::NX, NE = 12, NU = M::NU, NC = M::NC, NF = M::kFeet, NLIN = M::NLIN;
  (void)lane_id; (void)lane_mask;
  const int N = c.N;
  const M& m = *c.m;
  double *X = c.X, *U = c.U, *gLX = c.gLX, *glin = c.glin, *gmu = c.gmu, *DX = c.DX, *red = c.red;
  const double* wr = c.wr;
  const double hd = c.hd, hh = c.hh;
  // ---------------- expansions, lane k <- knot k: cost gradient + attitude Hessian block (21 doubles) and
  // the dynamics blocks (NLIN doubles).  The same X is used throughout the iteration: computed once,
  // knot-parallel, reused by the stationarity test and the backward pass.
  // (Computing the stationarity residuals of knot k in this same pass - lx and the linearisation blocks are in
  // registers here, the separate pass below re-reads them - was measured no faster: 1.687 M against 1.700 M solves/s at
  // B = 4096, run 14.)
  COOP_PHASE {
#pragma unroll 1
    for (int k = lane; k <= N; k += G) {
      double lx[NE], Hk[9], hphi;
      KnotLin4 Lk = {};
      cost_expand(m, cfg, k, X + k * NX, lx, &hphi);
      hphi_block<M>(cfg, X + k * NX, hphi, Hk);
#pragma unroll
      for (int i = 0; i < NE; ++i) gLX[k * L::kRow + Row::lx + i] = lx[i];
#pragma unroll
      for (int i = 0; i < 9; ++i) gLX[k * L::kRow + Row::Hphi + i] = Hk[i];
      if (k == N) {   // the backward pass starts from these: hand them over in shared memory
#pragma unroll
        for (int i = 0; i < NE; ++i) c.vec[L::vpv + i] = lx[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) c.vec[L::vQx + i] = Hk[i];
      }
      if (k < N) {
        coop_linearize(m, X + k * NX, U + k * NU, X + (k + 1) * NX, hd, hh, Lk);
        double* gl = glin + k * L::kLinStride;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          gl[i] = Lk.Aff[i];
          gl[9 + i] = Lk.Afw[i];
          gl[18 + i] = Lk.Cf[i];
          if (NLIN > 27) gl[(NLIN > 27 ? 27 : 0) + i] = Lk.Dw[i];
        }
      }
    }
  }
  COOP_SYNC();

  if (it > 0) {
    {
      // ---------------- stationarity with the Riccati duals of the accepted step (DX holds y_k)
      COOP_PHASE {
        double rx = 0, ru = 0;
#pragma unroll 1
        for (int k = lane; k <= N; k += G) {
          double lx[NE];
          KnotLin4 Lk = {};   // (an uninitialised block reaching the residuals at k = N made ptxas spill)
#pragma unroll
          for (int a = 0; a < NE; ++a) lx[a] = gLX[k * L::kRow + Row::lx + a];
          if (k < N) {
            const double* gl = glin + k * L::kLinStride;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
              Lk.Aff[i] = gl[i];
              Lk.Afw[i] = gl[9 + i];
              Lk.Cf[i] = gl[18 + i];
              if (NLIN > 27) Lk.Dw[i] = gl[(NLIN > 27 ? 27 : 0) + i];
            }
          }
          coop_knot_residuals<M>(m, c.rho, N, k, hd, hh, U, DX, gmu, wr, lx, Lk, rx, ru);
        }
        red[lane] = rx > ru ? rx : ru;
      }
      COOP_SYNC();
    }
    double stat = 0;
    for (int l = 0; l < G; ++l) stat = red[l] > stat ? red[l] : stat;
    COOP_SYNC();
    if (stat < o.tol_stationarity && c.viol < o.tol_primal_feasibility) {
      c.status = QMPC_STATUS_SUCCESS;
      return;
    }
    if (fabs(c.cost_decrease) < o.tol_cost_intermediate || stat < o.tol_stationarity) {
      // dual update (row-parallel), penalty update, merit refresh (knot-parallel)
      COOP_PHASE {
#pragma unroll 1
        for (int idx = lane; idx < N * NC; idx += G) {
          const int k = idx / NC, r = idx % NC, f = r / 6, rr = r % 6;
          const double* u = U + k * NU + 3 * f;
          double cc = m.CR[3 * rr] * u[0] + m.CR[3 * rr + 1] * u[1] + m.CR[3 * rr + 2] * u[2];
          if (rr == 4) cc += -m.fzc(k, f);
          const double est = gmu[idx] + c.rho * cc;
          gmu[idx] = pos_part(est);
        }
      }
      COOP_SYNC();
      {
        const double r = c.rho * o.penalty_scaling;
        c.rho = r < o.penalty_max ? r : o.penalty_max;
      }
      COOP_PHASE {
        double J = 0, vl = 0;
#pragma unroll 1
        for (int k = lane; k <= N; k += G) knot_merit(m, cfg, wr, k, N, X + k * NX, U + k * NU, gmu + k * NC, c.rho, J, vl);
        red[lane] = J;
        red[G + lane] = vl;
      }
      COOP_SYNC();
      c.phi = 0;
      c.viol = 0;
      for (int l = 0; l < G; ++l) {
        c.phi += red[l];
        c.viol = red[G + l] > c.viol ? red[G + l] : c.viol;
      }
      COOP_SYNC();
    }
  }
  // ---------------- AL terms of every (knot, foot) with the duals / penalty now in force: input gradient
  // g = R (u - u_ref) + J^T max(0, mu + rho c) and D block = R + rho J_a^T J_a.  (knot, foot)-parallel over the
  // 16 lanes - off the critical path of the Riccati recursion, where 4 lanes used to compute them knot by knot.
  COOP_PHASE {
#pragma unroll 1
    for (int idx = lane; idx < N * NF; idx += G) {
      const int k = idx / NF, f = idx % NF;
      const double* u = U + k * NU + 3 * f;
      double g0 = 0, g1 = 0, g2 = 0, hb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
      for (int r = 0; r < 6; ++r) {
        double cc = m.CR[3 * r] * u[0] + m.CR[3 * r + 1] * u[1] + m.CR[3 * r + 2] * u[2];
        if (r == 4) cc += -m.fzc(k, f);
        const double est = gmu[k * NC + 6 * f + r] + c.rho * cc;
        if (est > 0) {
          const double j0 = m.CR[3 * r], j1 = m.CR[3 * r + 1], j2 = m.CR[3 * r + 2];
          g0 += j0 * est; g1 += j1 * est; g2 += j2 * est;
          hb[0] += c.rho * j0 * j0; hb[1] += c.rho * j0 * j1; hb[2] += c.rho * j0 * j2;
          hb[3] += c.rho * j1 * j0; hb[4] += c.rho * j1 * j1; hb[5] += c.rho * j1 * j2;
          hb[6] += c.rho * j2 * j0; hb[7] += c.rho * j2 * j1; hb[8] += c.rho * j2 * j2;
        }
      }
      hb[0] += wr[3 * f]; hb[4] += wr[3 * f + 1]; hb[8] += wr[3 * f + 2];
      double* row = gLX + k * L::kRow;
#pragma unroll
      for (int a = 0; a < 9; ++a) row[Row::Dblk + 9 * f + a] = hb[a];
      row[Row::g + 3 * f] = wr[3 * f] * u[0] + g0;
      row[Row::g + 3 * f + 1] = wr[3 * f + 1] * u[1] + g1;
      row[Row::g + 3 * f + 2] = wr[3 * f + 2] * (u[2] - m.urefz(k, f)) + g2;
    }
  }
  COOP_SYNC();
}

// ------------------------------------------------------------------ Riccati backward pass.  Lane (br, bc) =
// (lane / 4, lane % 4) owns the 3x3 block (br, bc) of every 12x12 quantity; all inner indices are compile-time.
// Sets c.dphi0; on a non-positive pivot c.status = BACKWARD_FAILED.
template <class M, int G>
QMPC_HD inline void coop_phase_backward(CoopCtx<M, G>& c, COOP_ARGS_DECL) {
  using L = CoopLayout<M, G>;
  using Row = CoopRow<M>;
  constexpr int NU = M::NU, NF = M::kFeet, NLIN = M::NLIN;
  // memory offsets of the state blocks by role (position, attitude, linear velocity, angular velocity)
  constexpr int oP = 3 * (0 ^ M::kSwap), oA = 3 * (1 ^ M::kSwap), oV = 3 * (2 ^ M::kSwap), oW = 3 * (3 ^ M::kSwap);
  static_assert(G == 16, "block-per-lane mapping assumes 16 lanes per problem");
  (void)lane_id; (void)lane_mask;
  const int N = c.N;
  const M& m = *c.m;
  double *P = c.P, *PA = c.PA, *T = c.T, *PM = c.PM, *S = c.S, *Qux = c.Qux, *vec = c.vec, *gLX = c.gLX;
  double *gK = c.gK, *gP = c.gP, *gpv = c.gpv;
  const double* wq = c.wq;
  const double hd = c.hd, hh = c.hh, c1 = c.c1;
  double* SW = PM;
  double* scal = vec + L::vscal;
  double* row = vec;                     // the knot row [lx | Hphi | g | Dblk] staged here
  double* pvv = vec + L::vpv;
  bool bp_ok = true;
  double* Pc = P;   // value-function Hessian of knot k+1 (then the not-yet-corrected one of knot k)
  double* Pw = PA;  // work buffer: P A, then Quu and its Cholesky factor, then the new P (ping-pong)
  // knot row + (when they are not shared-memory residents) the linearisation blocks of knot k: copied into
  // `vec` / `lin` with 16-byte cp.async requests, issued one knot AHEAD (after phase E of knot k+1, when the
  // row of knot k+1 is dead) so that the L2 round trip hides behind the Cholesky and phase F
  auto stage_row = [&](int k) {
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_ROW_PREFETCH)
    coop_cp_async_row(row, gLX + (size_t)k * L::kRow, L::kRow, lane_id, G);
    if (!c.lin_stride_is_resident) coop_cp_async_row(c.lin, c.glin + (size_t)k * L::kLinStride, L::kLinStride, lane_id, G);
    asm volatile("cp.async.commit_group;" ::: "memory");
#else
    COOP_PHASE {
#pragma unroll 1
      for (int e = lane; e < L::kRow; e += G) row[e] = gLX[(size_t)k * L::kRow + e];
      if (!c.lin_stride_is_resident) {
#pragma unroll 1
        for (int e = lane; e < NLIN; e += G) c.lin[e] = c.glin[(size_t)k * L::kLinStride + e];
      }
    }
#endif
  };
  auto stage_wait = [&]() {
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_ROW_PREFETCH)
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
    COOP_SYNC();
  };
  // terminal knot: its cost gradient (-> p_N) and attitude Hessian block were left in shared memory by the
  // expansions pass of this iteration (pvv, vec + vQx: `pre` and the backward pass always run in the same kernel), so
  // the row of knot N - 1 can be requested first and its L2 round trip overlaps the set-up of P_N
  stage_row(N - 1);
  COOP_PHASE {
    const int br = lane >> 2, bc = lane & 3;
    double o[9];
    lxx_block<M>(wq, vec + L::vQx, br, bc, o);
    blk_store(Pc + 36 * br + 3 * bc, 12, o);
    blk_store_keep(gP + (size_t)N * 144 + 36 * br + 3 * bc, 12, o);
    if (lane < 12) st_keep(gpv + N * 12 + lane, pvv[lane]);
    if (lane == 0) scal[0] = 0.0;
  }
  COOP_SYNC();

#pragma unroll 1
  for (int k = N - 1; k >= 0 && bp_ok; --k) {
    stage_wait();   // the row of knot k (and its linearisation blocks) are in shared memory
    const double* lin = c.lin_stride_is_resident ? c.glin + (size_t)k * L::kLinStride : c.lin;
    const double* Aff = lin;
    const double* Afw = lin + 9;
    const double* Cf = lin + 18;
    const double* Dw = lin + (NLIN > 27 ? 27 : 0);
    (void)Dw;
    const double* pv = pvv;
    // ---- phase B: PA = P A (16 blocks), PM = P M (8 blocks), s = M^T p, Atp = A^T p
    COOP_PHASE {
      const int br = lane >> 2, bc = lane & 3, rc = bc ^ M::kSwap;   // rc: role of this lane's block column
      // the two small vectors first: their load -> fma -> store chains then run under the block products' loads
#ifndef QMPC_COOP_VEC_DIVERGENT
      constexpr bool kUniformVec = !M::kDw;
#else
      constexpr bool kUniformVec = false;
#endif
      if (kUniformVec) {
        // s = M^T p (6 entries) and Atp = A^T p (12 entries): seven divergent one-line paths, each a dependent
        // LDS -> FMA -> STS chain, when written per role.  Every one of them is an instance of
        //   v = m0 p0 + m1 p1 + m2 p2 + beta pb  =  fma(beta, pb, fma(m2, p2, fma(m0, p0, m1 p1)))
        // with operands that are matrix entries, entries of p, or the constants 0 / 1 / h / c1 of the block's table
        // (a product by 1 or 0 is exact): one uniform evaluation for Atp (lanes 0..11), one for s (lanes 0..5),
        // rounding for rounding what the role-specific expressions evaluate.
        const double *zero = c.I3 + 1, *one = c.I3, *hp = c.hI3, *c1p = c.wq + kCoopC1;
        (void)hp;
        if (lane < 12) {
          const int a = lane, ab = (a / 3) ^ M::kSwap, aa = a % 3;   // ab: role of row a
          const bool odd = ab & 1;
          const double* Mx = (ab == 1 ? Aff : Afw) + aa;
          const double *m0 = odd ? Mx : (ab == 2 ? hp : zero), *m1 = odd ? Mx + 3 : one, *m2 = odd ? Mx + 6 : zero;
          const double *p0 = odd ? pv + oA : pv + oP + aa, *p1 = odd ? pv + oA + 1 : pv + a, *p2 = odd ? pv + oA + 2 : pv + a;
          const double beta = ab == 3 ? 1.0 : 0.0;
          vec[L::vAtp + a] = coop_dot4(*m0, *p0, *m1, *p1, *m2, *p2, beta, pv[a]);
        }
        if (lane < 6) {
          const int e = lane;
          // moment rows: Cf^T p_A + h p_W; force rows: c1 p_P + h p_V, which the role-specific code evaluates as
          // fma(h, p_V, c1 p_P) (the compiler shares the "+ h p" tail of the two branches): c1 p_P is the rounded product
          const bool mom = e >= 3;
          const double* Cx = Cf + (mom ? e - 3 : 0);
          const double *m0 = mom ? Cx : zero, *m1 = mom ? Cx + 3 : c1p, *m2 = mom ? Cx + 6 : zero;
          const double *p0 = mom ? pv + oA : pv + oP + e, *p1 = mom ? pv + oA + 1 : pv + oP + e, *p2 = mom ? pv + oA + 2 : pv + oP + e;
          const double *pb = mom ? pv + oW + e - 3 : pv + oV + e;
          vec[L::vs + e] = coop_dot4(*m0, *p0, *m1, *p1, *m2, *p2, hd, *pb);
        }
      } else {
        if (lane < 6) {
          const int e = lane;
          double v;
          if (e < 3) v = hd * hh * pv[oP + e] + hd * pv[oV + e];
          else if (M::kDw) v = Cf[e - 3] * pv[oA] + Cf[e] * pv[oA + 1] + Cf[3 + e] * pv[oA + 2] +
                               (Dw[e - 3] * pv[oW] + Dw[e] * pv[oW + 1] + Dw[3 + e] * pv[oW + 2]);
          else v = Cf[e - 3] * pv[oA] + Cf[e] * pv[oA + 1] + Cf[3 + e] * pv[oA + 2] + hd * pv[oW + e - 3];
          vec[L::vs + e] = v;
        }
        if (lane < 12) {
          const int a = lane, ab = (a / 3) ^ M::kSwap, aa = a % 3;   // ab: role of row a
          double v;
          if (ab == 0) v = pv[a];
          else if (ab == 1) v = Aff[aa] * pv[oA] + Aff[3 + aa] * pv[oA + 1] + Aff[6 + aa] * pv[oA + 2];
          else if (ab == 2) v = hd * pv[oP + aa] + pv[a];
          else v = Afw[aa] * pv[oA] + Afw[3 + aa] * pv[oA + 1] + Afw[6 + aa] * pv[oA + 2] + pv[a];
          vec[L::vAtp + a] = v;
        }
      }
      const double* Pr = Pc + 36 * br;
      {
        // ONE block product for all 16 lanes (see the comment above blk_right): the even roles are the odd roles'
        // product with the constant blocks I / h I and the two operands exchanged - same operations, same roundings
        const bool odd = rc & 1;
        blk_right(Pr, 12, odd ? oA : (rc == 0 ? oP : oV), odd ? oW : (rc == 0 ? oV : oP),
                  odd ? (rc == 1 ? Aff : Afw) : c.I3, rc == 2 ? hd : (rc == 3 ? 1.0 : 0.0), Pw + 36 * br + 3 * bc, 12);
        if (bc < 2) {   // P M: moment columns X_A Cf + h X_W, force columns c1 X_P + h X_V = X_V (h I) + c1 X_P
          if (!M::kDw)
            blk_right(Pr, 12, bc == 1 ? oA : oV, bc == 1 ? oW : oP, bc == 1 ? Cf : c.hI3, bc == 1 ? hd : c1,
                      PM + 18 * br + 3 * bc, 6);
          else if (bc == 1) blk_right2(Pr, 12, oA, oW, Cf, Dw, PM + 18 * br + 3 * bc, 6);   // Euler model: X_A Cf + X_W Dw
          else blk_even(Pr, 12, oP, oV, c1, hd, PM + 18 * br + 3 * bc, 6);
        }
      }
    }
    COOP_SYNC();
    // ---- phase C: P <- A^T PA + lxx (16 blocks), T = M^T PA (8 blocks), S = M^T PM (4 blocks), Qx
    COOP_PHASE {
      const int br = lane >> 2, bc = lane & 3, rr = br ^ M::kSwap;   // rr: role of this lane's block row
      if (lane < 12) vec[L::vQx + lane] = vec[L::vAtp + lane] + row[Row::lx + lane];
      const double* Yc = Pw + 3 * bc;
      double* Pd = Pc + 36 * br + 3 * bc;
      {   // A^T (P A) + lxx: one product for all lanes, the cost Hessian block (zero off the diagonal) folded in
        const bool odd = rr & 1;
        double lxx[9];
        lxx_block<M>(wq, row + Row::Hphi, br, bc, lxx);
        blk_left_add(Yc, 12, odd ? oA : (rr == 0 ? oP : oV), odd ? oW : (rr == 0 ? oV : oP),
                     odd ? (rr == 1 ? Aff : Afw) : c.I3, rr == 2 ? hd : (rr == 3 ? 1.0 : 0.0), lxx, Pd, 12);
      }
      if (lane < 12) {
        // T = M^T (P A) on lanes 0..7 and S = M^T (P M) on lanes 8..11 with per-lane operands
        const bool isS = lane >= 8;
        const int r = isS ? (lane >> 1) & 1 : br, cc = isS ? (lane & 1) : bc, ldy = isS ? 6 : 12;
        const double* Y = isS ? PM + 3 * cc : Yc;
        double* dst = isS ? S + 18 * r + 3 * cc : T + 36 * r + 3 * cc;
        if (!M::kDw) {
          // ONE block product: moment rows Cf^T Y_A + h Y_W, force rows c1 Y_P + h Y_V = (h I)^T Y_V + c1 Y_P
          blk_left(Y, ldy, r == 1 ? oA : oV, r == 1 ? oW : oP, r == 1 ? Cf : c.hI3, r == 1 ? hd : c1, dst, ldy);
        } else {
          // Euler model: the moment rows need the two-matrix product (Cf^T Y_A + Dw^T Y_W): one call per row type
          if (r == 1) blk_left2(Y, ldy, oA, oW, Cf, Dw, dst, ldy);
          else blk_evenT(Y, ldy, oP, oV, c1, hd, dst, ldy);
        }
      }
    }
    COOP_SYNC();
    // ---- phase D: Qux = W^T T (NF x 4 blocks), SW = S W (2 x NF blocks), Qu = g + W^T s
    COOP_PHASE {
      const int br = lane >> 2, bc = lane & 3;
      if (lane < NU) {
        const int f = lane / 3, a = lane % 3;
        const double* IS = m.IS + 9 * f;
        const double* s = vec + L::vs;
        vec[L::vQu + lane] = (m.inv_mass * s[a] + IS[a] * s[3] + IS[3 + a] * s[4] + IS[6 + a] * s[5]) + row[Row::g + lane];
      }
      if (br < NF) blk_wt(T + 3 * bc, 12, m.inv_mass, m.IS + 9 * br, Qux + 36 * br + 3 * bc, 12, nullptr);
      if (br < 2 && bc < NF) blk_w(S + 18 * br, 6, m.inv_mass, m.IS + 9 * bc, SW + 3 * NU * br + 3 * bc, NU);
    }
    COOP_SYNC();
    // ---- phase E: Quu = D + W^T (S W)  (NF x NF blocks, into the work buffer)
    double* Quu = Pw;
    COOP_PHASE {
      const int br = lane >> 2, bc = lane & 3;
      if (br < NF && bc < NF)
        blk_wt(SW + 3 * bc, NU, m.inv_mass, m.IS + 9 * br, Quu + 3 * NU * br + 3 * bc, NU,
               br == bc ? row + Row::Dblk + 9 * br : nullptr);
    }
    COOP_SYNC();
    if (k > 0) stage_row(k - 1);   // the row of knot k is dead from here on: bring the next one in meanwhile
    // ---- Cholesky + both triangular solves, fused, per lane, entirely in registers.  Every lane
    //      factors the NU x NU Quu redundantly (the kernel is latency- and shared-memory-bound, not
    //      FLOP-bound: 16 lanes doing the same flops cost the same issue slots as one) and then
    //      solves for its own right-hand side (column `lane` of Qux; lane 12: Qu) with L still in
    //      registers: no column exchange, no barrier, no shared-memory traffic for L.  Straight-line code; a
    //      non-positive pivot poisons the lane's result and is reported through `ok`.
    COOP_PHASE {
      const int cix = lane < 12 ? lane : 12;   // lanes 13..15 shadow the Qu column and store nothing
      constexpr int NT = NU * (NU + 1) / 2;
#define QMPC_TRI(i_, l_) ((i_) * ((i_) + 1) / 2 + (l_))
      double Lr[NT], rd[NU], rhs[NU];
#define QMPC_DIVD(x_, i_) { (x_) = (x_) * rd[i_]; }
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_CHOL_LD64)
      {   // lower triangle of Quu with 16-byte loads (rows start 16-byte aligned: NU is even): 42 LDS.128 for 78 LDS.64
        static_assert(NU % 2 == 0, "Quu rows must start 16-byte aligned");
        const double2* Q2 = reinterpret_cast<const double2*>(Quu);
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
          for (int l = 0; l <= i; l += 2) {
            const double2 v = Q2[(NU * i + l) / 2];
            Lr[QMPC_TRI(i, l)] = v.x;
            if (l + 1 <= i) Lr[QMPC_TRI(i, l + 1)] = v.y;
          }
      }
#else
#pragma unroll
      for (int i = 0; i < NU; ++i)
#pragma unroll
        for (int l = 0; l <= i; ++l) Lr[QMPC_TRI(i, l)] = Quu[NU * i + l];
#endif
#pragma unroll
      for (int i = 0; i < NU; ++i) rhs[i] = cix < 12 ? Qux[12 * i + cix] : vec[L::vQu + i];
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
        dg = fma(0.5 * fma(-dg, dg, sjj), r0, dg);          // sqrt(sjj), correctly rounded  (the exact factor 0.5 moved onto r0,
                                                            // off the pivot chain: measured 1 % SLOWER, run 20)
        const double rdg = fma(fma(-dg, r0, 1.0), r0, r0);  // 1 / dg
        rd[j] = rdg;
#pragma unroll
        for (int i = j + 1; i < NU; ++i) {
          // first guess and correction with r0 itself (within an ulp of 1 / dg): the correction step returns the
          // correctly rounded a / dg from either reciprocal, and the pivot chain no longer waits for rdg
          const double a = Lr[QMPC_TRI(i, j)], q = a * r0;
          Lr[QMPC_TRI(i, j)] = fma(fma(-q, dg, a), r0, q);   // a / dg
        }
#endif
#pragma unroll
        for (int i = j + 1; i < NU; ++i)
#pragma unroll
          for (int l = j + 1; l <= i; ++l) Lr[QMPC_TRI(i, l)] -= Lr[QMPC_TRI(i, j)] * Lr[QMPC_TRI(l, j)];
        // forward substitution riding along: column j of L is final, so y_j and its updates can go now (same
        // operations, same order per entry as a separate loop: bit-identical; +0.3 % measured)
        QMPC_DIVD(rhs[j], j);
#pragma unroll
        for (int l = j + 1; l < NU; ++l) rhs[l] -= Lr[QMPC_TRI(l, j)] * rhs[j];
      }
      if (!ok) bp_ok = false;
      if (ok && lane <= 12) {
        if (cix < 12) {
#pragma unroll
          for (int i = 0; i < NU; ++i) Qux[12 * i + cix] = rhs[i];   // V = L^-1 Qux
        } else {
#pragma unroll
          for (int i = 0; i < NU; ++i) vec[L::vvu + i] = rhs[i];
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
        double* gKk = gK + (size_t)k * L::kKD;
        if (cix < 12) {
#pragma unroll
          for (int i = 0; i < NU; ++i) st_keep(gKk + i * 12 + cix, -rhs[i]);
        } else {
          double t = 0;
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            st_keep(gKk + NU * 12 + i, -rhs[i]);
            t += vec[L::vQu + i] * (-rhs[i]);
          }
          scal[0] += t;
        }
      }
    }
    COOP_SYNC();
    if (!bp_ok) break;
    // ---- phase F: new P = sym(P) - V^T V (16 blocks, written to the work buffer: no race with the
    //      transposed reads of Pc), pv <- Qx - V^T vu ; then swap the two buffers
    COOP_PHASE {
      const int br = lane >> 2, bc = lane & 3;
      double o[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#ifndef QMPC_COOP_F_UNROLL
#define QMPC_COOP_F_UNROLL 4   // measured +0.4 % for 2 over 1, +0.2 % for 4 over 2, 12 loses (profiles/r02_experiments.md)
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
        for (int l = 0; l < NU; ++l) t += Qux[12 * l + a] * vec[L::vvu + l];
        const double v = vec[L::vQx + a] - t;
        pvv[a] = v;
        st_keep(gpv + k * 12 + a, v);
      }
    }
    COOP_SYNC();
    { double* t = Pc; Pc = Pw; Pw = t; }
  }
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_ROW_PREFETCH)
  if (!bp_ok) asm volatile("cp.async.wait_all;" ::: "memory");   // a failed knot leaves its successor's row in flight
#endif
  if (!bp_ok) c.status = QMPC_STATUS_BACKWARD_FAILED;
  c.dphi0 = scal[0];
}

// ------------------------------------------------------------------ forward pass: speculative back-tracking
// line search (lane l rolls out alpha = decrease^(round*G + l); the first lane passing the Armijo test wins -
// identical to the sequential search), then the accepted step.
template <class M, int G>
QMPC_HD inline void coop_phase_forward(CoopCtx<M, G>& c, const QmpcConfig& cfg, const SolverOpts& o, int it, COOP_ARGS_DECL) {
  constexpr int NX = M::NX, NE = 12, NU = M::NU, NCAND = G;
  (void)lane_id; (void)lane_mask;
  const int N = c.N;
  const M& m = *c.m;
  double *X = c.X, *U = c.U, *red = c.red, *DX = c.DX, *gTX = c.gTX, *gTU = c.gTU;
  int acc_j = -1;
  double phin = 0, violn = 0;
  // (prefetch.global.L2 of the P_k rows the accepted step will read - issued here, or right after the roll-out - was
  // measured 0.5 ... 0.7 % slower, run 23: those rows are not what the accepted step waits for)
#pragma unroll 1
  for (int round = 0; round * G < o.ls_iters_max && acc_j < 0; ++round) {
    COOP_PHASE {
      const int j = round * G + lane;
      // every lane rolls out (lanes past ls_iters_max too: the roll-out stages the gains
      // cooperatively); their result is discarded below
      double J = NAN, vl = 0, alpha = 1.0;
      for (int q = 0; q < j; ++q) alpha *= o.ls_decrease;
      coop_rollout<M>(m, cfg, c.wr, N, c.h, X, U, c.gK, c.gmu, c.rho, alpha, 1, &J, &vl, gTX, gTU, lane, G, c.kstage, lane_mask, nullptr);
      red[lane] = j < o.ls_iters_max ? J : NAN;
      red[G + lane] = vl;
    }
    COOP_SYNC();
    {
      double alpha = 1.0;
      for (int q = 0; q < round * G; ++q) alpha *= o.ls_decrease;
      for (int l = 0; l < G && acc_j < 0; ++l) {
        const double pl = red[l];
        if (round * G + l < o.ls_iters_max && isfinite(pl) && pl <= c.phi + o.ls_c1 * alpha * c.dphi0) {
          acc_j = round * G + l;
          phin = pl;
          violn = red[G + l];
        }
        alpha *= o.ls_decrease;
      }
    }
    COOP_SYNC();
  }
  c.iters = it + 1;
  if (acc_j < 0) { c.status = QMPC_STATUS_LINESEARCH_FAILED; return; }
  // ---------------- accepted step: the winning lane's trial trajectory is already in the scratch.
  const int acc_lane = acc_j % NCAND;
  // (1) lane k <- knot k: dx_k = x_new (-) x_old into the shared dx buffer; (2) Riccati duals y_k = P_k dx_k + p_k
  // with lane a <- ROW a of every knot: the 12 active lanes read 12 consecutive rows of P_k (1152 contiguous
  // bytes per knot, 16-byte loads), four knots' loads in flight before the first FMA - three L2 round trips for
  // the whole horizon where lane-per-knot needed a dozen (the accept step was 12.8 % of the solve's time for 4 %
  // of its instructions, long-scoreboard 8.6).  Row sums run b = 0..11 as before: bit-identical.
  // (Dealing the 12 (N + 1) rows over all 16 lanes, two batches with the first one's loads issued before step (1),
  // measured 1.4 % SLOWER, run 15: not adopted.)
  double* dxs = c.dxs;
  // after the LAST iteration only the inputs are read again (result, warm-start buffer): no dx, no new X, no duals
  const bool more = it + 1 < o.iterations_max;
  COOP_PHASE {
#pragma unroll 1
    for (int k = lane; more && k <= N; k += G) {
      double xn[NX], dx[NE];
#pragma unroll
      for (int i = 0; i < NX; ++i) xn[i] = ld_stream(gTX + (size_t)(k * NX + i) * NCAND + acc_lane);
      state_diff<M>(xn, X + k * NX, dx);
#pragma unroll
      for (int i = 0; i < NE; ++i) dxs[k * NE + i] = dx[i];
      // X <- accepted states right here: the lane holds its knot's new state in registers and nobody reads the old
      // one again (the copy loop at the end took one L2 round trip per element: `unroll 1`, load -> store)
#pragma unroll
      for (int i = 0; i < NX; ++i) X[k * NX + i] = xn[i];
    }
    // U <- accepted inputs, four independent loads in flight per lane
#pragma unroll 1
    for (int e0 = 0; e0 < N * NU; e0 += 4 * G) {
      double t[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = e0 + q * G + lane;
        t[q] = ld_stream(gTU + (size_t)(e < N * NU ? e : N * NU - 1) * NCAND + acc_lane);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = e0 + q * G + lane;
        if (e < N * NU) U[e] = t[q];
      }
    }
  }
  COOP_SYNC();
  COOP_PHASE {
    // (the duals are only read by the NEXT iteration's stationarity test)
    if (lane < NE && more) {
      const int a = lane;
#ifndef QMPC_COOP_ACCEPT_KB
#define QMPC_COOP_ACCEPT_KB 4
#endif
      constexpr int KB = QMPC_COOP_ACCEPT_KB;
#pragma unroll 1
      for (int k0 = 0; k0 <= N; k0 += KB) {
        double2 pr[KB][6];
        double t[KB];
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          const int k = k0 + q <= N ? k0 + q : N;   // clamped: a redundant load instead of a branch
          const double2* rowp = reinterpret_cast<const double2*>(c.gP + (size_t)k * 144 + 12 * a);
#pragma unroll
          for (int b = 0; b < 6; ++b) pr[q][b] = rowp[b];
          t[q] = ld_keep(c.gpv + k * 12 + a);
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          const int k = k0 + q <= N ? k0 + q : N;
          const double* dx = dxs + k * NE;
#pragma unroll
          for (int b = 0; b < 6; ++b) { t[q] += pr[q][b].x * dx[2 * b]; t[q] += pr[q][b].y * dx[2 * b + 1]; }
        }
#pragma unroll
        for (int q = 0; q < KB; ++q)
          if (k0 + q <= N) DX[(k0 + q) * NE + a] = t[q];
      }
    }
  }
  COOP_SYNC();
#if defined(__CUDA_ARCH__) && !defined(QMPC_COOP_NO_DISCARD)
  // The 16 trial trajectories (33.7 KB per slot at N = 10) are dead from here on, and they were 78 % of the DRAM
  // traffic of round 1's kernel (769 kB per solve against 536 B algorithmic): dirty L2 lines written back to HBM when
  // the scratch of 2368 slots (169 MB) overflows the 126 MB L2.  discard.global.L2 drops the lines WITHOUT write-back
  // (every line is rewritten in full by the next line search: 16 lanes x 8 bytes = one 128-byte line per element),
  // so only the live 37 KB per slot compete for the L2.
  {
    const size_t nlines = ((size_t)(N + 1) * NX * G + (size_t)N * NU * G + 15) / 16;
    for (size_t e = (size_t)lane_id; e < nlines; e += G)
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(gTX + e * 16) : "memory");
  }
#endif
  c.cost_decrease = c.phi - phin;
  c.phi = phin;
  c.viol = violn;
}

// ------------------------------------------------------------------ result + warm-start buffer
template <class M, int G>
QMPC_HD inline void coop_phase_epilogue(CoopCtx<M, G>& c, QmpcResult* out, QmpcWarmStart* warm, int pid, double* stage30,
                                        COOP_ARGS_DECL) {
  constexpr int NU = M::NU;
  (void)lane_id; (void)lane_mask;
  const int N = c.N;
  // lane 0 assembles the 240-byte result in shared memory, the 16 lanes store it as 30 coalesced 8-byte words
  // (one lane issuing 30 dependent global stores sat at 4 % of the solve's time in round 1's profile)
  COOP_PHASE {
    if (lane == 0) {
      QmpcResult& r = *reinterpret_cast<QmpcResult*>(stage30);
      c.m->write_result(c.U, r);
      r.max_violation = c.viol;
      r.iterations = c.iters;
      r.status = c.status;
    }
  }
  COOP_SYNC();
  COOP_PHASE {
    static_assert(sizeof(QmpcResult) % 8 == 0, "QmpcResult is copied as 8-byte words");
    double* dst = reinterpret_cast<double*>(out + pid);
#pragma unroll 1
    for (int e = lane; e < (int)(sizeof(QmpcResult) / 8); e += G) dst[e] = stage30[e];
    if (warm) {
      if (lane == 0) warm[pid].valid = c.status != QMPC_STATUS_NONFINITE;
#pragma unroll 1
      for (int e = lane; e < N * 12; e += G) {
        const int k = e / 12, i = e % 12;
        warm[pid].u[k][i] = i < NU ? c.U[k * NU + i] : 0.0;
      }
    }
  }
  COOP_SYNC();
}

// ------------------------------------------------------------------ the whole solve, fused (one persistent kernel)
template <class M, int G>
QMPC_HD void coop_solve_one(const QmpcConfig& cfg, const SolverOpts& o, const typename M::Problem* in,
                            const unsigned char* sched, QmpcWarmStart* warm, QmpcResult* out,
                            int pid, double* sm, double* gs, int lane_id, unsigned lane_mask, int flags,
                            const double* wts) {
  using L = CoopLayout<M, G>;
  CoopCtx<M, G> c;
  c.bind(sm, gs, gs + L::gTX(o.N), o.N, o.h, flags, wts);
#ifndef QMPC_COOP_DX_GLOBAL
  // Riccati duals y_k of the accepted step (written by the accepted step, read by the next iteration's stationarity
  // test): for short horizons they fit behind the dx buffer in the P .. Qux region, which is dead from the end of the
  // backward pass to the start of the next one - shared memory instead of two L2 round trips per iteration.
  // Layout while they live: dx buffer [0, 12 (N+1)) | y [156, 156 + 12 (N+1)) | reduction slots [328, 360)
  if (o.N <= 12) c.DX = c.P + 156;
#endif
  coop_phase_setup<M, G>(c, cfg, o, in, sched, warm, pid, COOP_ARGS);
#pragma unroll 1
  for (int it = 0; it < o.iterations_max; ++it) {
    // block-level phase alignment: every thread of the block passes this barrier exactly iterations_max times
    // per problem wave, finished slots included (see the header comment)
    COOP_BLOCK_SYNC();
    if (c.status != QMPC_STATUS_MAX_ITERATIONS) {
#if defined(__CUDA_ARCH__) && defined(QMPC_COOP_BLOCK_SYNC)
      continue;   // finished: keep passing the barriers
#else
      break;
#endif
    }
    coop_phase_pre<M, G>(c, cfg, o, it, COOP_ARGS);
    if (c.status != QMPC_STATUS_MAX_ITERATIONS) continue;
    coop_phase_backward<M, G>(c, COOP_ARGS);
    if (c.status != QMPC_STATUS_MAX_ITERATIONS) continue;
    coop_phase_forward<M, G>(c, cfg, o, it, COOP_ARGS);
  }
  coop_phase_epilogue<M, G>(c, out, warm, pid, c.P, COOP_ARGS);
}

#ifdef __CUDACC__
// persistent launch: every group of G lanes is a "slot" that strides over the batch
#ifndef QMPC_COOP_BLOCK
#define QMPC_COOP_BLOCK 128        // 4 warps = 8 problems per block
#endif
#ifndef QMPC_COOP_MIN_BLOCKS
#define QMPC_COOP_MIN_BLOCKS (256 / QMPC_COOP_BLOCK)   // 8 warps per SM at 255 registers
#endif
template <class M, int G>
__global__ void __launch_bounds__(QMPC_COOP_BLOCK, QMPC_COOP_MIN_BLOCKS)
qmpc_coop_kernel(QmpcConfig cfg, SolverOpts o, const typename M::Problem* __restrict__ in,
                 const unsigned char* __restrict__ sched, QmpcWarmStart* __restrict__ warm,
                 QmpcResult* __restrict__ out,
                 double* __restrict__ scratch, int batch, int smem_per_problem, size_t scratch_per_slot, int wide,
                 int active_groups) {
  extern __shared__ __align__(16) double smem_pool[];
  if (threadIdx.x < kCoopBlockShared) smem_pool[threadIdx.x] = coop_block_const(cfg, o.h, threadIdx.x);
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
      coop_solve_one<M, G>(cfg, o, in, sched, warm, out, pid, sm, gs, lane_id, lane_mask, wide, smem_pool);
    } else {
      for (int it = 0; it < o.iterations_max; ++it) { COOP_BLOCK_SYNC(); }
    }
  }
#else
  for (int pid = slot; !idle && pid < batch; pid += nslots) {
    coop_solve_one<M, G>(cfg, o, in, sched, warm, out, pid, sm, gs, lane_id, lane_mask, wide, smem_pool);
  }
#endif
}
#endif

}  // namespace qmpc
