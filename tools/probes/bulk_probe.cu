// bulk_probe.cu — isolates the gain-staging pattern of coop_rollout (qmpc_coop.cuh): per 16-lane group a double
// buffer in shared memory filled by ONE cp.async.bulk per knot (mbarrier completion) vs 16-byte cp.async per lane.
// Two groups per warp with independent barriers, re-initialised per "roll-out".  Prints PASS/FAIL + timings.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
constexpr int kKD = 156, N = 10, G = 16;
__device__ __forceinline__ void mbar_init(double* mb, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(mb);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, unsigned bytes, double* mb) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst), b = (unsigned)__cvta_generic_to_shared(mb);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(double* mb, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(mb);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(b), "r"(parity) : "memory");
}
// mode 4: the exact instruction sequence libcu++ emits for cuda::device::memcpy_async_tx + barrier_arrive_tx + wait
// (copy first, then arrive.expect_tx returning a token, token-based try_wait, source through cvta.to.global)
__device__ __forceinline__ unsigned long long bulk_g2s_tok(double* dst, const double* src, unsigned bytes, double* mb) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst), b = (unsigned)__cvta_generic_to_shared(mb);
  const unsigned long long g = (unsigned long long)__cvta_generic_to_global(src);
  unsigned long long tok;
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(g), "r"(bytes), "r"(b) : "memory");
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 %0, [%1], %2;" : "=l"(tok) : "r"(b), "r"(bytes) : "memory");
  return tok;
}
__device__ __forceinline__ void mbar_wait_tok(double* mb, unsigned long long tok) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(mb);
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.shared.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "l"(tok) : "memory");
  }
}
template <int MODE>   // 0 = bulk (lane 0 of BOTH half-warps in one predicated instruction), 1 = ldgsts,
                      // 2 = bulk, the two half-warps issue from different code paths, 3 = bulk, only half-warp 0 works
__global__ void __launch_bounds__(128) stage_kernel(const double* __restrict__ gK, double* out, int rollouts, int nprob) {
  extern __shared__ __align__(16) double sm[];
  const int group = threadIdx.x / G, tl = threadIdx.x % G;
  const unsigned lane_mask = ((1u << G) - 1u) << ((threadIdx.x % 32) / G * G);
  double* kstage = sm + (size_t)group * (2 * kKD + 2);
  double* mbar = kstage + 2 * kKD;
  const int pid = blockIdx.x * (blockDim.x / G) + group;
  if (pid >= nprob) return;
  if ((MODE == 3 || MODE == 4) && ((threadIdx.x % 32) / G) != 0) return;
  const double* src = gK + (size_t)pid * N * kKD;
  double acc = 0;
  for (int r = 0; r < rollouts; ++r) {
    if (MODE != 1) {
      if (tl == 0) {
        mbar_init(mbar, 1); mbar_init(mbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncwarp(lane_mask);
    }
    auto stage = [&](int k) {
      if (MODE == 0 || MODE == 3) { if (tl == 0) bulk_g2s(kstage + (k & 1) * kKD, src + (size_t)k * kKD, kKD * 8u, mbar + (k & 1)); }
      else if (MODE == 2) {
        if (((threadIdx.x % 32) / G) == 0) { if (tl == 0) bulk_g2s(kstage + (k & 1) * kKD, src + (size_t)k * kKD, kKD * 8u, mbar + (k & 1)); }
        else { if (tl == 0) bulk_g2s(kstage + (k & 1) * kKD, src + (size_t)k * kKD, kKD * 8u, mbar + (k & 1)); }
      } else {
        for (int c = tl; 2 * c < kKD; c += G) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(kstage + (k & 1) * kKD + 2 * c);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + (size_t)k * kKD + 2 * c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    };
    if (MODE == 4) {
      for (int k = 0; k < N; ++k) {
        if (tl == 0) {
          const unsigned long long tok = bulk_g2s_tok(kstage + (k & 1) * kKD, src + (size_t)k * kKD, kKD * 8u, mbar + (k & 1));
          mbar_wait_tok(mbar + (k & 1), tok);
        }
        __syncwarp(lane_mask);
        const double* Kk = kstage + (k & 1) * kKD;
        for (int i = tl; i < kKD; i += G) acc += Kk[i] * (1 + k);
        __syncwarp(lane_mask);
      }
      continue;
    }
    if (MODE == 5) {   // mode 4's recipe, pipelined one knot ahead as the kernel does it
      unsigned long long tok[2] = {0, 0};
      if (tl == 0) tok[0] = bulk_g2s_tok(kstage, src, kKD * 8u, mbar);
      for (int k = 0; k < N; ++k) {
        if (tl == 0) mbar_wait_tok(mbar + (k & 1), tok[k & 1]);
        __syncwarp(lane_mask);
        if (k + 1 < N && tl == 0) tok[(k + 1) & 1] = bulk_g2s_tok(kstage + ((k + 1) & 1) * kKD, src + (size_t)(k + 1) * kKD, kKD * 8u, mbar + ((k + 1) & 1));
        const double* Kk = kstage + (k & 1) * kKD;
        for (int i = tl; i < kKD; i += G) acc += Kk[i] * (1 + k);
      }
      __syncwarp(lane_mask);
      continue;
    }
    stage(0);
    for (int k = 0; k < N; ++k) {
      if (MODE != 1) mbar_wait(mbar + (k & 1), (unsigned)((k >> 1) & 1));
      else asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp(lane_mask);
      if (k + 1 < N) stage(k + 1);
      const double* Kk = kstage + (k & 1) * kKD;
      for (int i = tl; i < kKD; i += G) acc += Kk[i] * (1 + k);
    }
    __syncwarp(lane_mask);
  }
  out[(size_t)pid * G + tl] = acc;
}
template <int MODE>
static int run(const double* d, double* o0, double* o1, int nprob, int rollouts, size_t smem) {
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  stage_kernel<1><<<nprob / 8, 128, smem>>>(d, o1, rollouts, nprob);
  cudaError_t er = cudaDeviceSynchronize();
  printf("ldgsts: %s\n", cudaGetErrorString(er));
  cudaEventRecord(e0);
  stage_kernel<1><<<nprob / 8, 128, smem>>>(d, o1, rollouts, nprob);
  cudaEventRecord(e1);
  stage_kernel<MODE><<<nprob / 8, 128, smem>>>(d, o0, rollouts, nprob);
  cudaEventRecord(e2);
  er = cudaDeviceSynchronize();
  printf("bulk mode %d: %s\n", MODE, cudaGetErrorString(er));
  if (er != cudaSuccess) return 1;
  float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
  std::vector<double> r0((size_t)nprob * G), r1((size_t)nprob * G);
  cudaMemcpy(r0.data(), o0, r0.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(r1.data(), o1, r1.size() * 8, cudaMemcpyDeviceToHost);
  size_t bad = 0, cmp = 0;
  for (size_t i = 0; i < r0.size(); ++i) { if ((MODE == 3 || MODE == 4) && ((i / G) & 1)) continue; ++cmp; bad += r0[i] != r1[i]; }
  printf("{\"mode\": %d, \"ldgsts_ms\": %.3f, \"bulk_ms\": %.3f, \"compared\": %zu, \"mismatches\": %zu, \"verdict\": \"%s\"}\n", MODE, a, b, cmp, bad, bad ? "FAIL" : "PASS");
  return bad != 0;
}
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int nprob = 2048 * 8, rollouts = 20;
  std::vector<double> h((size_t)nprob * N * kKD);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)((i * 2654435761u) % 1000) * 1e-3;
  double *d, *o0, *o1;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&o0, (size_t)nprob * G * 8); cudaMalloc(&o1, (size_t)nprob * G * 8);
  cudaMemset(o0, 0, (size_t)nprob * G * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  const size_t smem = 8 * (2 * kKD + 2) * 8;
  if (mode == 2) return run<2>(d, o0, o1, nprob, rollouts, smem);
  if (mode == 3) return run<3>(d, o0, o1, nprob, rollouts, smem);
  if (mode == 4) return run<4>(d, o0, o1, nprob, rollouts, smem);
  if (mode == 5) return run<5>(d, o0, o1, nprob, rollouts, smem);
  return run<0>(d, o0, o1, nprob, rollouts, smem);
}
