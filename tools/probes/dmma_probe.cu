// dmma_probe.cu — evidence for the tensor-core decision of the QuatMpc solve (north_star: "tensor cores used only
// for the dense contraction, each choice evidenced by ncu").  FP64 on sm_100a has no tcgen05 path (tcgen05.mma kinds
// are f16/tf32/f8f6f4/i8/mxf*); the only FP64 tensor instruction is the legacy warp-synchronous
// mma.sync.aligned.m8n8k4.f64 (DMMA).  This probe measures, on the device:
//   1. DFMA peak and dependent-chain latency (vector pipe)
//   2. DMMA m8n8k4 peak (8 independent accumulators per warp) and dependent-chain latency
//   3. the one dense product of the Riccati step, P <- P - V^T V (12x12x12, phase F of qmpc_coop.cuh), batched with
//      operands in shared memory, as (a) the coop kernel does it - 16 lanes per problem, lane = 3x3 block, DFMA - and
//      (b) with DMMA - one warp per problem, 12x12 padded to 16x16, 2x2 tiles x 3 k-steps = 12 DMMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dmma_probe dmma_probe.cu ; prints one JSON line.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x + i;
#pragma unroll 1
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int r = 0; r < 64 / CHAINS; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double c0[CHAINS], c1[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
#pragma unroll 1
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int r = 0; r < 16 / CHAINS; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) dmma(c0[i], c1[i], a, b);
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- phase F shape: Pn = 0.5 (P + P^T) - V^T V, 12x12, V and P in shared memory, REPS times per problem
constexpr int kProblemDoubles = 288;   // P (144) + V (144)
__global__ void __launch_bounds__(128) phaseF_dfma(const double* in, double* out, int reps) {
  extern __shared__ double sm[];
  const int group = threadIdx.x / 16, lane = threadIdx.x % 16;
  double* P = sm + (size_t)group * kProblemDoubles;
  double* V = P + 144;
  const size_t pid = (size_t)blockIdx.x * (blockDim.x / 16) + group;
  for (int e = lane; e < kProblemDoubles; e += 16) P[e] = in[pid * kProblemDoubles + e];
  __syncwarp();
  const int br = lane >> 2, bc = lane & 3;
  double o[9];
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i] = 0;
#pragma unroll 1
    for (int l = 0; l < 12; ++l) {
      const double* Vr = V + 12 * l + 3 * br;
      const double* Vc = V + 12 * l + 3 * bc;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) o[3 * a + b] += Vr[a] * Vc[b];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        o[3 * a + b] = 0.5 * (P[12 * (3 * br + a) + 3 * bc + b] + P[12 * (3 * bc + b) + 3 * br + a]) - o[3 * a + b];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) P[12 * (3 * br + a) + 3 * bc + b] = o[3 * a + b];
    __syncwarp();
  }
  for (int e = lane; e < 144; e += 16) out[pid * 144 + e] = P[e];
}

// one warp per problem; V^T V as C(16x16) = A(16x12) B(12x16) with A = V^T (row-major fragment = column of V),
// B = V; 4 output tiles of 8x8, 3 k-steps of 4.  Lane layout of m8n8k4.f64: A[row = lane/4][k = lane%4],
// B[k = lane%4][col = lane/4], C[row = lane/4][col = 2*(lane%4) + {0,1}].
__global__ void __launch_bounds__(128) phaseF_dmma(const double* in, double* out, int reps) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  double* P = sm + (size_t)warp * kProblemDoubles;
  double* V = P + 144;
  const size_t pid = (size_t)blockIdx.x * (blockDim.x / 32) + warp;
  for (int e = lane; e < kProblemDoubles; e += 32) P[e] = in[pid * kProblemDoubles + e];
  __syncwarp();
  const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0;
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
      const int k = 4 * ks + t;
      // A fragment for row tile mt: (V^T)[8 mt + g][k] = V[k][8 mt + g] ; B fragment for column tile nt: V[k][8 nt + g]
      const double v0 = V[12 * k + g];                       // columns 0..7
      const double v1 = (8 + g) < 12 ? V[12 * k + 8 + g] : 0.0;   // columns 8..15 (12..15 are padding)
      dmma(c[0][0], c[0][1], v0, v0);
      dmma(c[1][0], c[1][1], v0, v1);
      dmma(c[2][0], c[2][1], v1, v0);
      dmma(c[3][0], c[3][1], v1, v1);
    }
    double pn[4][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int row = 8 * mt + g, col = 8 * nt + 2 * t + j;
          pn[2 * mt + nt][j] = (row < 12 && col < 12) ? 0.5 * (P[12 * row + col] + P[12 * col + row]) - c[2 * mt + nt][j] : 0.0;
        }
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int row = 8 * mt + g, col = 8 * nt + 2 * t + j;
          if (row < 12 && col < 12) P[12 * row + col] = pn[2 * mt + nt][j];
        }
    __syncwarp();
  }
  for (int e = lane; e < 144; e += 32) out[pid * 144 + e] = P[e];
}

template <class F>
static double time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  const int iters = 2048;
  const double clk_ghz = prop.clockRate * 1e-6;
  // peaks: 8 blocks x 256 threads per SM
  double ms = time_ms([&] { dfma_kernel<16><<<sms * 8, 256>>>(out, iters, 0.999999, 1e-6); });
  const double dfma_tf = 2.0 * 64 * iters * (double)sms * 8 * 256 / (ms * 1e-3) / 1e12;
  ms = time_ms([&] { dmma_kernel<8><<<sms * 8, 256>>>(out, iters, 0.999999, 1e-6); });
  const double dmma_tf = 2.0 * 256 * 16 * iters * (double)sms * 8 * 8 / (ms * 1e-3) / 1e12;   // 256 FMA per warp-level DMMA
  // latencies: one warp per SM, one chain
  ms = time_ms([&] { dfma_kernel<1><<<sms, 32>>>(out, iters, 0.999999, 1e-6); });
  const double dfma_lat = ms * 1e-3 * clk_ghz * 1e9 / (64.0 * iters);
  ms = time_ms([&] { dmma_kernel<1><<<sms, 32>>>(out, iters, 0.999999, 1e-6); });
  const double dmma_lat = ms * 1e-3 * clk_ghz * 1e9 / (16.0 * iters);
  // phase F: 16384 problems, 64 repetitions each, occupancy as in the coop kernel (8 warps per SM)
  const int nprob = 16384, reps = 64;
  std::vector<double> h((size_t)nprob * kProblemDoubles);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 1e-3 * ((i * 2654435761u) % 1000) / 1000.0;
  double *din, *dout, *dout2;
  cudaMalloc(&din, h.size() * 8); cudaMalloc(&dout, (size_t)nprob * 144 * 8); cudaMalloc(&dout2, (size_t)nprob * 144 * 8);
  cudaMemcpy(din, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  const size_t smA = 8 * kProblemDoubles * 8, smB = 4 * kProblemDoubles * 8;
  const double msA = time_ms([&] { phaseF_dfma<<<nprob / 8, 128, smA>>>(din, dout, reps); });
  const double msB = time_ms([&] { phaseF_dmma<<<nprob / 4, 128, smB>>>(din, dout2, reps); });
  std::vector<double> ra((size_t)nprob * 144), rb((size_t)nprob * 144);
  cudaMemcpy(ra.data(), dout, ra.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(rb.data(), dout2, rb.size() * 8, cudaMemcpyDeviceToHost);
  double maxd = 0; size_t nbit = 0;
  for (size_t i = 0; i < ra.size(); ++i) { double d = ra[i] - rb[i]; if (d < 0) d = -d; if (d > maxd) maxd = d; nbit += ra[i] != rb[i]; }
  const double useful = 2.0 * 1728 * (double)nprob * reps;
  printf("{\"device\": \"%s\", \"sm_clock_ghz\": %.3f, \"dfma_peak_tflops\": %.2f, \"dmma_m8n8k4_peak_tflops\": %.2f, "
         "\"dfma_latency_cycles\": %.1f, \"dmma_latency_cycles\": %.1f, "
         "\"phaseF_12x12x12\": {\"problems\": %d, \"reps\": %d, \"dfma_16lanes_ms\": %.3f, \"dmma_warp_ms\": %.3f, "
         "\"dfma_useful_tflops\": %.2f, \"dmma_useful_tflops\": %.2f, \"max_abs_diff\": %.3e, \"entries_not_bit_identical\": %zu, "
         "\"note\": \"DMMA pads 12 to 16: 12 x 256 = 3072 FMA slots for 1728 useful\"}, "
         "\"cudaError\": \"%s\"}\n",
         prop.name, clk_ghz, dfma_tf, dmma_tf, dfma_lat, dmma_lat, nprob, reps, msA, msB, useful / (msA * 1e-3) / 1e12,
         useful / (msB * 1e-3) / 1e12, maxd, nbit, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
