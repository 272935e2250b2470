// fma_peak.cu — measures the FP64 / FP32 CUDA-core FMA throughput of the device.
//
// MEASURED_PEAKS.json (driver-written) only has HBM GB/s and bf16 tensor TFLOP/s.  The QuatMpc
// solve is bound by neither: its roofline is the vector FMA pipe, so the
// denominator has to be measured here.  16 independent FMA chains per thread, 256 threads per
// block, 8 blocks per SM, timed with CUDA events; returns TFLOP/s (1 FMA = 2 FLOP).
#include <cuda_runtime.h>

#include "../../include/qmpc.h"

template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(T* out, int iters, T a, T b) {
  T acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (T)(threadIdx.x + i);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = acc[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T>
static double measure(int device, int iters) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  T* out = nullptr;
  if (cudaMalloc(&out, sizeof(T) * blocks * threads) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    fma_chain_kernel<T><<<blocks, threads>>>(out, iters, (T)0.999999, (T)1e-6);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

extern "C" int qmpc_measure_fma_peak(int32_t device, double* fp64_tflops, double* fp32_tflops) {
  if (!fp64_tflops || !fp32_tflops) return QMPC_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return QMPC_ERR_CUDA;
  *fp64_tflops = measure<double>(device, 4096);
  *fp32_tflops = measure<float>(device, 8192);
  return (*fp64_tflops > 0 && *fp32_tflops > 0) ? QMPC_OK : QMPC_ERR_CUDA;
}
