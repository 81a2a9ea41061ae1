// FMA-pipe peak of this GPU, measured: the denominator of bench.py's compute roofline (FP64/FP32 FMA throughput
// is not in MEASURED_PEAKS.json, so it is measured on the box).  Register-resident independent FMA chains, no memory traffic.
#include <cuda_runtime.h>

#include <algorithm>

#include "../../include/mjt.h"

namespace {

constexpr int kChains = 8;      // independent accumulators per thread: hides the 4-cycle (f32) / 8-cycle (f64) dependent latency
constexpr int kIters = 4096;    // FMAs per chain
constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads) k_fma_peak(T* sink, T a, T b) {
  T acc[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc[c] = T(threadIdx.x + c);
#pragma unroll 1
  for (int i = 0; i < kIters; i += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < kChains; ++c) acc[c] = acc[c] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += acc[c];
  if (s == T(-1)) sink[0] = s;   // never true for the launch values; keeps the chains alive
}

template <typename T>
int measure(double* tflops, cudaStream_t st) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
  T* sink = nullptr;
  if (cudaMalloc(&sink, sizeof(T)) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * 16;   // 8 resident CTAs per SM, two waves
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {   // first launch warms up
    cudaEventRecord(e0, st);
    k_fma_peak<T><<<grid, kThreads, 0, st>>>(sink, T(0.999), T(0.001));
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * kChains * (double)kIters * kThreads * (double)grid;
    if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace

extern "C" int mjt_fma_peak(int is_f64, double* tflops, void* stream) {
  if (!tflops) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return is_f64 ? measure<double>(tflops, st) : measure<float>(tflops, st);
}
