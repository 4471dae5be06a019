// ld_probe.cu — micro-benchmarks that measure, on the box, the denominators of this path's roofline:
// the non-fused FP64 (DADD/DMUL) issue rate, the non-fused FP32 rate and the random 8-byte gather rate
// from an L2-resident window the size of the DFIRE table.  Used by bench.py only.
#include <cuda_runtime.h>

#include <string>

#include "../../include/lightdock_b200.h"

namespace {

template <typename T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <>
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
template <>
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
template <typename T>
__device__ __forceinline__ T add_rn(T a, T b);
template <>
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
template <>
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// 8 independent chains per thread of (a*c)+d, never fused: 16 flops per inner iteration per thread.
template <typename T>
__global__ void __launch_bounds__(256) flop_probe(T *out, int iters, T c, T d) {
  T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
    a0 = add_rn(mul_rn(a0, c), d); a1 = add_rn(mul_rn(a1, c), d); a2 = add_rn(mul_rn(a2, c), d);
    a3 = add_rn(mul_rn(a3, c), d); a4 = add_rn(mul_rn(a4, c), d); a5 = add_rn(mul_rn(a5, c), d);
    a6 = add_rn(mul_rn(a6, c), d); a7 = add_rn(mul_rn(a7, c), d);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// Random 8-byte loads (index stream independent of the loaded data, 4 loads in flight per iteration).
__global__ void __launch_bounds__(256) gather_probe(const double *__restrict__ table, unsigned n, double *out,
                                                    int iters) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  double acc = 0.0;
  for (int i = 0; i < iters; ++i) {
    unsigned i0 = s % n; s = s * 1664525u + 1013904223u;
    unsigned i1 = s % n; s = s * 1664525u + 1013904223u;
    unsigned i2 = s % n; s = s * 1664525u + 1013904223u;
    unsigned i3 = s % n; s = s * 1664525u + 1013904223u;
    acc += __ldg(table + i0) + __ldg(table + i1) + __ldg(table + i2) + __ldg(table + i3);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
cudaError_t best_ms(F &&launch, int reps, float *best) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  *best = 1e30f;
  launch();  // warm-up
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < *best) *best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return cudaGetLastError();
}

}  // namespace

extern "C" int ld_probe_peaks(int32_t device, double *fp64_nonfma_tflops, double *fp32_nonfma_tflops,
                              double *l2_gather_gloads) {
  if (cudaSetDevice(device) != cudaSuccess) return LD_ECUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LD_ECUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  void *out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return LD_ENOMEM;
  float ms = 0;
  cudaError_t e = cudaSuccess;
  if (fp64_nonfma_tflops) {
    const int iters = 4096;
    e = best_ms([&] { flop_probe<double><<<blocks, threads>>>((double *)out, iters, 1.0000001, 1e-9); }, 5, &ms);
    *fp64_nonfma_tflops = (double)blocks * threads * iters * 16.0 / (ms * 1e-3) / 1e12;
  }
  if (e == cudaSuccess && fp32_nonfma_tflops) {
    const int iters = 8192;
    e = best_ms([&] { flop_probe<float><<<blocks, threads>>>((float *)out, iters, 1.0000001f, 1e-9f); }, 5, &ms);
    *fp32_nonfma_tflops = (double)blocks * threads * iters * 16.0 / (ms * 1e-3) / 1e12;
  }
  if (e == cudaSuccess && l2_gather_gloads) {
    double *table = nullptr;
    if (cudaMalloc(&table, (size_t)LD_DFIRE_TABLE_LEN * sizeof(double)) != cudaSuccess) {
      cudaFree(out);
      return LD_ENOMEM;
    }
    cudaMemset(table, 0, (size_t)LD_DFIRE_TABLE_LEN * sizeof(double));
    const int iters = 512;
    e = best_ms([&] { gather_probe<<<blocks, threads>>>(table, LD_DFIRE_TABLE_LEN, (double *)out, iters); }, 5, &ms);
    *l2_gather_gloads = (double)blocks * threads * iters * 4.0 / (ms * 1e-3) / 1e9;
    cudaFree(table);
  }
  cudaFree(out);
  return e == cudaSuccess ? LD_OK : LD_ECUDA;
}
