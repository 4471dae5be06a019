// Micro-benchmark (run under gpurun): issue / pipe cost of the packed FP32x2 instructions of sm_100a (FFMA2)
// against scalar register-form FFMA, alone and mixed with ALU-pipe work (funnel shifts ptxas cannot fold).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/f32x2_probe tools/probes/f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned shf(unsigned a, unsigned b, unsigned c) { unsigned r; asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// MODE 0: 16 scalar FFMA (register operands)   1: 8 FFMA2 (same flops)   2: 8 SHF
//      3: 16 FFMA + 8 SHF                       4: 8 FFMA2 + 8 SHF        5: 8 FFMA2 + 16 SHF   6: 16 FFMA + 16 SHF
template <int MODE>
__global__ void __launch_bounds__(512) probe(float *out, int iters, float c, float d, unsigned sh) {
  float f[16]; u64 p[8]; unsigned q[16];
  for (int i = 0; i < 16; ++i) { f[i] = c + i + threadIdx.x; q[i] = i * 2654435761u + threadIdx.x; }
  for (int i = 0; i < 8; ++i) p[i] = (u64)__float_as_uint(f[i]) | ((u64)__float_as_uint(f[i + 8]) << 32);
  const u64 c2 = (u64)__float_as_uint(c) | ((u64)__float_as_uint(c) << 32), d2 = (u64)__float_as_uint(d) | ((u64)__float_as_uint(d) << 32);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 3 || MODE == 6) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = ffma1(f[i], c, d);
    }
    if (MODE == 1 || MODE == 4 || MODE == 5) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], c2, d2);
    }
    if (MODE == 2 || MODE == 3 || MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = shf(q[i], q[i + 8], sh);
    }
    if (MODE == 5 || MODE == 6) {
#pragma unroll
      for (int i = 0; i < 16; ++i) q[i] = shf(q[i], q[(i + 5) & 15], sh);
    }
  }
  float acc = 0;
  for (int i = 0; i < 16; ++i) acc += f[i] + (float)q[i];
  for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char *name, int fp_instr, int alu_instr) {
  float *out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  probe<MODE><<<148 * 4, 512>>>(out, 100, 0.999f, 0.001f, 3);
  cudaEventRecord(e0); probe<MODE><<<148 * 4, 512>>>(out, iters, 0.999f, 0.001f, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // 16 warps per scheduler; cycles per iteration per scheduler at 1.965 GHz
  const double cyc = ms * 1e-3 * 1.965e9 / iters;
  printf("%-22s %7.3f ms  %6.1f cycles/iter/SMSP for 16 warps x (%2d FP + %2d ALU instr) -> %.2f issue slots/cycle\n", name, ms, cyc,
         fp_instr, alu_instr, 16.0 * (fp_instr + alu_instr) / cyc);
  cudaFree(out);
}
int main() {
  run<0>("16 FFMA", 16, 0); run<1>("8 FFMA2", 8, 0); run<2>("8 SHF", 0, 8);
  run<3>("16 FFMA + 8 SHF", 16, 8); run<4>("8 FFMA2 + 8 SHF", 8, 8); run<5>("8 FFMA2 + 16 SHF", 8, 16); run<6>("16 FFMA + 16 SHF", 16, 16);
  return 0;
}
