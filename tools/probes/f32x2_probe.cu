// Micro-benchmark (run under gpurun): issue cost of the packed FP32x2 instructions of sm_100a
// (FFMA2/FADD2/FMUL2) against scalar FFMA/FADD, alone and mixed with ALU-pipe work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/f32x2_probe tools/probes/f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512) probe(float *out, int iters, float seed) {
  float f[8]; u64 p[8]; unsigned q[8];
  for (int i = 0; i < 8; ++i) { f[i] = seed + i + threadIdx.x; p[i] = (u64)__float_as_uint(f[i]) | ((u64)__float_as_uint(f[i] + 1) << 32); q[i] = i + threadIdx.x; }
  const float c = 0.999f, d = 0.001f; const u64 c2 = (u64)__float_as_uint(c) | ((u64)__float_as_uint(c) << 32), d2 = (u64)__float_as_uint(d) | ((u64)__float_as_uint(d) << 32);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) f[i] = ffma1(f[i], c, d);                       // 8 FFMA
      if (MODE == 1) p[i] = ffma2(p[i], c2, d2);                     // 8 FFMA2 (16 fma)
      if (MODE == 2) f[i] = fadd1(f[i], d);                          // 8 FADD
      if (MODE == 3) p[i] = fadd2(p[i], d2);                         // 8 FADD2
      if (MODE == 4) { f[i] = ffma1(f[i], c, d); q[i] = lop(q[i], 0x5a5a5a5au + i); }   // FFMA + LOP
      if (MODE == 5) { p[i] = ffma2(p[i], c2, d2); q[i] = lop(q[i], 0x5a5a5a5au + i); } // FFMA2 + LOP
      if (MODE == 6) { p[i] = ffma2(p[i], c2, d2); q[i] = lop(q[i], 0x5a5a5a5au + i); q[i] = lop(q[i], 0x1234567u + i); } // FFMA2 + 2 LOP
      if (MODE == 7) { f[i] = ffma1(f[i], c, d); f[i] = fadd1(f[i], d); q[i] = lop(q[i], 0x5a5a5a5au + i); q[i] = lop(q[i], 0x1234567u + i); } // 2 fp + 2 alu
    }
  }
  float acc = 0; for (int i = 0; i < 8; ++i) acc += f[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + (float)q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char *name, int instr_per_iter) {
  float *out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  probe<MODE><<<148 * 4, 512>>>(out, 100, 1.f);
  cudaEventRecord(e0); probe<MODE><<<148 * 4, 512>>>(out, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double warp_instr = (double)148 * 4 * 16 * iters * instr_per_iter;  // warps * instrs
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-28s %8.3f ms  %.3f warp-instr/clk/SMSP (of the counted %d per iter)\n", name, ms, warp_instr / cycles / (148 * 4), instr_per_iter);
  cudaFree(out);
}
int main() {
  run<0>("FFMA", 8); run<1>("FFMA2", 8); run<2>("FADD", 8); run<3>("FADD2", 8);
  run<4>("FFMA+LOP", 16); run<5>("FFMA2+LOP", 16); run<6>("FFMA2+2LOP", 24); run<7>("FFMA+FADD+2LOP", 32);
  return 0;
}
