// Micro-benchmark: MUFU.EX2 throughput on sm_100a (f32, f16x2, bf16x2) and its interaction with FMA / HMMA issue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB000B000u + i + threadIdx.x; }
  float acc[4] = {0, 0, 0, 0};
  uint32_t afrag[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b0 = 0x3f803f80u, b1 = 0x3f803f80u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 3 || MODE == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) {  // + 4 FFMA per MUFU
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[0]) : "f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[1]) : "f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[2]) : "f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[3]) : "f"(a[i]));
      }
      if (MODE == 4 && (i & 1) == 0) {  // + 1 HMMA per 2 MUFU
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]) : "r"(afrag[0]), "r"(afrag[1]), "r"(afrag[2]), "r"(afrag[3]), "r"(b0), "r"(b1));
      }
      if (MODE == 5 && (i & 1) == 0) {  // HMMA only
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]) : "r"(afrag[0]), "r"(afrag[1]), "r"(afrag[2]), "r"(afrag[3]), "r"(b0), "r"(b1));
      }
    }
  }
  float s = acc[0] + acc[1] + acc[2] + acc[3];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int warps, double ops_per_iter_thread) {
  float* d; cudaMalloc(&d, 4);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, warps * 32>>>(d, 100);
  cudaEventRecord(e0);
  k<MODE><<<148, warps * 32>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double per_clk_sm = ops_per_iter_thread * iters * warps * 32.0 / (ms * 1e-3 * 1.9e9);
  printf("%-28s warps/SM %2d : %8.3f ms  -> %6.2f lane-ops/clk/SM (at 1.9 GHz), %.2f clk per warp-instr per SMSP\n", name, warps, ms, per_clk_sm, 32.0 * 4 / per_clk_sm);
  cudaFree(d);
}

int main() {
  for (int w : {4, 8, 12, 16}) run<0>("ex2.f32", w, 8);
  for (int w : {4, 8, 16}) run<1>("ex2.f16x2 (instr)", w, 8);
  for (int w : {4, 8, 16}) run<2>("ex2.bf16x2 (instr)", w, 8);
  for (int w : {4, 12, 16}) run<3>("ex2.f32 + 4 FFMA (mufu)", w, 8);
  for (int w : {4, 12, 16}) run<4>("ex2.f32 + 0.5 HMMA (mufu)", w, 8);
  for (int w : {4, 12, 16}) run<5>("HMMA only (hmma)", w, 4);
  return 0;
}
