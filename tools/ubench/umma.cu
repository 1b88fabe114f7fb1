// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16, SS mode) as a function of N, of the A-operand
// layout the implicit-GEMM kernel uses (dense tile / three dx-shifted halo copies / one wide halo box) and of concurrent
// shared-memory traffic from other warps (the epilogue's staging stores, TMA writes).  Answers: is the "60 + 0.55 N cycles"
// per MMA measured inside conv_gemm_tcgen05_kernel a hardware floor or a property of that kernel?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../fbanet_b200/csrc -o umma umma.cu && ./umma
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_ptx.cuh"

using namespace fbanet;

constexpr int HALO_COPY = 18 * 8 * 128;   // one dx-shifted copy: 18 rows x 1 KB
constexpr int WROW = 16 * 128;            // wide box: 2 KB per image row

// MODE 0: dense A tile (every tap re-reads the same 16 KB);  1: three halo copies (54 KB);  3: wide box (36 KB, 2 KB atom stride)
template <int MODE>
__device__ __forceinline__ void issue36(uint32_t tmem_d, uint32_t idesc, uint64_t desc_hi, uint64_t desc_wide, uint32_t a_lo, uint32_t b_lo,
                                        uint32_t b_step_lo, int b_slabs) {
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    uint32_t at = a_lo;
    uint64_t dA = desc_hi;
    if (MODE == 1) at += (uint32_t)(((t % 3) * HALO_COPY + (t / 3) * 1024) >> 4);
    if (MODE == 3) { at += (uint32_t)(((t / 3) * WROW + (3 + t % 3) * 128) >> 4); dA = desc_wide; }
    const uint32_t bt = b_lo + (uint32_t)(t % b_slabs) * b_step_lo;
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, dA + (uint64_t)(at + 2 * k), desc_hi + (uint64_t)(bt + 2 * k), idesc, 1u);
  }
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) umma_kernel(long long* cycles, int N, int reps, int bg_warps, int bg_store) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done[2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // layout: A region 54 KB | B region 9 slabs (as many as fit below 160 KB) | background region 32 KB
  const uint32_t b_bytes = (uint32_t)N * 128;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + 55 * 1024;
  int b_slabs = (int)((105 * 1024) / b_bytes);
  if (b_slabs > 9) b_slabs = 9;
  uint8_t* smem_bg = smem + 160 * 1024;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (threadIdx.x == 0) { mbar_init(&done[0], 1); mbar_init(&done[1], 1); fence_barrier_init(); stop = 0; }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(N);
    const uint64_t desc_hi = make_sw128_desc(0);
    const uint64_t desc_wide = (desc_hi & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)(WROW >> 4) << 32);
    const uint32_t a_lo = smem_u32(smem_a) >> 4, b_lo = smem_u32(smem_b) >> 4;
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 = warm-up
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
          issue36<MODE>(tmem_base + (uint32_t)((r & 1) * N), idesc, desc_hi, desc_wide, a_lo, b_lo, b_bytes >> 4, b_slabs);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&done[pass]);
      __syncwarp();
      mbar_wait(&done[pass], 0);
      t1 = clock64();
    }
    if (lane == 0) { cycles[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp >= 4 && warp < 4 + bg_warps) {
    // background shared-memory traffic: 16-byte loads (or stores) over a private 32 KB region until the MMA warp finishes
    uint4 acc = make_uint4(0, 0, 0, 0);
    uint8_t* base = smem_bg + ((warp - 4) & 7) * 4096;
    int n = 0;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4* q = reinterpret_cast<uint4*>(base + ((i * 512 + lane * 16) & 4095));
        if (bg_store) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(q)), "r"(acc.x), "r"(acc.y), "r"(acc.z), "r"(acc.w) : "memory");
        else { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(q))); acc.x += v.x; acc.y ^= v.y; }
      }
      ++n;
    }
    if (acc.x == 0x12345678u && n == -1) cycles[0] = acc.y;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

template <int MODE>
static void run(const char* name, int N, int grid, int bg_warps, int bg_store, int reps = 200) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  const int smem = 193 * 1024 + 1024;
  cudaFuncSetAttribute(umma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  umma_kernel<MODE><<<grid, 512, smem>>>(d, N, reps, bg_warps, bg_store);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  if (e != cudaSuccess) { printf("%s N=%d: %s\n", name, N, cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < grid; ++i) s += (double)h[i];
  // the kernel runs the sequence twice (warm-up pass + timed pass): effective clock = 2 x cycles / elapsed time
  printf("%-6s N=%3d grid=%3d bg=%2d%s reps=%6d  %7.1f cycles/MMA  (floor %d)  kernel %.3f ms -> ~%.0f MHz\n", name, N, grid, bg_warps,
         bg_warps ? (bg_store ? "st" : "ld") : "  ", reps, s / grid / (36.0 * reps), N / 2, ms, 2.0 * s / grid / ms / 1e3);
  cudaFree(d);
}

int main() {
  const int Ns[] = {16, 64, 128, 256};
  for (int grid : {1, 148})
    for (int N : Ns) {
      run<0>("dense", N, grid, 0, 0);
      run<1>("halo3", N, grid, 0, 0);
      run<3>("wide", N, grid, 0, 0);
    }
  for (int N : Ns)
    for (int bg : {4, 8, 12})
      for (int st : {0, 1}) run<1>("halo3", N, 148, bg, st);
  // sustained: ~10-40 ms of back-to-back MMAs on all SMs (power cap -> clock)
  for (int N : Ns) run<1>("halo3", N, 148, 0, 0, 20000);
  return 0;
}
