// Probe: what does a tiled TMA load with elementStrides = {1, 2, 2, 1} put in shared memory, and how many bytes does it report to the
// mbarrier?  Tensor [1][8][16][64] bf16 channels-last with value = 100 * y + x in every channel; box {64, BX, BY, 1} at (0, -1, -1, 0).
// The wait is BOUNDED (a wrong byte count must not hang the GPU).   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_stride tma_stride.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../fbanet_b200/csrc/common.cuh"
#include "../../fbanet_b200/csrc/tc_ptx.cuh"
using namespace fbanet;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, uint32_t expect, int cx, int cy, uint16_t* out, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xFFFF;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_proxy_async();
    mbar_expect_tx(&bar, expect);
    tma_load_4d(smem, &map, &bar, 0, cx, cy, 0);
  }
  __syncwarp();
  int done = 0;
  if (threadIdx.x < 32) {
    for (int it = 0; it < (1 << 20) && !done; ++it) done = mbar_test_all(&bar, 0) ? 1 : 0;
    if (threadIdx.x == 0) *status = done;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main(int argc, char** argv) {
  const int H = 8, W = 16, C = 64;
  const int BX = argc > 1 ? atoi(argv[1]) : 8, BY = argc > 2 ? atoi(argv[2]) : 4, ES = argc > 3 ? atoi(argv[3]) : 2;
  const uint32_t expect = argc > 4 ? (uint32_t)atoi(argv[4]) : (uint32_t)(((BX + ES - 1) / ES) * ((BY + ES - 1) / ES) * 128);
  __nv_bfloat16* h = (__nv_bfloat16*)malloc(H * W * C * 2);
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c) h[(y * W + x) * C + c] = __float2bfloat16((float)(100 * y + x));
  __nv_bfloat16* d; cudaMalloc(&d, H * W * C * 2); cudaMemcpy(d, h, H * W * C * 2, cudaMemcpyHostToDevice);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeFn encode = (EncodeFn)sym;
  CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  const cuuint32_t box[4] = {64, (cuuint32_t)BX, (cuuint32_t)BY, 1};
  const cuuint32_t estr[4] = {1, (cuuint32_t)ES, (cuuint32_t)ES, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box {64,%d,%d,1} elementStrides {1,%d,%d,1} expect_tx %u: encode -> %d\n", BX, BY, ES, ES, expect, (int)r);
  if (r != CUDA_SUCCESS) return 0;
  uint16_t* dout; int* dst; cudaMalloc(&dout, 16384); cudaMalloc(&dst, 4); cudaMemset(dst, 0xff, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(map, expect, -1, -1, dout, dst);
  cudaError_t e = cudaDeviceSynchronize();
  int st = -1; cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
  uint16_t* ho = (uint16_t*)malloc(16384); cudaMemcpy(ho, dout, 16384, cudaMemcpyDeviceToHost);
  printf("sync: %s, barrier completed: %d\n", cudaGetErrorString(e), st);
  for (int row = 0; row < 40; ++row) {           // 128-byte rows; logical chunk 0 sits at physical chunk (row & 7)
    const uint16_t v = ho[row * 64 + ((0 ^ (row & 7)) * 8)];
    if (v == 0xFFFF) { printf("row %2d: untouched\n", row); if (row > 34) break; continue; }
    uint32_t u = (uint32_t)v << 16; float f; memcpy(&f, &u, 4);
    printf("row %2d: %g\n", row, f);
  }
  return 0;
}
