// Narrow host I/O for the end-to-end path: the reference's data path is 8-bit on both ends -- uint8 frames normalised to [0,1] on
// the way in (train.py:82-83: `data["lr_frames"].astype(float32) / 255.0`) and the clamped SR image written as 8-bit PNG on the way out
// (test_in_any_resolution.py:93-101: `torch.clamp(x, 0, 1)` then torchvision `ToPILImage` = `mul(255).byte()`, i.e. truncation).
// These two conversions run on the device, so the pinned host buffers -- and the PCIe / host-memory traffic of 8 ranks -- shrink 4x.
//   mode U8_TO_F32:  dst_f32 = (float)src_u8 / 255.0f           (IEEE division, as numpy / jax do)
//   mode F32_TO_U8:  dst_u8  = (uint8)trunc(clamp(src_f32, 0, 1) * 255.0f)
//   mode F32_TO_F16: dst_f16 = half_rn(src_f32)                 (no clamp: callers clamp, models/fba_net.py:320 does not)
#include <cuda_fp16.h>

#include "common.cuh"

namespace fbanet {

template <int MODE>
__global__ void __launch_bounds__(256) convert_io_kernel(const void* __restrict__ src, void* __restrict__ dst, int64_t n16) {
  // 16 elements per thread: one 16-byte u8 vector / four float4
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    if (MODE == FBANET_CONVERT_U8_TO_F32) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      float4* o = reinterpret_cast<float4*>(dst) + 4 * i;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = make_float4(__fdiv_rn((float)(w[k] & 0xff), 255.f), __fdiv_rn((float)((w[k] >> 8) & 0xff), 255.f),
                           __fdiv_rn((float)((w[k] >> 16) & 0xff), 255.f), __fdiv_rn((float)(w[k] >> 24), 255.f));
    } else {
      const float4* s = reinterpret_cast<const float4*>(src) + 4 * i;
      float f[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float4 v = __ldg(s + k); f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w; }
      if (MODE == FBANET_CONVERT_F32_TO_U8) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t b[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) b[e] = (uint32_t)(fminf(fmaxf(f[4 * k + e], 0.f), 1.f) * 255.f);   // float -> unsigned: truncation
          w[k] = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        }
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const __half2 h = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
          w[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        uint4* o = reinterpret_cast<uint4*>(dst) + 2 * i;
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
}

}  // namespace fbanet

using namespace fbanet;

extern "C" int fbanet_convert_io_sm100(const fbanet_convert_io_params* p, void* stream) {
  if (!p || !p->src || !p->dst || p->n <= 0 || (p->n % 16)) return FBANET_E_BADSHAPE;
  if (((uintptr_t)p->src % 16) || ((uintptr_t)p->dst % 16)) return FBANET_E_ALIGN;
  const int64_t n16 = p->n / 16;
  int64_t blocks = (n16 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  switch (p->mode) {
    case FBANET_CONVERT_U8_TO_F32: convert_io_kernel<FBANET_CONVERT_U8_TO_F32><<<(unsigned)blocks, 256, 0, st>>>(p->src, p->dst, n16); break;
    case FBANET_CONVERT_F32_TO_U8: convert_io_kernel<FBANET_CONVERT_F32_TO_U8><<<(unsigned)blocks, 256, 0, st>>>(p->src, p->dst, n16); break;
    case FBANET_CONVERT_F32_TO_F16: convert_io_kernel<FBANET_CONVERT_F32_TO_F16><<<(unsigned)blocks, 256, 0, st>>>(p->src, p->dst, n16); break;
    default: return FBANET_E_DTYPE;
  }
  return check_launch();
}
