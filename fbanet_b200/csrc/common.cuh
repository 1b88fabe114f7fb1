// Shared device/host helpers for the fbanet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fbanet_b200.h"

namespace fbanet {

void set_last_error(cudaError_t e);
int check_launch();  // cudaGetLastError -> FBANET_OK / FBANET_E_LAUNCH

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// tanh-GELU as jax.nn.gelu(approximate=True): 0.5 u (1 + tanh(sqrt(2/pi) (u + 0.044715 u^3)))
__device__ __forceinline__ float gelu_tanh(float u) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (u + k1 * u * u * u);
  return 0.5f * u * (1.0f + tanhf(inner));
}
// bf16-path GELU: single MUFU.TANH (abs err ~5e-4, below bf16 output rounding)
__device__ __forceinline__ float gelu_tanh_fast(float u) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  float th;
  const float inner = fmaf(u * u, k0k1, k0) * u;   // k0 (u + k1 u^3): FMUL, FFMA, FMUL
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(inner));
  const float hu = 0.5f * u;
  return fmaf(hu, th, hu);                          // 0.5 u (1 + tanh): FMUL, FFMA
}
// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack_f2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack_f2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma_f2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul_f2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add_f2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// two bf16 packed in a 32-bit word -> fp32 pair (exact: bf16 is the high half of fp32)
__device__ __forceinline__ f32x2 bf16x2_to_f2(uint32_t u) { return pack_f2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
// tanh-GELU of a pair, bf16-path accuracy (MUFU.TANH): 5 packed ops + 2 MUFU
__device__ __forceinline__ f32x2 gelu_tanh_fast_f2(f32x2 u) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const f32x2 inner = mul_f2(fma_f2(mul_f2(u, u), pack_f2(k0k1, k0k1), pack_f2(k0, k0)), u);
  float i0, i1, t0, t1;
  unpack_f2(inner, i0, i1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(i0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(i1));
  const f32x2 hu = mul_f2(u, pack_f2(0.5f, 0.5f));
  return fma_f2(hu, pack_f2(t0, t1), hu);
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(f32x2 v) {
  float lo, hi;
  unpack_f2(v, lo, hi);
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float gelu_erf(float u) { return 0.5f * u * (1.0f + erff(u * 0.7071067811865476f)); }

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  switch (act) {
    case FBANET_ACT_RELU: return v > 0.f ? v : 0.f;
    case FBANET_ACT_PRELU: return v > 0.f ? v : alpha * v;
    case FBANET_ACT_GELU_TANH: return gelu_tanh(v);
    case FBANET_ACT_GELU_ERF: return gelu_erf(v);
    default: return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16-byte vector of T
template <typename T> struct Vec16;
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<bf16> { static constexpr int N = 8; };

template <typename T, int N> __device__ __forceinline__ void load_vec(const T* p, float (&o)[N]);
template <> __device__ __forceinline__ void load_vec<float, 4>(const float* p, float (&o)[4]) {
  float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void load_vec<bf16, 8>(const bf16* p, float (&o)[8]) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void load_vec<bf16, 4>(const bf16* p, float (&o)[4]) {
  uint2 v = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 2; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}
template <typename T, int N> __device__ __forceinline__ void store_vec(T* p, const float (&o)[N]);
template <> __device__ __forceinline__ void store_vec<float, 4>(float* p, const float (&o)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
template <> __device__ __forceinline__ void store_vec<bf16, 8>(bf16* p, const float (&o)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}
template <> __device__ __forceinline__ void store_vec<bf16, 4>(bf16* p, const float (&o)[4]) {
  uint2 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 2; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
  *reinterpret_cast<uint2*>(p) = v;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace fbanet
