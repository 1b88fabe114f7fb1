// SURVEY 8f-3, backward bricks of the training step (what torch autograd did for train.py.bak:163-169 and what
// eqx.filter_value_and_grad does at train.py:63): weight / bias gradients of every convolution and linear layer, LayerNorm backward,
// activation backward.  The DATA gradient of the stride-1 convolutions and of the linear layers needs no kernel of its own: it is
// the forward implicit GEMM (fbanet_conv_gemm_sm100, tcgen05 in bf16) run on dY with flipped / transposed weights
// (fbanet_b200/train.py: dgrad_weight).
//
// First correct version: CUDA-core fp32 accumulation everywhere, fixed-order (bit-reproducible) two-stage reductions.  The weight
// gradient is a GEMM with K = pixels whose two operands both arrive pixel-major (MN-major for the tensor core): the tcgen05 form
// (transposing shared-memory descriptors, split-K over CTAs) is DESIGN.md 8c's next step; this kernel is its parity reference.
#include "common.cuh"

namespace fbanet {

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// weight gradient: CTA tile 64 output channels x 64 input channels of ONE tap, over one chunk of the pixel axis
//   grid = (ceil(Cout/64), KH*KW*ceil(Cin/64), splits), 256 threads, 4 x 4 accumulators per thread
// ------------------------------------------------------------------------------------------------
constexpr int WG_T = 64;      // tile edge (channels)
constexpr int WG_P = 16;      // pixels per shared-memory step
constexpr int WG_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(WG_THREADS) wgrad_kernel(const fbanet_wgrad_params p, int64_t chunk) {
  __shared__ __align__(16) float sdy[WG_P][WG_T + 4];
  __shared__ __align__(16) float sx[WG_P][WG_T + 4];
  const int ci_tiles = (p.Cin + WG_T - 1) / WG_T;
  const int tap = blockIdx.y / ci_tiles, ci0 = (blockIdx.y % ci_tiles) * WG_T, co0 = blockIdx.x * WG_T;
  const int ky = tap / p.KW, kx = tap % p.KW;
  const int64_t P = (int64_t)p.N * p.Ho * p.Wo;
  const int64_t p_begin = (int64_t)blockIdx.z * chunk, p_end = min(P, p_begin + chunk);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const T* __restrict__ X = static_cast<const T*>(p.x);
  const T* __restrict__ DY = static_cast<const T*>(p.dy);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool want_bias = p.db != nullptr && blockIdx.y == 0;   // the tap-0 / ci-tile-0 CTAs also own the bias gradient

  for (int64_t pb = p_begin; pb < p_end; pb += WG_P) {
    // 16 pixels x 64 channels of each operand: thread t loads elements t, t + 256, ... (channels fastest: coalesced)
#pragma unroll
    for (int k = 0; k < WG_P * WG_T / WG_THREADS; ++k) {
      const int e = tid + k * WG_THREADS, row = e / WG_T, col = e % WG_T;
      const int64_t pix = pb + row;
      float vdy = 0.f, vx = 0.f;
      if (pix < p_end) {
        const int n = (int)(pix / ((int64_t)p.Ho * p.Wo));
        const int rem = (int)(pix - (int64_t)n * p.Ho * p.Wo);
        const int yo = rem / p.Wo, xo = rem - yo * p.Wo;
        if (co0 + col < p.Cout) vdy = to_f32<T>(DY[(int64_t)n * p.dy_img_stride + (int64_t)rem * p.dy_ld + co0 + col]);
        const int yi = yo * p.stride + ky - p.pad, xi = xo * p.stride + kx - p.pad;
        if (ci0 + col < p.Cin && yi >= 0 && yi < p.H && xi >= 0 && xi < p.W)
          vx = to_f32<T>(X[(int64_t)n * p.x_img_stride + ((int64_t)yi * p.W + xi) * p.x_ld + ci0 + col]);
      }
      sdy[row][col] = vdy;
      sx[row][col] = vx;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WG_P; ++r) {
      const float4 av = *reinterpret_cast<const float4*>(&sdy[r][ty * 4]);   // row pitch 68 floats: 16-byte aligned
      const float4 bv = *reinterpret_cast<const float4*>(&sx[r][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (want_bias && tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bsum[i] += a[i];
      }
    }
    __syncthreads();
  }
  // partial[split][co][tap][ci]  followed by  partial_bias[split][co]
  const int64_t K = (int64_t)p.KH * p.KW * p.Cin;
  float* part = p.partial + (int64_t)blockIdx.z * p.Cout * K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < p.Cin) part[(int64_t)co * K + (int64_t)tap * p.Cin + ci] = acc[i][j];
    }
    if (want_bias && tx == 0) p.partial[(int64_t)p.splits * p.Cout * K + (int64_t)blockIdx.z * p.Cout + co] = bsum[i];
  }
}

// sum the splits in order; dw in the torch layout [Cout][Cin][KH][KW]
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const fbanet_wgrad_params p) {
  const int64_t K = (int64_t)p.KH * p.KW * p.Cin, total = (int64_t)p.Cout * K;
  const int taps = p.KH * p.KW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total + p.Cout; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < total) {   // i = (co * Cin + ci) * taps + tap
      const int tap = (int)(i % taps);
      const int64_t r = i / taps;
      const int ci = (int)(r % p.Cin);
      const int64_t co = r / p.Cin;
      const int64_t src = co * K + (int64_t)tap * p.Cin + ci;
      float s = 0.f;
      for (int z = 0; z < p.splits; ++z) s += p.partial[(int64_t)z * total + src];
      p.dw[i] = p.accumulate ? p.dw[i] + s : s;
    } else if (p.db != nullptr) {
      const int64_t co = i - total;
      float s = 0.f;
      for (int z = 0; z < p.splits; ++z) s += p.partial[(int64_t)p.splits * total + (int64_t)z * p.Cout + co];
      p.db[co] = p.accumulate ? p.db[co] + s : s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward: one warp per row, lane l owns channels l, l + 32, ... (C <= 256)
// ------------------------------------------------------------------------------------------------
constexpr int LNB_WARPS = 8;
constexpr int LNB_MAXK = 8;

template <typename T>
__global__ void __launch_bounds__(LNB_WARPS * 32) layernorm_bwd_kernel(const fbanet_layernorm_bwd_params p) {
  __shared__ float sred[LNB_WARPS][2][LNB_MAXK * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, C = p.C;
  const T* __restrict__ X = static_cast<const T*>(p.x);
  const T* __restrict__ DY = static_cast<const T*>(p.dy);
  T* __restrict__ DX = static_cast<T*>(p.dx);
  float gam[LNB_MAXK], dg[LNB_MAXK], db[LNB_MAXK];
#pragma unroll
  for (int k = 0; k < LNB_MAXK; ++k) {
    const int c = lane + 32 * k;
    gam[k] = c < C ? p.gamma[c] : 0.f;
    dg[k] = 0.f;
    db[k] = 0.f;
  }
  const float inv_c = 1.f / (float)C;
  for (int64_t row = (int64_t)blockIdx.x * LNB_WARPS + warp; row < p.rows; row += (int64_t)gridDim.x * LNB_WARPS) {
    float x[LNB_MAXK], dy[LNB_MAXK];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LNB_MAXK; ++k) {
      const int c = lane + 32 * k;
      x[k] = c < C ? to_f32<T>(X[row * C + c]) : 0.f;
      dy[k] = c < C ? to_f32<T>(DY[row * C + c]) : 0.f;
      s += x[k];
    }
    const float mean = warp_sum(s) * inv_c;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < LNB_MAXK; ++k) {
      const float d = (lane + 32 * k < C) ? x[k] - mean : 0.f;
      v = fmaf(d, d, v);
    }
    const float rstd = rsqrtf(warp_sum(v) * inv_c + p.eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LNB_MAXK; ++k) {
      const bool in = lane + 32 * k < C;
      x[k] = in ? (x[k] - mean) * rstd : 0.f;       // xhat
      const float g = dy[k] * gam[k];
      s1 += g;
      s2 = fmaf(g, x[k], s2);
      dg[k] = fmaf(dy[k], x[k], dg[k]);
      db[k] += dy[k];
    }
    const float m1 = warp_sum(s1) * inv_c, m2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int k = 0; k < LNB_MAXK; ++k) {
      const int c = lane + 32 * k;
      if (c < C) DX[row * C + c] = from_f32<T>(rstd * (dy[k] * gam[k] - m1 - x[k] * m2));
    }
  }
#pragma unroll
  for (int k = 0; k < LNB_MAXK; ++k) {
    sred[warp][0][lane + 32 * k] = dg[k];
    sred[warp][1][lane + 32 * k] = db[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_WARPS; ++w) { a += sred[w][0][c]; b += sred[w][1][c]; }
    p.partial[((int64_t)blockIdx.x * 2 + 0) * C + c] = a;
    p.partial[((int64_t)blockIdx.x * 2 + 1) * C + c] = b;
  }
}

__global__ void __launch_bounds__(256) layernorm_bwd_finish_kernel(const fbanet_layernorm_bwd_params p, int blocks) {
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int z = 0; z < blocks; ++z) {
      a += p.partial[((int64_t)z * 2 + 0) * p.C + c];
      b += p.partial[((int64_t)z * 2 + 1) * p.C + c];
    }
    if (p.dgamma) p.dgamma[c] = p.accumulate ? p.dgamma[c] + a : a;
    if (p.dbeta) p.dbeta[c] = p.accumulate ? p.dbeta[c] + b : b;
  }
}

// ------------------------------------------------------------------------------------------------
// activation backward
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(float x, int act, float alpha) {
  switch (act) {
    case FBANET_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case FBANET_ACT_PRELU: return x > 0.f ? 1.f : alpha;     // forward: v > 0 ? v : alpha v (apply_act)
    case FBANET_ACT_GELU_TANH: {
      const float k0 = 0.7978845608028654f, k1 = 0.044715f;
      const float t = tanhf(k0 * (x + k1 * x * x * x));
      return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k0 * (1.f + 3.f * k1 * x * x);
    }
    case FBANET_ACT_GELU_ERF:
      return 0.5f * (1.f + erff(x * 0.7071067811865476f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
    default: return 1.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const fbanet_act_bwd_params p) {
  __shared__ float sred[256];
  const T* __restrict__ X = static_cast<const T*>(p.x);
  const T* __restrict__ DY = static_cast<const T*>(p.dy);
  T* __restrict__ DX = static_cast<T*>(p.dx);
  const float alpha = (p.act == FBANET_ACT_PRELU && p.alpha) ? *p.alpha : 0.f;
  float da = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = to_f32<T>(X[i]), dy = to_f32<T>(DY[i]);
    DX[i] = from_f32<T>(dy * act_grad(x, p.act, alpha));
    if (!(x > 0.f)) da = fmaf(dy, x, da);
  }
  if (p.act == FBANET_ACT_PRELU && p.dalpha) {
    sred[threadIdx.x] = da;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) sred[threadIdx.x] += sred[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) p.partial[blockIdx.x] = sred[0];
  }
}

// training-mode forward of a stand-alone activation: y = act(x) with the PRE-activation kept by the caller for act_bwd_kernel
// (the inference path applies activations inside the producing GEMM's epilogue and never stores x)
template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const fbanet_act_fwd_params p) {
  const T* __restrict__ X = static_cast<const T*>(p.x);
  T* __restrict__ Y = static_cast<T*>(p.y);
  const float alpha = (p.act == FBANET_ACT_PRELU && p.alpha) ? *p.alpha : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x)
    Y[i] = from_f32<T>(apply_act(to_f32<T>(X[i]), p.act, alpha));
}

__global__ void act_bwd_finish_kernel(const fbanet_act_bwd_params p, int blocks) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int z = 0; z < blocks; ++z) s += p.partial[z];
    *p.dalpha = p.accumulate ? *p.dalpha + s : s;
  }
}

}  // namespace fbanet

using namespace fbanet;

extern "C" int fbanet_wgrad_sm100(const fbanet_wgrad_params* p, void* stream) {
  if (!p || !p->x || !p->dy || !p->dw || !p->partial) return FBANET_E_BADSHAPE;
  if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->Cin <= 0 || p->Cout <= 0 || p->KH <= 0 || p->KW <= 0 || p->stride <= 0 || p->pad < 0 ||
      p->splits <= 0 || p->splits > 65535 || p->x_ld < p->Cin || p->dy_ld < p->Cout)
    return FBANET_E_BADSHAPE;
  if (p->Ho != (p->H + 2 * p->pad - p->KH) / p->stride + 1 || p->Wo != (p->W + 2 * p->pad - p->KW) / p->stride + 1 || p->Ho <= 0 || p->Wo <= 0)
    return FBANET_E_BADSHAPE;
  if (p->dtype != FBANET_F32 && p->dtype != FBANET_BF16) return FBANET_E_DTYPE;
  const int64_t P = (int64_t)p->N * p->Ho * p->Wo;
  const int64_t chunk = cdiv64(cdiv64(P, p->splits), WG_P) * WG_P;
  const int64_t ktiles = (int64_t)p->KH * p->KW * ((p->Cin + WG_T - 1) / WG_T);
  if (ktiles > 65535) return FBANET_E_BADSHAPE;
  const dim3 grid((unsigned)((p->Cout + WG_T - 1) / WG_T), (unsigned)ktiles, (unsigned)p->splits);
  if (p->dtype == FBANET_F32) wgrad_kernel<float><<<grid, WG_THREADS, 0, (cudaStream_t)stream>>>(*p, chunk);
  else wgrad_kernel<bf16><<<grid, WG_THREADS, 0, (cudaStream_t)stream>>>(*p, chunk);
  const int64_t total = (int64_t)p->Cout * p->KH * p->KW * p->Cin + p->Cout;
  int64_t blocks = cdiv64(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  wgrad_finish_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}

extern "C" int fbanet_layernorm_bwd_blocks(int64_t rows) {
  if (rows <= 0) return -1;
  const int64_t b = cdiv64(rows, LNB_WARPS);
  return (int)(b < 148 * 4 ? b : 148 * 4);
}

extern "C" int fbanet_layernorm_bwd_sm100(const fbanet_layernorm_bwd_params* p, void* stream) {
  if (!p || !p->x || !p->dy || !p->gamma || !p->dx || !p->partial || p->rows <= 0 || p->C <= 0 || p->C > LNB_MAXK * 32 || !(p->eps > 0.f))
    return FBANET_E_BADSHAPE;
  if (p->dtype != FBANET_F32 && p->dtype != FBANET_BF16) return FBANET_E_DTYPE;
  const int blocks = fbanet_layernorm_bwd_blocks(p->rows);
  if (p->dtype == FBANET_F32) layernorm_bwd_kernel<float><<<blocks, LNB_WARPS * 32, 0, (cudaStream_t)stream>>>(*p);
  else layernorm_bwd_kernel<bf16><<<blocks, LNB_WARPS * 32, 0, (cudaStream_t)stream>>>(*p);
  if (p->dgamma || p->dbeta) layernorm_bwd_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*p, blocks);
  return check_launch();
}

extern "C" int fbanet_act_bwd_blocks(int64_t n) {
  if (n <= 0) return -1;
  const int64_t b = cdiv64(n, 256 * 4);
  return (int)(b < 148 * 8 ? b : 148 * 8);
}

extern "C" int fbanet_act_bwd_sm100(const fbanet_act_bwd_params* p, void* stream) {
  if (!p || !p->x || !p->dy || !p->dx || p->n <= 0) return FBANET_E_BADSHAPE;
  if (p->act < FBANET_ACT_RELU || p->act > FBANET_ACT_GELU_ERF) return FBANET_E_BADSHAPE;
  if (p->act == FBANET_ACT_PRELU && (!p->alpha || (p->dalpha && !p->partial))) return FBANET_E_BADSHAPE;
  if (p->dtype != FBANET_F32 && p->dtype != FBANET_BF16) return FBANET_E_DTYPE;
  const int blocks = fbanet_act_bwd_blocks(p->n);
  if (p->dtype == FBANET_F32) act_bwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  else act_bwd_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  if (p->act == FBANET_ACT_PRELU && p->dalpha) act_bwd_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*p, blocks);
  return check_launch();
}

extern "C" int fbanet_act_fwd_sm100(const fbanet_act_fwd_params* p, void* stream) {
  if (!p || !p->x || !p->y || p->n <= 0) return FBANET_E_BADSHAPE;
  if (p->act < FBANET_ACT_NONE || p->act > FBANET_ACT_GELU_ERF) return FBANET_E_BADSHAPE;
  if (p->act == FBANET_ACT_PRELU && !p->alpha) return FBANET_E_BADSHAPE;
  if (p->dtype != FBANET_F32 && p->dtype != FBANET_BF16) return FBANET_E_DTYPE;
  const int blocks = fbanet_act_bwd_blocks(p->n);
  if (p->dtype == FBANET_F32) act_fwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  else act_fwd_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}
