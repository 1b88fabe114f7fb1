// K6 (CUDA-core version): windowed multi-head self-attention with cyclic shift, relative position
// bias and the Swin shift mask computed analytically.  One CTA per (window, head); one thread per
// query row.  fp32 maths throughout (parity path); storage dtype T.
#include "common.cuh"

namespace fbanet {

// region id along one axis of the *shifted* grid (layers/fba_net.py:151-163): slices
// (0:-win), (-win:-shift), (-shift:)
__device__ __forceinline__ int shift_region(int v, int L, int win, int shift) { return v < L - win ? 0 : (v < L - shift ? 1 : 2); }

template <typename T, int DH>
__global__ void __launch_bounds__(128) window_attention_simt_kernel(const fbanet_attn_params p) {
  extern __shared__ float smem[];
  const int win = p.win, N = win * win;
  float* ks = smem;                 // [N][DH]
  float* vs = ks + N * DH;          // [N][DH]
  float* S = vs + N * DH;           // [N][N+1]
  int* tok = reinterpret_cast<int*>(S + N * (N + 1));  // [N] token offset (pixel index within the image)
  int* reg = tok + N;               // [N] shift-mask region id

  const int nwx = p.W / win, nwy = p.H / win;
  const int wid = blockIdx.x;       // global window id over the batch
  const int head = blockIdx.y;
  const int b = wid / (nwx * nwy), wl = wid % (nwx * nwy);
  const int wy = wl / nwx, wx = wl % nwx;
  const int tid = threadIdx.x;
  const T* qkv = reinterpret_cast<const T*>(p.qkv);
  const int64_t img_tok0 = (int64_t)b * p.H * p.W;

  for (int i = tid; i < N; i += blockDim.x) {
    const int ys = wy * win + i / win, xs = wx * win + i % win;  // coordinates in the shifted grid
    const int y = (ys + p.shift) % p.H, x = (xs + p.shift) % p.W;  // roll(-shift): shifted[ys] = x[(ys+shift)%H]
    tok[i] = y * p.W + x;
    reg[i] = p.shift > 0 ? shift_region(ys, p.H, win, p.shift) * 3 + shift_region(xs, p.W, win, p.shift) : 0;
  }
  __syncthreads();
  for (int e = tid; e < N * DH; e += blockDim.x) {
    const int j = e / DH, d = e % DH;
    const T* row = qkv + (img_tok0 + tok[j]) * p.qkv_ld + head * DH + d;
    ks[e] = to_f32<T>(row[p.C]);
    vs[e] = to_f32<T>(row[2 * p.C]);
  }
  __syncthreads();

  const int i = tid;
  if (i < N) {
    float q[DH];
    const T* qrow = qkv + (img_tok0 + tok[i]) * p.qkv_ld + head * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) q[d] = to_f32<T>(qrow[d]) * p.scale;
    const int yi = i / win, xi = i % win, ri = reg[i];
    float* Si = S + i * (N + 1);
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) {
      const float* kj = ks + j * DH;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s = fmaf(q[d], kj[d], s);
      const int yj = j / win, xj = j % win;
      const int idx = (yi - yj + win - 1) * (2 * win - 1) + (xi - xj + win - 1);
      s += __ldg(p.bias_table + idx * p.heads + head);
      if (ri != reg[j]) s += -100.0f;
      Si[j] = s;
      mx = fmaxf(mx, s);
    }
    float sum = 0.f;
    for (int j = 0; j < N; ++j) { const float e = expf(Si[j] - mx); Si[j] = e; sum += e; }
    const float inv = 1.0f / sum;
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
    for (int j = 0; j < N; ++j) {
      const float pj = Si[j] * inv;
      const float* vj = vs + j * DH;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(pj, vj[d], o[d]);
    }
    T* orow = reinterpret_cast<T*>(p.out) + (img_tok0 + tok[i]) * p.out_ld + head * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) orow[d] = from_f32<T>(o[d]);
  }
}

template <typename T, int DH>
static int launch_attn(const fbanet_attn_params* p, cudaStream_t s) {
  const int N = p->win * p->win;
  const size_t smem = (size_t)(2 * N * DH + N * (N + 1)) * sizeof(float) + 2 * N * sizeof(int);
  if (smem > 227 * 1024) return FBANET_E_BADSHAPE;
  auto kern = window_attention_simt_kernel<T, DH>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
  dim3 grid(p->B * (p->H / p->win) * (p->W / p->win), p->heads);
  kern<<<grid, 128, smem, s>>>(*p);
  return check_launch();
}

template <typename T>
static int dispatch_dh(const fbanet_attn_params* p, cudaStream_t s) {
  switch (p->C / p->heads) {
    case 8: return launch_attn<T, 8>(p, s);
    case 16: return launch_attn<T, 16>(p, s);
    case 32: return launch_attn<T, 32>(p, s);
    case 64: return launch_attn<T, 64>(p, s);
    default: return FBANET_E_BADSHAPE;
  }
}

int window_attention_simt_launch(const fbanet_attn_params* p, cudaStream_t s) {
  if (p->dtype == FBANET_F32) return dispatch_dh<float>(p, s);
  if (p->dtype == FBANET_BF16) return dispatch_dh<bf16>(p, s);
  return FBANET_E_DTYPE;
}

int window_attention_validate(const fbanet_attn_params* p) {
  if (!p || !p->qkv || !p->out || !p->bias_table) return FBANET_E_BADSHAPE;
  if (p->B <= 0 || p->heads <= 0 || p->win <= 0 || p->C % p->heads) return FBANET_E_BADSHAPE;
  if (p->H % p->win || p->W % p->win || p->shift < 0 || p->shift >= p->win) return FBANET_E_BADSHAPE;
  if (p->win * p->win > 128) return FBANET_E_BADSHAPE;
  if (p->qkv_ld < 3 * p->C || p->out_ld < p->C) return FBANET_E_BADSHAPE;
  return FBANET_OK;
}

}  // namespace fbanet
