// SURVEY 8f-3, second set of backward bricks of the training step: the layers whose gradient is NOT a dense GEMM --
//   * LeFF depthwise 3x3 (layers/locally_enhanced_feed_forward.py:39-52): data, weight and bias gradient,
//   * windowed attention (layers/window_attention.py:159-248 + the shift / partition / mask of layers/fba_net.py:149-238):
//     dq, dk, dv and the gradient of the relative-position table,
//   * the Federated-Affinity gate (blocks/federated_affinity_fusion.py:79-105): gradient through sigmoid(|a_f - a_0|) into the
//     features and into temporal_attn1's weights,
//   * DropPath residual  out = skip + scale[burst] * x  (layers/drop_path.py:39-63, layers/fba_net.py:245-248), forward AND backward
//     (the backward of the branch is the same row scaling applied to dy).
// First correct versions: CUDA-core fp32 arithmetic, fixed-order (bit-reproducible) two-stage reductions, no atomics.  All four are
// HBM / L2-bound index work except the attention core, whose tensor-core form is DESIGN.md 8c's next step.
#include "common.cuh"

namespace fbanet {

static inline int64_t cdiv64s(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int SP_CT = 64;        // channels per CTA (fastest thread index: coalesced channels-last rows)
constexpr int SP_ROWS = 4;       // pixel lanes per CTA
constexpr int SP_THREADS = SP_CT * SP_ROWS;
constexpr int SP_MAX_BLOCKS_X = 148 * 4;

// fixed-order reduction of NV per-thread sums over the SP_ROWS pixel lanes of a CTA -> partial[blockIdx.x][v][C]
template <int NV>
__device__ __forceinline__ void cta_reduce_store(float (&acc)[NV], float* sred, float* partial, int C, int c) {
  const int lane_c = threadIdx.x % SP_CT, r = threadIdx.x / SP_CT;
#pragma unroll
  for (int v = 0; v < NV; ++v) sred[(r * NV + v) * SP_CT + lane_c] = acc[v];
  __syncthreads();
  if (r == 0 && c < C) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float s = 0.f;
#pragma unroll
      for (int rr = 0; rr < SP_ROWS; ++rr) s += sred[(rr * NV + v) * SP_CT + lane_c];
      partial[((int64_t)blockIdx.x * NV + v) * C + c] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 backward.  y(q,c) = sum_tap w[tap][c] x(q + tap - 1, c) + b[c]
//   dx(q,c) = sum_tap w[tap][c] dy(q - (tap - 1), c);  dw[c][tap] = sum_q dy(q,c) x(q + tap - 1, c);  db[c] = sum_q dy(q,c)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(SP_THREADS) dwconv_bwd_kernel(const fbanet_dwconv_bwd_params p) {
  __shared__ float sred[SP_ROWS * 10 * SP_CT];
  const int c = blockIdx.y * SP_CT + threadIdx.x % SP_CT, r = threadIdx.x / SP_CT;
  const bool live = c < p.C;
  const T* __restrict__ X = static_cast<const T*>(p.x);
  const T* __restrict__ DY = static_cast<const T*>(p.dy);
  T* __restrict__ DX = static_cast<T*>(p.dx);
  float w[9], acc[10];
#pragma unroll
  for (int t = 0; t < 9; ++t) w[t] = live ? p.weight[t * p.C + c] : 0.f;
#pragma unroll
  for (int t = 0; t < 10; ++t) acc[t] = 0.f;
  const int64_t P = (int64_t)p.N * p.H * p.W;
  if (live) {
    for (int64_t pix = (int64_t)blockIdx.x * SP_ROWS + r; pix < P; pix += (int64_t)gridDim.x * SP_ROWS) {
      const int64_t n = pix / ((int64_t)p.H * p.W);
      const int rem = (int)(pix - n * p.H * p.W), y = rem / p.W, x = rem - y * p.W;
      const int64_t img = n * p.H * p.W;
      const float dyc = to_f32<T>(DY[pix * p.C + c]);
      float dx = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int t = ky * 3 + kx;
          const int ys = y + ky - 1, xs = x + kx - 1;          // forward tap of THIS output pixel: weight gradient
          if (ys >= 0 && ys < p.H && xs >= 0 && xs < p.W) acc[t] = fmaf(dyc, to_f32<T>(X[(img + (int64_t)ys * p.W + xs) * p.C + c]), acc[t]);
          const int yo = y - (ky - 1), xo = x - (kx - 1);      // output pixel that read THIS input pixel through tap t
          if (yo >= 0 && yo < p.H && xo >= 0 && xo < p.W) dx = fmaf(w[t], to_f32<T>(DY[(img + (int64_t)yo * p.W + xo) * p.C + c]), dx);
        }
      acc[9] += dyc;
      if (DX) DX[pix * p.C + c] = from_f32<T>(dx);
    }
  }
  cta_reduce_store<10>(acc, sred, p.partial, p.C, c);
}

// dw[c][tap] / db[c] in the torch layouts ([C,1,3,3], [C]); sums the CTAs' partials in order
__global__ void __launch_bounds__(256) dwconv_bwd_finish_kernel(const fbanet_dwconv_bwd_params p, int blocks) {
  const int total = 10 * p.C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i / p.C, c = i % p.C;
    float s = 0.f;
    for (int z = 0; z < blocks; ++z) s += p.partial[((int64_t)z * 10 + v) * p.C + c];
    if (v < 9) {
      if (p.dw) p.dw[c * 9 + v] = p.accumulate ? p.dw[c * 9 + v] + s : s;
    } else if (p.db) {
      p.db[c] = p.accumulate ? p.db[c] + s : s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Federated-Affinity gate backward.  Forward (fbanet_faf_gate_sm100): s_f(p) = sum_{tap,c} wsum[tap][c] feat_f(p + tap - 1, c),
//   g_f = sigmoid(|s_f - s_0|), gated_f = feat_f * g_f (f >= 1), gated_0 = feat_0.
// pass 1 (one warp per pixel): ds_f(p) = sign(s_f - s_0) g_f (1 - g_f) sum_c dgated_f(p,c) feat_f(p,c);  ds_0 = -sum_f ds_f
// pass 2 (thread = channel, pixel lanes): dfeat_f(q,c) = dgated_f(q,c) g_f(q) + sum_tap wsum[tap][c] ds_f(q - (tap - 1));
//                                         dwsum[tap][c] = sum_{b,f,q} feat_f(q,c) ds_f(q - (tap - 1))
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) faf_gate_bwd_dscore_kernel(const fbanet_faf_gate_bwd_params p) {
  const int lane = threadIdx.x & 31;
  const int64_t HW = (int64_t)p.H * p.W, total = (int64_t)p.B * HW;
  const T* __restrict__ FEAT = static_cast<const T*>(p.feat);
  const T* __restrict__ DG = static_cast<const T*>(p.dgated);
  for (int64_t bp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); bp < total; bp += (int64_t)gridDim.x * 8) {
    const int64_t b = bp / HW, pix = bp - b * HW;
    const float s0 = p.score[(b * p.F + 0) * HW + pix];
    float ds0 = 0.f;
    for (int f = 1; f < p.F; ++f) {
      const T* fr = FEAT + ((b * p.F + f) * HW + pix) * p.C;
      const T* dr = DG + ((b * HW + pix) * p.F + f) * p.C;
      float dot = 0.f;
      for (int c = lane; c < p.C; c += 32) dot = fmaf(to_f32<T>(dr[c]), to_f32<T>(fr[c]), dot);
      dot = warp_sum(dot);
      const float d = p.score[(b * p.F + f) * HW + pix] - s0;
      const float g = p.gate[(b * (p.F - 1) + f - 1) * HW + pix];
      const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      const float ds = sgn * g * (1.f - g) * dot;
      ds0 -= ds;
      if (lane == 0) p.dscore[(b * p.F + f) * HW + pix] = ds;
    }
    if (lane == 0) p.dscore[(b * p.F + 0) * HW + pix] = ds0;
  }
}

template <typename T>
__global__ void __launch_bounds__(SP_THREADS) faf_gate_bwd_dfeat_kernel(const fbanet_faf_gate_bwd_params p) {
  __shared__ float sred[SP_ROWS * 9 * SP_CT];
  const int c = blockIdx.y * SP_CT + threadIdx.x % SP_CT, r = threadIdx.x / SP_CT;
  const bool live = c < p.C;
  const T* __restrict__ FEAT = static_cast<const T*>(p.feat);
  const T* __restrict__ DG = static_cast<const T*>(p.dgated);
  T* __restrict__ DF = static_cast<T*>(p.dfeat);
  float w[9], acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) { w[t] = live ? p.wsum[t * p.C + c] : 0.f; acc[t] = 0.f; }
  const int64_t HW = (int64_t)p.H * p.W, P = (int64_t)p.B * p.F * HW;
  if (live) {
    for (int64_t e = (int64_t)blockIdx.x * SP_ROWS + r; e < P; e += (int64_t)gridDim.x * SP_ROWS) {
      const int64_t bf = e / HW;                  // b * F + f
      const int pix = (int)(e - bf * HW), y = pix / p.W, x = pix - y * p.W;
      const int64_t b = bf / p.F;
      const int f = (int)(bf - b * p.F);
      const float g = f == 0 ? 1.f : p.gate[(b * (p.F - 1) + f - 1) * HW + pix];
      const float ft = to_f32<T>(FEAT[e * p.C + c]);
      float d = to_f32<T>(DG[((b * HW + pix) * p.F + f) * p.C + c]) * g;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yo = y - (ky - 1), xo = x - (kx - 1);
          if (yo >= 0 && yo < p.H && xo >= 0 && xo < p.W) {
            const float ds = p.dscore[bf * HW + (int64_t)yo * p.W + xo];
            d = fmaf(w[ky * 3 + kx], ds, d);
            acc[ky * 3 + kx] = fmaf(ft, ds, acc[ky * 3 + kx]);
          }
        }
      DF[e * p.C + c] = from_f32<T>(d);
    }
  }
  cta_reduce_store<9>(acc, sred, p.partial, p.C, c);
}

__global__ void __launch_bounds__(256) faf_gate_bwd_finish_kernel(const fbanet_faf_gate_bwd_params p, int blocks) {
  const int total = 9 * p.C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i / p.C, c = i % p.C;
    float s = 0.f;
    for (int z = 0; z < blocks; ++z) s += p.partial[((int64_t)z * 9 + v) * p.C + c];
    p.dwsum[i] = p.accumulate ? p.dwsum[i] + s : s;      // [9][C], the layout of wsum
  }
}

// ------------------------------------------------------------------------------------------------
// DropPath residual: out[b][i] = (skip ? skip[b][i] : 0) + scale[b] * x[b][i]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) drop_path_add_kernel(const fbanet_drop_path_params p) {
  const T* __restrict__ X = static_cast<const T*>(p.x);
  const T* __restrict__ S = static_cast<const T*>(p.skip);
  T* __restrict__ O = static_cast<T*>(p.out);
  const int64_t total = (int64_t)p.B * p.per_burst;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = p.scale[i / p.per_burst];
    // a dropped branch contributes exactly nothing (x * 0 would turn an inf / nan of the branch into nan; jnp does the same
    // multiplication, but the gradient of a dropped branch must not poison the flat buffer)
    const float v = s == 0.f ? 0.f : __fmul_rn(s, to_f32<T>(X[i]));      // product rounded before the add, as x * noise is
    O[i] = from_f32<T>(S ? __fadd_rn(to_f32<T>(S[i]), v) : v);
  }
}

// ------------------------------------------------------------------------------------------------
// window attention backward: one CTA per (window, head), thread i owns query row i in the row phases and key column i in the
// column phases.  P and dS live in shared memory ([N][N+1] fp32 each).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int shift_region_b(int v, int L, int win, int shift) { return v < L - win ? 0 : (v < L - shift ? 1 : 2); }

template <typename T, int DH>
__global__ void __launch_bounds__(128) window_attention_bwd_kernel(const fbanet_attn_bwd_params p) {
  extern __shared__ float smem[];
  const int win = p.win, N = win * win, LD = N + 1;
  float* qs = smem;                 // [N][DH]  q * scale
  float* ks = qs + N * DH;          // [N][DH]
  float* vs = ks + N * DH;          // [N][DH]
  float* gs = vs + N * DH;          // [N][DH]  dO
  float* Pm = gs + N * DH;          // [N][LD]  softmax probabilities
  float* dS = Pm + N * LD;          // [N][LD]  dP, then dS
  int* tok = reinterpret_cast<int*>(dS + N * LD);
  int* reg = tok + N;

  const int nwx = p.W / win, nwy = p.H / win;
  const int wid = blockIdx.x, head = blockIdx.y;
  const int b = wid / (nwx * nwy), wl = wid % (nwx * nwy);
  const int wy = wl / nwx, wx = wl % nwx;
  const int tid = threadIdx.x;
  const T* qkv = static_cast<const T*>(p.qkv);
  const T* dout = static_cast<const T*>(p.dout);
  T* dqkv = static_cast<T*>(p.dqkv);
  const int64_t img_tok0 = (int64_t)b * p.H * p.W;

  for (int i = tid; i < N; i += blockDim.x) {
    const int ys = wy * win + i / win, xs = wx * win + i % win;    // coordinates in the shifted grid
    const int y = (ys + p.shift) % p.H, x = (xs + p.shift) % p.W;
    tok[i] = y * p.W + x;
    reg[i] = p.shift > 0 ? shift_region_b(ys, p.H, win, p.shift) * 3 + shift_region_b(xs, p.W, win, p.shift) : 0;
  }
  __syncthreads();
  for (int e = tid; e < N * DH; e += blockDim.x) {
    const int j = e / DH, d = e % DH;
    const T* row = qkv + (img_tok0 + tok[j]) * p.qkv_ld + head * DH + d;
    qs[e] = to_f32<T>(row[0]) * p.scale;
    ks[e] = to_f32<T>(row[p.C]);
    vs[e] = to_f32<T>(row[2 * p.C]);
    gs[e] = to_f32<T>(dout[(img_tok0 + tok[j]) * p.dout_ld + head * DH + d]);
  }
  __syncthreads();

  const int i = tid;
  if (i < N) {
    float* Pi = Pm + i * LD;
    float* dSi = dS + i * LD;
    const int yi = i / win, xi = i % win, ri = reg[i];
    {   // probabilities of row i, exactly as the forward forms them
      float q[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) q[d] = qs[i * DH + d];
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
        Pi[j] = s;
        mx = fmaxf(mx, s);
      }
      float sum = 0.f;
      for (int j = 0; j < N; ++j) { const float e = expf(Pi[j] - mx); Pi[j] = e; sum += e; }
      const float inv = 1.0f / sum;
      for (int j = 0; j < N; ++j) Pi[j] *= inv;
    }
    {   // dP = dO v^T, delta = sum_j P dP, dS = P (dP - delta)
      float g[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) g[d] = gs[i * DH + d];
      float delta = 0.f;
      for (int j = 0; j < N; ++j) {
        const float* vj = vs + j * DH;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) s = fmaf(g[d], vj[d], s);
        dSi[j] = s;
        delta = fmaf(Pi[j], s, delta);
      }
      for (int j = 0; j < N; ++j) dSi[j] = Pi[j] * (dSi[j] - delta);
    }
    {   // dq_i = scale * sum_j dS_ij k_j
      float a[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) a[d] = 0.f;
      for (int j = 0; j < N; ++j) {
        const float s = dSi[j];
        const float* kj = ks + j * DH;
#pragma unroll
        for (int d = 0; d < DH; ++d) a[d] = fmaf(s, kj[d], a[d]);
      }
      T* o = dqkv + (img_tok0 + tok[i]) * p.dqkv_ld + head * DH;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = from_f32<T>(a[d] * p.scale);
    }
  }
  __syncthreads();
  if (i < N) {   // column phases: thread i is key / value row j = i
    const int j = i;
    float a[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) a[d] = 0.f;
    for (int r = 0; r < N; ++r) {        // dk_j = sum_r dS_rj (q_r * scale)
      const float s = dS[r * LD + j];
      const float* qr = qs + r * DH;
#pragma unroll
      for (int d = 0; d < DH; ++d) a[d] = fmaf(s, qr[d], a[d]);
    }
    T* o = dqkv + (img_tok0 + tok[j]) * p.dqkv_ld + p.C + head * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) { o[d] = from_f32<T>(a[d]); a[d] = 0.f; }
    for (int r = 0; r < N; ++r) {        // dv_j = sum_r P_rj dO_r
      const float s = Pm[r * LD + j];
      const float* gr = gs + r * DH;
#pragma unroll
      for (int d = 0; d < DH; ++d) a[d] = fmaf(s, gr[d], a[d]);
    }
    o += p.C;
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = from_f32<T>(a[d]);
  }
  if (p.dbias) {   // this window's share of the table gradient: row r = (dy, dx) collects dS over the pairs at that offset
    const int R1 = 2 * win - 1, R = R1 * R1;
    float* part = p.partial + ((int64_t)wid * p.heads + head) * R;
    for (int r = tid; r < R; r += blockDim.x) {
      const int oy = r / R1 - (win - 1), ox = r % R1 - (win - 1);   // yi - yj, xi - xj
      float s = 0.f;
      for (int q = 0; q < N; ++q) {
        const int yj = q / win - oy, xj = q % win - ox;
        if (yj >= 0 && yj < win && xj >= 0 && xj < win) s += dS[q * LD + yj * win + xj];
      }
      part[r] = s;
    }
  }
}

// dbias[r][h] (+)= sum over the windows, in order
__global__ void __launch_bounds__(256) window_attention_bwd_finish_kernel(const fbanet_attn_bwd_params p, int windows) {
  const int R = (2 * p.win - 1) * (2 * p.win - 1), total = R * p.heads;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / p.heads, h = i % p.heads;
    float s = 0.f;
    for (int w = 0; w < windows; ++w) s += p.partial[((int64_t)w * p.heads + h) * R + r];
    p.dbias[i] = p.accumulate ? p.dbias[i] + s : s;
  }
}

template <typename T, int DH>
static int launch_attn_bwd(const fbanet_attn_bwd_params* p, cudaStream_t s) {
  const int N = p->win * p->win;
  const size_t smem = (size_t)(4 * N * DH + 2 * N * (N + 1)) * sizeof(float) + 2 * N * sizeof(int);
  if (smem > 227 * 1024) return FBANET_E_BADSHAPE;
  auto kern = window_attention_bwd_kernel<T, DH>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
  const int windows = p->B * (p->H / p->win) * (p->W / p->win);
  kern<<<dim3(windows, p->heads), 128, smem, s>>>(*p);
  if (p->dbias) {
    const int total = (2 * p->win - 1) * (2 * p->win - 1) * p->heads;
    window_attention_bwd_finish_kernel<<<(total + 255) / 256, 256, 0, s>>>(*p, windows);
  }
  return check_launch();
}

template <typename T>
static int dispatch_attn_bwd(const fbanet_attn_bwd_params* p, cudaStream_t s) {
  switch (p->C / p->heads) {
    case 8: return launch_attn_bwd<T, 8>(p, s);
    case 16: return launch_attn_bwd<T, 16>(p, s);
    case 32: return launch_attn_bwd<T, 32>(p, s);
    case 64: return launch_attn_bwd<T, 64>(p, s);
    default: return FBANET_E_BADSHAPE;
  }
}

static int spatial_blocks(int64_t units) {
  const int64_t b = cdiv64s(units, SP_ROWS);
  return (int)(b < SP_MAX_BLOCKS_X ? b : SP_MAX_BLOCKS_X);
}

}  // namespace fbanet

using namespace fbanet;

extern "C" int fbanet_dwconv_bwd_blocks(int64_t pixels) { return pixels <= 0 ? -1 : spatial_blocks(pixels); }

extern "C" int fbanet_dwconv3x3_bwd_sm100(const fbanet_dwconv_bwd_params* p, void* stream) {
  if (!p || !p->x || !p->dy || !p->weight || !p->partial || p->N <= 0 || p->H <= 0 || p->W <= 0 || p->C <= 0) return FBANET_E_BADSHAPE;
  if (!p->dx && !p->dw && !p->db) return FBANET_E_BADSHAPE;
  if (p->dtype != FBANET_F32 && p->dtype != FBANET_BF16) return FBANET_E_DTYPE;
  const int blocks = spatial_blocks((int64_t)p->N * p->H * p->W);
  const dim3 grid((unsigned)blocks, (unsigned)((p->C + SP_CT - 1) / SP_CT));
  if (grid.y > 65535) return FBANET_E_BADSHAPE;
  if (p->dtype == FBANET_F32) dwconv_bwd_kernel<float><<<grid, SP_THREADS, 0, (cudaStream_t)stream>>>(*p);
  else dwconv_bwd_kernel<bf16><<<grid, SP_THREADS, 0, (cudaStream_t)stream>>>(*p);
  if (p->dw || p->db) dwconv_bwd_finish_kernel<<<(10 * p->C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, blocks);
  return check_launch();
}

extern "C" int fbanet_faf_gate_bwd_blocks(int64_t frame_pixels) { return frame_pixels <= 0 ? -1 : spatial_blocks(frame_pixels); }

extern "C" int fbanet_faf_gate_bwd_sm100(const fbanet_faf_gate_bwd_params* p, void* stream) {
  if (!p || !p->feat || !p->dgated || !p->gate || !p->score || !p->wsum || !p->dfeat || !p->dscore || !p->dwsum || !p->partial)
    return FBANET_E_BADSHAPE;
  if (p->B <= 0 || p->F < 2 || p->H <= 0 || p->W <= 0 || p->C <= 0) return FBANET_E_BADSHAPE;
  if (p->dtype != FBANET_F32 && p->dtype != FBANET_BF16) return FBANET_E_DTYPE;
  const int64_t BP = (int64_t)p->B * p->H * p->W;
  int64_t b1 = cdiv64s(BP, 8);
  if (b1 > 148 * 8) b1 = 148 * 8;
  const int blocks = spatial_blocks(BP * p->F);
  const dim3 grid((unsigned)blocks, (unsigned)((p->C + SP_CT - 1) / SP_CT));
  if (p->dtype == FBANET_F32) {
    faf_gate_bwd_dscore_kernel<float><<<(unsigned)b1, 256, 0, (cudaStream_t)stream>>>(*p);
    faf_gate_bwd_dfeat_kernel<float><<<grid, SP_THREADS, 0, (cudaStream_t)stream>>>(*p);
  } else {
    faf_gate_bwd_dscore_kernel<bf16><<<(unsigned)b1, 256, 0, (cudaStream_t)stream>>>(*p);
    faf_gate_bwd_dfeat_kernel<bf16><<<grid, SP_THREADS, 0, (cudaStream_t)stream>>>(*p);
  }
  faf_gate_bwd_finish_kernel<<<(9 * p->C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, blocks);
  return check_launch();
}

extern "C" int fbanet_drop_path_add_sm100(const fbanet_drop_path_params* p, void* stream) {
  if (!p || !p->x || !p->out || !p->scale || p->B <= 0 || p->per_burst <= 0) return FBANET_E_BADSHAPE;
  if (p->dtype != FBANET_F32 && p->dtype != FBANET_BF16) return FBANET_E_DTYPE;
  int64_t blocks = cdiv64s((int64_t)p->B * p->per_burst, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (p->dtype == FBANET_F32) drop_path_add_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  else drop_path_add_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}

extern "C" int64_t fbanet_attn_bwd_partial_floats(int32_t B, int32_t H, int32_t W, int32_t heads, int32_t win) {
  if (B <= 0 || H <= 0 || W <= 0 || heads <= 0 || win <= 0 || H % win || W % win) return -1;
  return (int64_t)B * (H / win) * (W / win) * heads * (2 * win - 1) * (2 * win - 1);
}

extern "C" int fbanet_window_attention_bwd_sm100(const fbanet_attn_bwd_params* p, void* stream) {
  if (!p || !p->qkv || !p->dout || !p->dqkv || !p->bias_table) return FBANET_E_BADSHAPE;
  if (p->dbias && !p->partial) return FBANET_E_BADSHAPE;
  if (p->B <= 0 || p->heads <= 0 || p->win <= 0 || p->C <= 0 || p->C % p->heads) return FBANET_E_BADSHAPE;
  if (p->H <= 0 || p->W <= 0 || p->H % p->win || p->W % p->win || p->shift < 0 || p->shift >= p->win) return FBANET_E_BADSHAPE;
  if (p->win * p->win > 128) return FBANET_E_BADSHAPE;
  if (p->qkv_ld < 3 * p->C || p->dqkv_ld < 3 * p->C || p->dout_ld < p->C) return FBANET_E_BADSHAPE;
  if (p->dtype == FBANET_F32) return dispatch_attn_bwd<float>(p, (cudaStream_t)stream);
  if (p->dtype == FBANET_BF16) return dispatch_attn_bwd<bf16>(p, (cudaStream_t)stream);
  return FBANET_E_DTYPE;
}
