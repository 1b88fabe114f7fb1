// K6 (tensor-core version, bf16): windowed multi-head self-attention, one CTA = one head x a chunk of
// windows, one warp = 16 query rows.  S = (q*scale) k^T and O = P v run on warp-level bf16 MMAs with the
// score tile resident in registers (the softmax'd S fragments are re-used directly as the A operand of
// P v, so S/P never touch shared memory).  The contraction depth per head is only d_h = 16 or 64, so the
// op is bound by the softmax (MUFU exp + fp32 ALU), not by the tensor pipe -- see DESIGN.md "K6".
//
// Relative-position bias (+ -inf key padding) for the CTA's head is expanded once per CTA into shared
// memory [N][NP]; the Swin shift mask is evaluated analytically and only for windows that touch the
// wrapped edge.
#include "common.cuh"

namespace fbanet {

__device__ __forceinline__ int shift_region_tc(int v, int L, int win, int shift) { return v < L - win ? 0 : (v < L - shift ? 1 : 2); }

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// NT = number of 8-key tiles (keys padded to 8*NT, a multiple of 16); MT = number of 16-row query tiles
template <int DH, int NT, int MT>
__global__ void __launch_bounds__(MT * 32) window_attention_tc_kernel(const fbanet_attn_params p, const int win_chunk) {
  constexpr int NP = NT * 8;          // padded keys
  constexpr int KS = DH + 8;          // K row stride (elements): conflict-free fragment loads
  constexpr int VS = NP + 8;          // V^T row stride
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int win = p.win, N = win * win;
  float* biasS = reinterpret_cast<float*>(smem_attn);                 // [N][NP] bias (+ -inf for padded keys)
  bf16* Ks = reinterpret_cast<bf16*>(biasS + N * NP);                 // [NP][KS]
  bf16* Vt = Ks + NP * KS;                                            // [DH][VS]
  int* tok = reinterpret_cast<int*>(Vt + DH * VS);                    // [NP]
  int* reg = tok + NP;                                                // [NP]

  const int head = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nwx = p.W / win, nwy = p.H / win, nw_img = nwx * nwy;
  const int total_windows = p.B * nw_img;
  const bf16* qkv = reinterpret_cast<const bf16*>(p.qkv);

  // ---- expand the relative-position bias of this head once ----
  for (int e = tid; e < N * NP; e += blockDim.x) {
    const int i = e / NP, j = e - i * NP;
    float b = -1e30f;
    if (j < N) {
      const int yi = i / win, xi = i - yi * win, yj = j / win, xj = j - yj * win;
      b = __ldg(p.bias_table + ((yi - yj + win - 1) * (2 * win - 1) + (xi - xj + win - 1)) * p.heads + head);
    }
    biasS[e] = b;
  }

  const int w_begin = blockIdx.x * win_chunk;
  const int w_end = min(w_begin + win_chunk, total_windows);
  for (int wid = w_begin; wid < w_end; ++wid) {
    const int b = wid / nw_img, wl = wid - b * nw_img;
    const int wy = wl / nwx, wx = wl - wy * nwx;
    const int64_t img_tok0 = (int64_t)b * p.H * p.W;
    const bool masked = p.shift > 0 && (wy == nwy - 1 || wx == nwx - 1);
    __syncthreads();  // previous window fully consumed (and biasS visible on the first trip)
    for (int i = tid; i < NP; i += blockDim.x) {
      int tk = 0, rg = 0;
      if (i < N) {
        const int iy = i / win, ix = i - iy * win;
        const int ys = wy * win + iy, xs = wx * win + ix;
        const int y = (ys + p.shift) % p.H, x = (xs + p.shift) % p.W;
        tk = y * p.W + x;
        rg = p.shift > 0 ? shift_region_tc(ys, p.H, win, p.shift) * 3 + shift_region_tc(xs, p.W, win, p.shift) : 0;
      }
      tok[i] = tk;
      reg[i] = rg;
    }
    __syncthreads();
    // ---- K -> Ks[key][ch], V -> Vt[ch][key]; 16-byte global loads ----
    constexpr int VEC = 8, CPR = DH / VEC;  // 16-byte chunks per row
    for (int e = tid; e < NP * CPR; e += blockDim.x) {
      const int j = e / CPR, c = (e - j * CPR) * VEC;
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (j < N) {
        const bf16* row = qkv + (img_tok0 + tok[j]) * p.qkv_ld + head * DH + c;
        kv = *reinterpret_cast<const uint4*>(row + p.C);
        vv = *reinterpret_cast<const uint4*>(row + 2 * p.C);
      }
      *reinterpret_cast<uint4*>(Ks + j * KS + c) = kv;
      const bf16* ve = reinterpret_cast<const bf16*>(&vv);
#pragma unroll
      for (int u = 0; u < VEC; ++u) Vt[(c + u) * VS + j] = ve[u];
    }
    __syncthreads();

    // ---- this warp's 16 query rows ----
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < N, v1 = r1 < N;
    uint32_t qa[DH / 16][4];
    {
      const bf16* q0 = qkv + (img_tok0 + tok[v0 ? r0 : 0]) * p.qkv_ld + head * DH;
      const bf16* q1 = qkv + (img_tok0 + tok[v1 ? r1 : 0]) * p.qkv_ld + head * DH;
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const int c = kk * 16 + 2 * t;
        const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
        __nv_bfloat162 x00 = v0 ? *reinterpret_cast<const __nv_bfloat162*>(q0 + c) : z;
        __nv_bfloat162 x10 = v1 ? *reinterpret_cast<const __nv_bfloat162*>(q1 + c) : z;
        __nv_bfloat162 x01 = v0 ? *reinterpret_cast<const __nv_bfloat162*>(q0 + c + 8) : z;
        __nv_bfloat162 x11 = v1 ? *reinterpret_cast<const __nv_bfloat162*>(q1 + c + 8) : z;
        qa[kk][0] = *reinterpret_cast<uint32_t*>(&x00);
        qa[kk][1] = *reinterpret_cast<uint32_t*>(&x10);
        qa[kk][2] = *reinterpret_cast<uint32_t*>(&x01);
        qa[kk][3] = *reinterpret_cast<uint32_t*>(&x11);
      }
    }
    // S = q k^T  (scale applied afterwards in fp32: (q.k)*scale == (q*scale).k up to rounding of the product)
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        uint32_t bfrag[2];
        const bf16* kp = Ks + (nt * 8 + g) * KS + kk * 16 + 2 * t;
        bfrag[0] = *reinterpret_cast<const uint32_t*>(kp);
        bfrag[1] = *reinterpret_cast<const uint32_t*>(kp + 8);
        mma_bf16_16816(s[nt], qa[kk], bfrag);
      }
    }
    // + bias (+ mask), row max
    const float* b0 = biasS + (v0 ? r0 : 0) * NP;
    const float* b1 = biasS + (v1 ? r1 : 0) * NP;
    const int rg0 = reg[v0 ? r0 : 0], rg1 = reg[v1 ? r1 : 0];
    float m0 = -1e30f, m1 = -1e30f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int j = nt * 8 + 2 * t;
      const float2 bb0 = *reinterpret_cast<const float2*>(b0 + j);
      const float2 bb1 = *reinterpret_cast<const float2*>(b1 + j);
      s[nt][0] = fmaf(s[nt][0], p.scale, bb0.x);
      s[nt][1] = fmaf(s[nt][1], p.scale, bb0.y);
      s[nt][2] = fmaf(s[nt][2], p.scale, bb1.x);
      s[nt][3] = fmaf(s[nt][3], p.scale, bb1.y);
      if (masked) {
        const int ja = reg[j], jb = reg[j + 1];
        if (ja != rg0) s[nt][0] -= 100.f;
        if (jb != rg0) s[nt][1] -= 100.f;
        if (ja != rg1) s[nt][2] -= 100.f;
        if (jb != rg1) s[nt][3] -= 100.f;
      }
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
    constexpr float LOG2E = 1.4426950408889634f;
    const float mm0 = m0 * LOG2E, mm1 = m1 * LOG2E;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = exp2f(fmaf(s[nt][0], LOG2E, -mm0));
      s[nt][1] = exp2f(fmaf(s[nt][1], LOG2E, -mm0));
      s[nt][2] = exp2f(fmaf(s[nt][2], LOG2E, -mm1));
      s[nt][3] = exp2f(fmaf(s[nt][3], LOG2E, -mm1));
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // O = P v
    float o[DH / 8][4];
#pragma unroll
    for (int dn = 0; dn < DH / 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < NT / 2; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int dn = 0; dn < DH / 8; ++dn) {
        uint32_t bfrag[2];
        const bf16* vp = Vt + (dn * 8 + g) * VS + ks * 16 + 2 * t;
        bfrag[0] = *reinterpret_cast<const uint32_t*>(vp);
        bfrag[1] = *reinterpret_cast<const uint32_t*>(vp + 8);
        mma_bf16_16816(o[dn], pa, bfrag);
      }
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    bf16* out = reinterpret_cast<bf16*>(p.out);
    if (v0) {
      bf16* op = out + (img_tok0 + tok[r0]) * p.out_ld + head * DH + 2 * t;
#pragma unroll
      for (int dn = 0; dn < DH / 8; ++dn) *reinterpret_cast<__nv_bfloat162*>(op + dn * 8) = __floats2bfloat162_rn(o[dn][0] * i0, o[dn][1] * i0);
    }
    if (v1) {
      bf16* op = out + (img_tok0 + tok[r1]) * p.out_ld + head * DH + 2 * t;
#pragma unroll
      for (int dn = 0; dn < DH / 8; ++dn) *reinterpret_cast<__nv_bfloat162*>(op + dn * 8) = __floats2bfloat162_rn(o[dn][2] * i1, o[dn][3] * i1);
    }
  }
}

template <int DH, int NT, int MT>
static int launch_tc(const fbanet_attn_params* p, cudaStream_t s) {
  constexpr int NP = NT * 8;
  const int N = p->win * p->win;
  const size_t smem = (size_t)N * NP * 4 + (size_t)NP * (DH + 8) * 2 + (size_t)DH * (NP + 8) * 2 + 2 * NP * 4;
  auto kern = window_attention_tc_kernel<DH, NT, MT>;
  static size_t opted = 0;
  if (smem > opted) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
    opted = smem;
  }
  const int total_windows = p->B * (p->H / p->win) * (p->W / p->win);
  // enough CTAs for ~4 waves of 148 SMs x 2 resident CTAs, but >= 4 windows per CTA to amortise the bias expansion
  int chunk = (total_windows * p->heads + 1183) / 1184;
  if (chunk < 4) chunk = 4;
  if (chunk > 32) chunk = 32;
  dim3 grid((total_windows + chunk - 1) / chunk, p->heads);
  kern<<<grid, MT * 32, smem, s>>>(*p, chunk);
  return check_launch();
}

int window_attention_tc_supported(const fbanet_attn_params* p) {
  if (p->dtype != FBANET_BF16) return 0;
  const int dh = p->C / p->heads, N = p->win * p->win;
  if (dh != 16 && dh != 32 && dh != 64) return 0;
  if (N != 100 && N != 64 && N != 25 && N != 16) return 0;
  if ((p->qkv_ld % 8) || (p->out_ld % 2) || ((uintptr_t)p->qkv % 16) || ((uintptr_t)p->out % 4)) return 0;
  return 1;
}

template <int DH>
static int dispatch_n(const fbanet_attn_params* p, cudaStream_t s) {
  switch (p->win * p->win) {
    case 100: return launch_tc<DH, 14, 7>(p, s);
    case 64: return launch_tc<DH, 8, 4>(p, s);
    case 25: return launch_tc<DH, 4, 2>(p, s);
    case 16: return launch_tc<DH, 2, 1>(p, s);
    default: return FBANET_E_UNSUPPORTED;
  }
}

int window_attention_tc_launch(const fbanet_attn_params* p, cudaStream_t s) {
  if (!window_attention_tc_supported(p)) return FBANET_E_UNSUPPORTED;
  switch (p->C / p->heads) {
    case 16: return dispatch_n<16>(p, s);
    case 32: return dispatch_n<32>(p, s);
    case 64: return dispatch_n<64>(p, s);
    default: return FBANET_E_UNSUPPORTED;
  }
}

}  // namespace fbanet
