// K6 (tensor-core version, bf16): windowed multi-head self-attention, one CTA = one head x a chunk of
// windows, one warp = 16 query rows.  S = (q*scale) k^T and O = P v run on warp-level bf16 MMAs with the
// score tile resident in registers (the softmax'd S fragments are re-used directly as the A operand of
// P v, so S/P never touch shared memory).  The contraction depth per head is only d_h = 16 or 64, so the
// op is bound by the softmax (MUFU exp + fp32 ALU), not by the tensor pipe -- see DESIGN.md "K6".
//
// Relative-position bias (+ -inf key padding) for the CTA's head is expanded once per CTA into shared
// memory [N][NP]; the Swin shift mask is evaluated analytically and only for windows that touch the
// wrapped edge.
#include <stdlib.h>

#include "common.cuh"

namespace fbanet {

__device__ __forceinline__ int shift_region_tc(int v, int L, int win, int shift) { return v < L - win ? 0 : (v < L - shift ? 1 : 2); }

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1])
               : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}

// NT = number of 8-key tiles (keys padded to 8*NT, a multiple of 16); MT = number of 16-row query tiles.
// K and V of a window are staged row-major [key][DH+8] with cp.async into a double buffer: the copies for
// window i+1 are in flight while window i is computed.  Fragments come from ldmatrix (.trans for V).
// Column DH of the V tile holds 1.0, so one extra MMA per key step yields the softmax denominator
// (sum of the bf16-rounded probabilities that P.v actually used) instead of 56 FADDs + shuffles per thread.
// Windows that do not wrap around the image edge take a fast path: token = window base + per-thread constant.
template <int DH, int NT, int MT, bool ONES>
__global__ void __launch_bounds__(MT * 32, (MT == 7) ? (DH == 16 ? 3 : 2) : 1) window_attention_tc_kernel(const fbanet_attn_params p, const int win_chunk) {
  constexpr int NP = NT * 8;          // padded keys
  constexpr int KS = DH + 8;          // row stride (elements): conflict-free ldmatrix rows; column DH = ones column of V
  constexpr int CPR = DH / 8;         // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int win = p.win, N = win * win;
  float* biasS = reinterpret_cast<float*>(smem_attn);                 // [N][NP] bias in log2 units (+ -inf for padded keys)
  bf16* KV = reinterpret_cast<bf16*>(biasS + N * NP);                 // [2 buffers][K | V][NP][KS]
  int* tokS = reinterpret_cast<int*>(KV + 4 * NP * KS);               // [2][NP] token index inside the image (wrapping windows)
  int* regS = tokS + 2 * NP;                                          // [2][NP] shift-mask region id

  const int head = blockIdx.x;   // heads fastest: the CTAs sharing a window's token rows run together (L2 reuse)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nwx = p.W / win, nwy = p.H / win, nw_img = nwx * nwy;
  const int total_windows = p.B * nw_img;
  const bf16* qkv = reinterpret_cast<const bf16*>(p.qkv) + head * DH;
  bf16* outp = reinterpret_cast<bf16*>(p.out) + head * DH + 2 * t;

  const int w_begin = blockIdx.y * win_chunk;
  const int w_end = min(w_begin + win_chunk, total_windows);

  // per-thread staging slots are the same for every window: (key j, 16-byte chunk c, in-window y/x)
  constexpr int SLOTS = (NP * CPR + MT * 32 - 1) / (MT * 32);
  int sj[SLOTS], sc[SLOTS], siy[SLOTS], six[SLOTS];
#pragma unroll
  for (int k = 0; k < SLOTS; ++k) {
    const int e = tid + k * MT * 32;
    sj[k] = e / CPR;
    sc[k] = (e - sj[k] * CPR) * 8;
    siy[k] = sj[k] / win;
    six[k] = sj[k] - siy[k] * win;
  }
  // this thread's two query rows
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const bool v0 = r0 < N, v1 = r1 < N;
  const int q0y = (v0 ? r0 : 0) / win, q0x = (v0 ? r0 : 0) % win, q1y = (v1 ? r1 : 0) / win, q1x = (v1 ? r1 : 0) % win;

  // K,V copies (and, for wrapping windows, token/region tables) of window `wid` into buffer `buf`
  auto stage = [&](int wid, int buf) {
    const int b = wid / nw_img, wl = wid - b * nw_img;
    const int wy = wl / nwx, wx = wl - wy * nwx;
    const bool wrap = p.shift > 0 && (wy == nwy - 1 || wx == nwx - 1);
    const int ty0 = wy * win + p.shift, tx0 = wx * win + p.shift;
    const bf16* ibase = qkv + (int64_t)b * p.H * p.W * p.qkv_ld;
    bf16* Kb = KV + (size_t)buf * 2 * NP * KS;
    bf16* Vb = Kb + NP * KS;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int j = sj[k], c = sc[k];
      if (j >= NP) continue;
      if (j < N) {
        int y = ty0 + siy[k], x = tx0 + six[k];
        if (wrap) {
          if (y >= p.H) y -= p.H;
          if (x >= p.W) x -= p.W;
          if (c == 0) {
            tokS[buf * NP + j] = y * p.W + x;
            regS[buf * NP + j] = shift_region_tc(wy * win + siy[k], p.H, win, p.shift) * 3 + shift_region_tc(wx * win + six[k], p.W, win, p.shift);
          }
        }
        const bf16* row = ibase + (int64_t)(y * p.W + x) * p.qkv_ld + c;
        cp_async16(Kb + j * KS + c, row + p.C);
        cp_async16(Vb + j * KS + c, row + 2 * p.C);
        if (c == 0) *reinterpret_cast<uint4*>(Vb + j * KS + DH) = make_uint4(0x00003F80u, 0, 0, 0);   // bf16 1.0 in column DH
      } else {
        *reinterpret_cast<uint4*>(Kb + j * KS + c) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(Vb + j * KS + c) = make_uint4(0, 0, 0, 0);
        if (c == 0) *reinterpret_cast<uint4*>(Vb + j * KS + DH) = make_uint4(0, 0, 0, 0);
      }
    }
    cp_async_commit();
  };

  if (w_begin < w_end) stage(w_begin, 0);
  // ---- relative-position bias of this head -> smem, in log2 units (overlaps the first copies) ----
  if (p.bias_expanded) {   // host-expanded [heads][N][NP]
    const float4* src = reinterpret_cast<const float4*>(p.bias_expanded + (size_t)head * N * NP);
    for (int e = tid; e < N * NP / 4; e += blockDim.x) reinterpret_cast<float4*>(biasS)[e] = __ldg(src + e);
  } else {
    for (int e = tid; e < N * NP; e += blockDim.x) {
      const int i = e / NP, j = e - i * NP;
      float b = -1e30f;
      if (j < N) {
        const int yi = i / win, xi = i - yi * win, yj = j / win, xj = j - yj * win;
        b = __ldg(p.bias_table + ((yi - yj + win - 1) * (2 * win - 1) + (xi - xj + win - 1)) * p.heads + head);
      }
      biasS[e] = b * 1.4426950408889634f;   // exp(x) = 2^(x log2 e)
    }
  }
  const float scale2 = p.q_prescaled ? 1.f : p.scale * 1.4426950408889634f;
  const float* bb0p = biasS + (v0 ? r0 : 0) * NP + 2 * t;
  const float* bb1p = biasS + (v1 ? r1 : 0) * NP + 2 * t;

  for (int wid = w_begin; wid < w_end; ++wid) {
    const int buf = (wid - w_begin) & 1;
    cp_async_wait_all();
    __syncthreads();  // window `wid` staged; every warp is done with the other buffer
    if (wid + 1 < w_end) stage(wid + 1, buf ^ 1);

    const int b = wid / nw_img, wl = wid - b * nw_img;
    const int wy = wl / nwx, wx = wl - wy * nwx;
    const bool wrap = p.shift > 0 && (wy == nwy - 1 || wx == nwx - 1);
    const int64_t img_tok0 = (int64_t)b * p.H * p.W;
    const bf16* Kb = KV + (size_t)buf * 2 * NP * KS;
    const bf16* Vb = Kb + NP * KS;
    const int* tok = tokS + buf * NP;
    const int* reg = regS + buf * NP;
    int tok0, tok1;
    if (wrap) {
      tok0 = tok[v0 ? r0 : 0];
      tok1 = tok[v1 ? r1 : 0];
    } else {
      const int wb = (wy * win + p.shift) * p.W + wx * win + p.shift;
      tok0 = wb + q0y * p.W + q0x;
      tok1 = wb + q1y * p.W + q1x;
    }

    // ---- this warp's 16 query rows ----
    uint32_t qa[DH / 16][4];
    {
      const bf16* q0 = qkv + (img_tok0 + tok0) * p.qkv_ld + 2 * t;
      const bf16* q1 = qkv + (img_tok0 + tok1) * p.qkv_ld + 2 * t;
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        qa[kk][0] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + kk * 16) : 0u;
        qa[kk][1] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + kk * 16) : 0u;
        qa[kk][2] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 8) : 0u;
        qa[kk][3] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 8) : 0u;
      }
    }
    // S = q k^T  (scale applied afterwards in fp32).  One ldmatrix.x4 = B fragments of two key tiles.
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    const int mi = lane >> 3;
    const bf16* kfrag = Kb + ((mi >> 1) * 8 + (lane & 7)) * KS + (mi & 1) * 8;
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        // matrices: (keys nt*8.., ch kk*16..+7), (same keys, ch +8), (keys (nt+1)*8.., ch ..+7), (.., ch +8)
        uint32_t kb[4];
        ldmatrix_x4(kb, kfrag + nt * 8 * KS + kk * 16);
        const uint32_t b0[2] = {kb[0], kb[1]}, b1[2] = {kb[2], kb[3]};
        mma_bf16_16816(s[nt], qa[kk], b0);
        mma_bf16_16816(s[nt + 1], qa[kk], b1);
      }
    }
    // + bias (+ mask), row max
    float m0 = -1e30f, m1 = -1e30f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float2 bb0 = *reinterpret_cast<const float2*>(bb0p + nt * 8);
      const float2 bb1 = *reinterpret_cast<const float2*>(bb1p + nt * 8);
      s[nt][0] = fmaf(s[nt][0], scale2, bb0.x);
      s[nt][1] = fmaf(s[nt][1], scale2, bb0.y);
      s[nt][2] = fmaf(s[nt][2], scale2, bb1.x);
      s[nt][3] = fmaf(s[nt][3], scale2, bb1.y);
    }
    if (wrap) {   // Swin shift mask, only where the window straddles the wrapped edge
      constexpr float M = 100.f * 1.4426950408889634f;
      const int rg0 = reg[v0 ? r0 : 0], rg1 = reg[v1 ? r1 : 0];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int ja = reg[nt * 8 + 2 * t], jb = reg[nt * 8 + 2 * t + 1];
        if (ja != rg0) s[nt][0] -= M;
        if (jb != rg0) s[nt][1] -= M;
        if (ja != rg1) s[nt][2] -= M;
        if (jb != rg1) s[nt][3] -= M;
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    // O = P v and l = P 1 ; V fragments via ldmatrix.trans from the row-major [key][ch] tile
    float o[DH / 8][4], osum[4] = {0.f, 0.f, 0.f, 0.f}, lsum0 = 0.f, lsum1 = 0.f;
#pragma unroll
    for (int dn = 0; dn < DH / 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    const bf16* vfrag = Vb + ((mi & 1) * 8 + (lane & 7)) * KS + (mi >> 1) * 8;
    const bf16* vones = Vb + (lane & 15) * KS + DH;
#pragma unroll
    for (int ks = 0; ks < NT / 2; ++ks) {
      uint32_t pa[4];
      float e[8];
      e[0] = ex2_approx(s[2 * ks][0] - m0); e[1] = ex2_approx(s[2 * ks][1] - m0);
      e[2] = ex2_approx(s[2 * ks][2] - m1); e[3] = ex2_approx(s[2 * ks][3] - m1);
      e[4] = ex2_approx(s[2 * ks + 1][0] - m0); e[5] = ex2_approx(s[2 * ks + 1][1] - m0);
      e[6] = ex2_approx(s[2 * ks + 1][2] - m1); e[7] = ex2_approx(s[2 * ks + 1][3] - m1);
      if (!ONES) { lsum0 += (e[0] + e[1]) + (e[4] + e[5]); lsum1 += (e[2] + e[3]) + (e[6] + e[7]); }
      pa[0] = pack_bf16(e[0], e[1]);
      pa[1] = pack_bf16(e[2], e[3]);
      pa[2] = pack_bf16(e[4], e[5]);
      pa[3] = pack_bf16(e[6], e[7]);
#pragma unroll
      for (int dn = 0; dn < DH / 8; dn += 2) {
        // matrices: (keys 16ks..+7, ch dn*8), (keys +8.., ch dn*8), (keys 16ks.., ch (dn+1)*8), (keys +8.., ch (dn+1)*8)
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, vfrag + ks * 16 * KS + dn * 8);
        const uint32_t b0[2] = {vb[0], vb[1]}, b1[2] = {vb[2], vb[3]};
        mma_bf16_16816(o[dn], pa, b0);
        mma_bf16_16816(o[dn + 1], pa, b1);
      }
      if (ONES) {
        uint32_t ob[2];
        ldmatrix_x2_trans(ob, vones + ks * 16 * KS);
        mma_bf16_16816(osum, pa, ob);
      }
    }
    float l0, l1;
    if (ONES) {   // column 0 of the ones tile lives in the t == 0 lane of each quad
      l0 = __shfl_sync(0xffffffffu, osum[0], lane & ~3);
      l1 = __shfl_sync(0xffffffffu, osum[2], lane & ~3);
    } else {
      lsum0 += __shfl_xor_sync(0xffffffffu, lsum0, 1);
      lsum0 += __shfl_xor_sync(0xffffffffu, lsum0, 2);
      lsum1 += __shfl_xor_sync(0xffffffffu, lsum1, 1);
      lsum1 += __shfl_xor_sync(0xffffffffu, lsum1, 2);
      l0 = lsum0; l1 = lsum1;
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    if (v0) {
      bf16* op = outp + (img_tok0 + tok0) * p.out_ld;
#pragma unroll
      for (int dn = 0; dn < DH / 8; ++dn) *reinterpret_cast<__nv_bfloat162*>(op + dn * 8) = __floats2bfloat162_rn(o[dn][0] * i0, o[dn][1] * i0);
    }
    if (v1) {
      bf16* op = outp + (img_tok0 + tok1) * p.out_ld;
#pragma unroll
      for (int dn = 0; dn < DH / 8; ++dn) *reinterpret_cast<__nv_bfloat162*>(op + dn * 8) = __floats2bfloat162_rn(o[dn][2] * i1, o[dn][3] * i1);
    }
  }
  cp_async_wait_all();
}

template <int DH, int NT, int MT, bool ONES>
static int launch_tc(const fbanet_attn_params* p, cudaStream_t s) {
  constexpr int NP = NT * 8;
  const int N = p->win * p->win;
  const size_t smem = (size_t)N * NP * 4 + (size_t)4 * NP * (DH + 8) * 2 + (size_t)4 * NP * 4;
  auto kern = window_attention_tc_kernel<DH, NT, MT, ONES>;
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
  dim3 grid(p->heads, (total_windows + chunk - 1) / chunk);
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
    case 100: {
      static const char* ones = getenv("FBANET_ATTN_ONES");   // experiment switch: softmax denominator via a ones column of V
      return (ones && ones[0] == '1') ? launch_tc<DH, 14, 7, true>(p, s) : launch_tc<DH, 14, 7, false>(p, s);
    }
    case 64: return launch_tc<DH, 8, 4, false>(p, s);
    case 25: return launch_tc<DH, 4, 2, false>(p, s);
    case 16: return launch_tc<DH, 2, 1, false>(p, s);
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
