// K6, head_dim = 16 / window 10x10 specialisation (bf16): the decoder and bottleneck stages of the model
// (dim 256 x 16 heads, dim 128 x 8 heads -- 75 % of all score elements of a forward).
//
// One WARP owns one (window, head) item end to end; a CTA is 12 independent warps working on windows of the
// same head, so the only CTA-wide barrier is the one after the head's relative-position bias has been
// expanded into shared memory.  Per item the warp
//   * stages the head's q | k | v rows of the (cyclically shifted) window with cp.async -- 100 rows x 32 B each,
//     XOR-swizzled so every ldmatrix phase is bank-conflict free (the next item's rows are in flight while this
//     one computes; q is double buffered, k/v are refilled as soon as their fragments are in registers);
//   * keeps ALL of K (14 key tiles) and V (7 key steps x 2 channel tiles) as MMA B fragments in registers
//     (56 registers) and walks the 7 query tiles: S = bias + q k^T is ONE m16n8k16 MMA per key tile whose
//     accumulator is initialised with the bias fragment (stored fragment-major in smem: one LDS.128 per tile),
//     softmax in registers, O = P v with the probabilities re-used as the A operand, and the row sum from one
//     extra MMA against an all-ones B fragment;
//   * writes O over the q rows it has consumed and streams it out with 16-byte stores.
// The q columns of the fused qkv GEMM may be pre-multiplied by scale*log2(e) (`q_prescaled`), which removes
// the per-score FFMA.  Bound: MUFU.EX2 (one per score) -- see DESIGN.md "K6".
#include <stdlib.h>
#include <initializer_list>
#include <type_traits>

#include "common.cuh"

namespace fbanet {

namespace {

constexpr int WIN = 10, NTOK = 100, NT = 14, MT = 7, WARPS = 12;
constexpr int NTU = 13;                       // key tiles that hold real keys (tile 13 = keys 104..111 is padding)
constexpr int ROWB = 32;                      // bytes per staged row (16 bf16)
constexpr int TILE_BYTES = 112 * ROWB;        // one q / k / v buffer
constexpr int WARP_BYTES = 4 * TILE_BYTES;    // q0 | q1 | k | v
constexpr int BIAS_BYTES = MT * NT * 32 * 16; // float4 per (query tile, key tile, lane)
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // volatile: keeps the hand-made MUFU / MMA interleave
  return y;
}
// exp2 of two scores on the FMA pipe (no MUFU): x = n + f with n = round(x) taken from the low mantissa bits of x + 1.5 * 2^23,
// 2^f on [-0.5, 0.5] as a degree-3 minimax polynomial (relative error 7.5e-5 = 2^-13.7: the result is rounded to bf16, 2^-9, right
// after), and n added to the exponent field.  Scores are clamped to [-125, 126] first: the -1e30 of the padded keys and the shift
// mask become 2^-125 (nothing), and a logit beyond 2^60 still drives the row sum out of the band that triggers the exact path.
__device__ __forceinline__ uint32_t ex2_poly_pack2(float x0, float x1) {
  x0 = fminf(fmaxf(x0, -125.f), 126.f);
  x1 = fminf(fmaxf(x1, -125.f), 126.f);
  const f32x2 x = pack_f2(x0, x1);
  const f32x2 xf = add_f2(x, pack_f2(12582912.f, 12582912.f));
  const f32x2 n = add_f2(xf, pack_f2(-12582912.f, -12582912.f));
  const f32x2 f = fma_f2(n, pack_f2(-1.f, -1.f), x);
  f32x2 r = fma_f2(pack_f2(0.055171649903059006f, 0.055171649903059006f), f, pack_f2(0.2426111251115799f, 0.2426111251115799f));
  r = fma_f2(r, f, pack_f2(0.6932609677314758f, 0.6932609677314758f));
  r = fma_f2(r, f, pack_f2(0.9999280571937561f, 0.9999280571937561f));
  float r0, r1, n0, n1;
  unpack_f2(r, r0, r1);
  unpack_f2(xf, n0, n1);
  const float y0 = __uint_as_float(__float_as_uint(r0) + (__float_as_uint(n0) << 23));
  const float y1 = __uint_as_float(__float_as_uint(r1) + (__float_as_uint(n1) << 23));
  __nv_bfloat162 v = __floats2bfloat162_rn(y0, y1);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void cp16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// byte offset of 16-byte chunk `c` (0/1) of row `row` inside a swizzled [112][32 B] tile
__device__ __forceinline__ uint32_t swz(int row, int c) { return (uint32_t)(row * ROWB + ((c ^ ((row >> 2) & 1)) << 4)); }

// POLY: how many of the four probability pairs of a 16-key step take the FMA-pipe exp2 (0 = all on MUFU.EX2, the default).  Measured
// (profiles/r2_t_attn_poly_mlp_ab.log, dim 128 x 8 heads @160^2, batch 64): 0.592 / 0.601 / 0.678 / 0.751 ms for 0 / 1 / 2 / 3 pairs
// -- every pair moved off the XU pipe costs ~12 issue slots and the kernel has none to spare (XU 66 % busy, issue-bound), so the
// experiment stays behind FBANET_ATTN_POLY and MUFU.EX2 stays the path.
template <bool PRESCALED, int POLY>
__global__ void __launch_bounds__(WARPS * 32, 1) window_attention_dh16_kernel(const fbanet_attn_params p, const int win_chunk) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, mi = lane >> 3, l7 = lane & 7;
  const int head = blockIdx.x;   // heads fastest: CTAs that share token rows are co-resident (DRAM page / L2 locality)
  const int H = p.H, W = p.W, shift = p.shift;
  const int nwx = W / WIN, nwy = H / WIN, nw_img = nwx * nwy;
  const int total_windows = p.B * nw_img;
  const int w_begin = blockIdx.y * win_chunk, w_end = min(w_begin + win_chunk, total_windows);

  float4* biasF = reinterpret_cast<float4*>(smem);
  const uint32_t smem_u = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t wbase = smem_u + BIAS_BYTES + warp * WARP_BYTES;
  const uint32_t Ks = wbase + 2 * TILE_BYTES, Vs = wbase + 3 * TILE_BYTES;
  uint8_t* wgen = smem + BIAS_BYTES + warp * WARP_BYTES;

  const bf16* qkv = reinterpret_cast<const bf16*>(p.qkv) + head * 16;
  bf16* outp = reinterpret_cast<bf16*>(p.out) + head * 16;
  const int64_t ld = p.qkv_ld, old = p.out_ld;
  const int C = p.C;

  // q | k | v rows of window `wid` -> (q buffer `qb`, k, v); 200 (row, chunk) slots over 32 lanes: lane pair = one 32-byte row.
  // All in-image offsets are 32-bit (an image's qkv rows span < 2^31 elements); the swizzled destination of slot i is a lane
  // constant + 512*i because 16 rows keep bit 2 of the row index.
  const int j0 = lane >> 1, c8 = (lane & 1) * 8;
  const uint32_t d0 = swz(j0, lane & 1);
  const uint32_t ldu = (uint32_t)p.qkv_ld, oldu = (uint32_t)p.out_ld;
  auto stage = [&](int b, int wy, int wx, int qb) {
    const int ty0 = wy * WIN + shift, tx0 = wx * WIN + shift;
    const bf16* ibq = qkv + (int64_t)b * H * W * ld;
    const bf16* ibk = ibq + C;
    const bf16* ibv = ibk + C;
    const uint32_t Qs = wbase + qb * TILE_BYTES + d0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int j = j0 + 16 * i;
      if (i < 6 || j < NTOK) {
        const int jy = (j * 205) >> 11, jx = j - jy * WIN;
        int y = ty0 + jy, x = tx0 + jx;
        y -= (y >= H) ? H : 0;
        x -= (x >= W) ? W : 0;
        const uint32_t off = (uint32_t)(y * W + x) * ldu + c8;
        cp16(Qs + i * 512, ibq + off);
        cp16(Ks + d0 + i * 512, ibk + off);
        cp16(Vs + d0 + i * 512, ibv + off);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // zero the padding rows (100..111) of this warp's buffers once: padded keys must be finite, padded v rows zero
  for (int e = lane; e < 4 * 12 * 2; e += 32) {
    const int buf = e / 24, r = NTOK + (e % 24) / 2, c = e & 1;
    *reinterpret_cast<uint4*>(wgen + buf * TILE_BYTES + r * ROWB + c * 16) = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  // window coordinates (image b, window row wy, window column wx) advance by WARPS windows per item without divisions
  int wid = w_begin + warp;
  int nb = wid / nw_img, nwy_ = (wid - nb * nw_img) / nwx, nwx_ = wid - nb * nw_img - nwy_ * nwx;   // "next" item = first item
  if (wid < w_end) stage(nb, nwy_, nwx_, 0);

  // ---- relative-position bias of this head, fragment-major, log2 units; padded keys -> -1e30 ----
  for (int e = tid; e < MT * NT * 32 * 4; e += WARPS * 32) {
    const int comp = e & 3, ln = (e >> 2) & 31, tile = e >> 7;
    const int mt = tile / NT, nt = tile - mt * NT;
    const int row = mt * 16 + (ln >> 2) + (comp >> 1) * 8, col = nt * 8 + 2 * (ln & 3) + (comp & 1);
    float b = 0.f;
    if (col >= NTOK) b = -1e30f;
    else if (row < NTOK) {
      const int yi = row / WIN, xi = row - yi * WIN, yj = col / WIN, xj = col - yj * WIN;
      b = __ldg(p.bias_table + ((yi - yj + WIN - 1) * (2 * WIN - 1) + (xi - xj + WIN - 1)) * p.heads + head) * LOG2E;
    }
    reinterpret_cast<float*>(smem)[e] = b;
  }
  __syncthreads();

  // per-thread key bitmasks for the Swin shift mask: bit (2*nt+u) <-> key nt*8+2t+u in the lower/right band
  uint32_t keyHy = 0, keyHx = 0;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = nt * 8 + 2 * t + u, jy = (j * 205) >> 11, jx = j - jy * WIN;
      if (jy >= WIN - shift) keyHy |= 1u << (2 * nt + u);
      if (jx >= WIN - shift) keyHx |= 1u << (2 * nt + u);
    }
  const float scale2 = p.scale * LOG2E;
  const uint32_t ones = 0x3F803F80u;   // bf16x2 (1, 1): B fragment of the row-sum MMA

  // lane-constant ldmatrix offsets
  const uint32_t kfo0 = swz((mi >> 1) * 8 + l7, mi & 1);            // + pr*16 rows: (mi>>1)*8 + l7 < 16, swizzle bit = bit 2 of row
  const uint32_t vfo0 = swz((mi & 1) * 8 + l7, mi >> 1);
  const uint32_t qfo0 = swz((mi & 1) * 8 + l7, mi >> 1);

  for (int it = 0; wid < w_end; wid += WARPS, ++it) {
    const int qb = it & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- K and V fragments of the whole window -> registers ----
    uint32_t kf[7][4], vf[7][4];
#pragma unroll
    for (int pr = 0; pr < 7; ++pr) {
      ldsm4(kf[pr], Ks + pr * 16 * ROWB + kfo0);     // 16 rows = 512 B: row bits 2.. unchanged mod the swizzle (16*k keeps bit 2)
      ldsm4t(vf[pr], Vs + pr * 16 * ROWB + vfo0);
    }
    __syncwarp();
    const int b = nb, wy = nwy_, wx = nwx_;
    nwx_ += WARPS;
    while (nwx_ >= nwx) { nwx_ -= nwx; ++nwy_; }
    while (nwy_ >= nwy) { nwy_ -= nwy; ++nb; }
    if (wid + WARPS < w_end) stage(nb, nwy_, nwx_, qb ^ 1);

    const bool wrap = shift > 0 && (wy == nwy - 1 || wx == nwx - 1);
    const uint32_t wyl = (shift > 0 && wy == nwy - 1) ? 0xffffffffu : 0u, wxl = (shift > 0 && wx == nwx - 1) ? 0xffffffffu : 0u;
    const uint32_t Qs = wbase + qb * TILE_BYTES;
    uint8_t* Qg = wgen + qb * TILE_BYTES;

    // ---- the 7 query tiles ----
    // Softmax is shift invariant, and both fp32 and bf16 carry 8 exponent bits, so exp2(s - c) keeps full relative precision
    // for ANY shift c within ~2^+-60 of the row maximum.  The fast path therefore uses c = 0: no row-max pass, no shuffles, no
    // subtract per score, and -- because nothing depends on the whole row any more -- the keys are STREAMED: 16 keys at a
    // time go bias -> score MMA -> exp -> pack -> P v MMA, two steps in flight, ~16 live score registers instead of 52.
    // The row sum doubles as the safety check: unless 2^-60 < l < 2^60 for every row of the tile (NaN fails too) the tile is
    // redone with the exact row maximum folded into the accumulator init.
    uint32_t qa[4];
    float o0[4], o1[4], os[4];
    const uint32_t wyl_ = wyl, wxl_ = wxl;
    auto tile_masks = [&](const int mt, uint32_t& m0, uint32_t& m1) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      const int r0y = (r0 * 205) >> 11, r0x = r0 - r0y * WIN, r1y = (r1 * 205) >> 11, r1x = r1 - r1y * WIN;
      m0 = (wyl_ & (r0y >= WIN - shift ? ~keyHy : keyHy)) | (wxl_ & (r0x >= WIN - shift ? ~keyHx : keyHx));
      m1 = (wyl_ & (r1y >= WIN - shift ? ~keyHy : keyHy)) | (wxl_ & (r1x >= WIN - shift ? ~keyHx : keyHx));
    };
    // scores of key tile nt for query tile mt, in log2 units, minus the per-row shift (c0, c1)
    auto score_tile = [&](float (&sc)[4], const int mt, const int nt, const uint32_t m0, const uint32_t m1, const float c0, const float c1,
                          auto shifted_tag, auto wrap_tag) {
      constexpr bool SHIFTED = decltype(shifted_tag)::value;
      constexpr bool WRAP = decltype(wrap_tag)::value;   // compile-time copy of `wrap`: no branch per key tile
      constexpr float M = 100.f * LOG2E;
      const float4 bv = biasF[(mt * NT + nt) * 32 + lane];
      if (PRESCALED) {
        sc[0] = bv.x; sc[1] = bv.y; sc[2] = bv.z; sc[3] = bv.w;
        if (WRAP) {   // Swin shift mask (-100 in natural units) folded into the accumulator init
          if (m0 & (1u << (2 * nt))) sc[0] -= M;
          if (m0 & (2u << (2 * nt))) sc[1] -= M;
          if (m1 & (1u << (2 * nt))) sc[2] -= M;
          if (m1 & (2u << (2 * nt))) sc[3] -= M;
        }
        if (SHIFTED) { sc[0] -= c0; sc[1] -= c0; sc[2] -= c1; sc[3] -= c1; }
        mma16816(sc, qa, kf[nt >> 1][(nt & 1) * 2], kf[nt >> 1][(nt & 1) * 2 + 1]);
      } else {
        sc[0] = sc[1] = sc[2] = sc[3] = 0.f;
        mma16816(sc, qa, kf[nt >> 1][(nt & 1) * 2], kf[nt >> 1][(nt & 1) * 2 + 1]);
        sc[0] = fmaf(sc[0], scale2, bv.x); sc[1] = fmaf(sc[1], scale2, bv.y);
        sc[2] = fmaf(sc[2], scale2, bv.z); sc[3] = fmaf(sc[3], scale2, bv.w);
        if (WRAP) {
          if (m0 & (1u << (2 * nt))) sc[0] -= M;
          if (m0 & (2u << (2 * nt))) sc[1] -= M;
          if (m1 & (1u << (2 * nt))) sc[2] -= M;
          if (m1 & (2u << (2 * nt))) sc[3] -= M;
        }
        if (SHIFTED) { sc[0] -= c0; sc[1] -= c0; sc[2] -= c1; sc[3] -= c1; }
      }
    };
    // one query tile, keys streamed in 7 steps of 16: O = P v and l = P 1 with p = exp2(s - c)
    // LASTQ (query tile 6 = rows 96..111): the tile's upper half (rows 104..111) is padding for every lane, so its probabilities
    // -- half of the tile's MUFU.EX2, 1/14 of the kernel's -- are not computed (P = 0 there; those output rows are never stored).
    auto stream_tile = [&](const int mt, const uint32_t m0, const uint32_t m1, const float c0, const float c1, auto shifted_tag, auto wrap_tag,
                           auto last_tag) {
      constexpr bool LASTQ = decltype(last_tag)::value;
      float sa[2][4], sb[2][4];   // two steps in flight: [step parity][..] for key tiles 2ks (sa) and 2ks+1 (sb)
#pragma unroll
      for (int e = 0; e < 4; ++e) o0[e] = o1[e] = os[e] = 0.f;
      score_tile(sa[0], mt, 0, m0, m1, c0, c1, shifted_tag, wrap_tag);
      score_tile(sb[0], mt, 1, m0, m1, c0, c1, shifted_tag, wrap_tag);
#pragma unroll
      for (int ks = 0; ks < 7; ++ks) {
        const int cur = ks & 1, nxt = cur ^ 1;
        if (2 * ks + 2 < NTU) score_tile(sa[nxt], mt, 2 * ks + 2, m0, m1, c0, c1, shifted_tag, wrap_tag);
        if (2 * ks + 3 < NTU) score_tile(sb[nxt], mt, 2 * ks + 3, m0, m1, c0, c1, shifted_tag, wrap_tag);
        uint32_t pa[4];
        pa[0] = POLY >= 1 ? ex2_poly_pack2(sa[cur][0], sa[cur][1]) : pack2(ex2(sa[cur][0]), ex2(sa[cur][1]));
        pa[1] = LASTQ ? 0u : (POLY >= 3 ? ex2_poly_pack2(sa[cur][2], sa[cur][3]) : pack2(ex2(sa[cur][2]), ex2(sa[cur][3])));
        if (2 * ks + 1 < NTU) {
          pa[2] = pack2(ex2(sb[cur][0]), ex2(sb[cur][1]));
          pa[3] = LASTQ ? 0u : (POLY >= 2 ? ex2_poly_pack2(sb[cur][2], sb[cur][3]) : pack2(ex2(sb[cur][2]), ex2(sb[cur][3])));
        } else {
          pa[2] = pa[3] = 0u;   // keys 104..111 are padding
        }
        mma16816(o0, pa, vf[ks][0], vf[ks][1]);
        mma16816(o1, pa, vf[ks][2], vf[ks][3]);
        mma16816(os, pa, ones, ones);   // every column = sum of the bf16 probabilities P v used
      }
    };
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      uint32_t m0 = 0, m1 = 0;
      if (wrap) tile_masks(mt, m0, m1);
      ldsm4(qa, Qs + mt * 16 * ROWB + qfo0);
      if (mt == MT - 1) {
        if (wrap) stream_tile(mt, m0, m1, 0.f, 0.f, std::false_type{}, std::true_type{}, std::true_type{});
        else stream_tile(mt, 0u, 0u, 0.f, 0.f, std::false_type{}, std::false_type{}, std::true_type{});
      } else if (wrap) stream_tile(mt, m0, m1, 0.f, 0.f, std::false_type{}, std::true_type{}, std::false_type{});
      else stream_tile(mt, 0u, 0u, 0.f, 0.f, std::false_type{}, std::false_type{}, std::false_type{});
      const bool ok = ((os[0] > 8.6e-19f && os[0] < 1.15e18f) || r0 >= NTOK) && ((os[2] > 8.6e-19f && os[2] < 1.15e18f) || r1 >= NTOK);
      if (!__all_sync(0xffffffffu, ok)) {   // out-of-band row sum (huge logits): exact row maximum, then redo.  Cold.
        float mx0 = -3.0e38f, mx1 = -3.0e38f;
#pragma unroll 1
        for (int nt = 0; nt < NTU; ++nt) {
          float sc[4];
          const float4 bv = biasF[(mt * NT + nt) * 32 + lane];
          constexpr float M = 100.f * LOG2E;
          sc[0] = sc[1] = sc[2] = sc[3] = 0.f;
          const int pr = nt >> 1;
          uint32_t kb0 = 0, kb1 = 0;
#pragma unroll
          for (int q = 0; q < 7; ++q) if (q == pr) { kb0 = (nt & 1) ? kf[q][2] : kf[q][0]; kb1 = (nt & 1) ? kf[q][3] : kf[q][1]; }
          mma16816(sc, qa, kb0, kb1);
          const float sq = PRESCALED ? 1.f : scale2;
          sc[0] = fmaf(sc[0], sq, bv.x) - (((m0 >> (2 * nt)) & 1u) ? M : 0.f);
          sc[1] = fmaf(sc[1], sq, bv.y) - (((m0 >> (2 * nt)) & 2u) ? M : 0.f);
          sc[2] = fmaf(sc[2], sq, bv.z) - (((m1 >> (2 * nt)) & 1u) ? M : 0.f);
          sc[3] = fmaf(sc[3], sq, bv.w) - (((m1 >> (2 * nt)) & 2u) ? M : 0.f);
          mx0 = fmaxf(mx0, fmaxf(sc[0], sc[1]));
          mx1 = fmaxf(mx1, fmaxf(sc[2], sc[3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        stream_tile(mt, m0, m1, mx0, mx1, std::true_type{}, std::true_type{}, std::false_type{});   // m0 = m1 = 0 when the window does not wrap
      }
      float i0, i1;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i0) : "f"(os[0]));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(os[2]));
      // O tile over the consumed q rows (same swizzle): (row, ch 2t..2t+1) and (row, ch 8+2t..)
      if (r0 < NTOK) {
        *reinterpret_cast<uint32_t*>(Qg + swz(r0, 0) + 4 * t) = pack2(o0[0] * i0, o0[1] * i0);
        *reinterpret_cast<uint32_t*>(Qg + swz(r0, 1) + 4 * t) = pack2(o1[0] * i0, o1[1] * i0);
      }
      if (r1 < NTOK) {
        *reinterpret_cast<uint32_t*>(Qg + swz(r1, 0) + 4 * t) = pack2(o0[2] * i1, o0[3] * i1);
        *reinterpret_cast<uint32_t*>(Qg + swz(r1, 1) + 4 * t) = pack2(o1[2] * i1, o1[3] * i1);
      }
    }
    __syncwarp();
    // ---- stream the 100 x 32 B output rows ----
    {
      const int ty0 = wy * WIN + shift, tx0 = wx * WIN + shift;
      bf16* obase = outp + (int64_t)b * H * W * old;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const int j = j0 + 16 * i;
        if (i < 6 || j < NTOK) {
          const int jy = (j * 205) >> 11, jx = j - jy * WIN;
          int y = ty0 + jy, x = tx0 + jx;
          y -= (y >= H) ? H : 0;
          x -= (x >= W) ? W : 0;
          const uint4 v = *reinterpret_cast<const uint4*>(Qg + d0 + i * 512);
          *reinterpret_cast<uint4*>(obase + ((uint32_t)(y * W + x) * oldu + c8)) = v;
        }
      }
    }
    __syncwarp();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace

int window_attention_dh16_supported(const fbanet_attn_params* p) {
  if (p->dtype != FBANET_BF16 || p->win != WIN || p->C != p->heads * 16) return 0;
  if ((p->qkv_ld % 8) || (p->out_ld % 8) || ((uintptr_t)p->qkv % 16) || ((uintptr_t)p->out % 16)) return 0;
  return 1;
}

int window_attention_dh16_launch(const fbanet_attn_params* p, cudaStream_t s) {
  constexpr size_t smem = BIAS_BYTES + (size_t)WARPS * WARP_BYTES;
  static int n_sm = 0;
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaSuccess;
    for (auto fn : {window_attention_dh16_kernel<true, 0>, window_attention_dh16_kernel<true, 1>, window_attention_dh16_kernel<false, 0>})
      if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0;
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
    opted = true;
  }
  const int total_windows = p->B * (p->H / WIN) * (p->W / WIN);
  // one CTA per SM (smem-bound); every CTA serves one head: heads x floor(SMs / heads) CTAs
  int nchunks = n_sm / p->heads;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > total_windows) nchunks = total_windows;
  const int chunk = (total_windows + nchunks - 1) / nchunks;
  dim3 grid(p->heads, (total_windows + chunk - 1) / chunk);
  // FBANET_ATTN_POLY=1: one probability pair per step on the FMA-pipe exp2 (experiment, slower: see the kernel's comment)
  static const int poly = [] { const char* e = getenv("FBANET_ATTN_POLY"); return e ? atoi(e) : 0; }();
  if (p->q_prescaled) {
    if (poly <= 0) window_attention_dh16_kernel<true, 0><<<grid, WARPS * 32, smem, s>>>(*p, chunk);
    else window_attention_dh16_kernel<true, 1><<<grid, WARPS * 32, smem, s>>>(*p, chunk);
  } else window_attention_dh16_kernel<false, 0><<<grid, WARPS * 32, smem, s>>>(*p, chunk);
  return check_launch();
}

}  // namespace fbanet
