// K3/K4/K5/K9 on the 5th-gen tensor cores: implicit-GEMM convolution / linear layer, bf16 x bf16 -> fp32.
//
//   out[pixel, co] = act( sum_{step} A_step[pixel, 0:64] . B_step[co, 0:64] + bias[co] ) + residual
//
// One K-step = one (tap, source, 64-channel chunk).  The A tile of a step is a TMA box
// {64 channels, tw, th, 1 image} of the channels-last source shifted by the tap offset (dx, dy): TMA
// zero-fills the out-of-image part, so there is no im2col buffer and no halo code, and the box lands in
// shared memory already in the canonical K-major SWIZZLE_128B layout tcgen05.mma reads.  The B tile is a
// {64, BN} box of the packed weights [Cout][K].  Accumulators live in TMEM (2 x BN fp32 columns, double
// buffered) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (256 threads, persistent, 1 CTA/SM): warp 0 = TMA producer (one lane), warp 1 = MMA issuer
// (one lane), warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> registers -> bias/act/residual ->
// bf16 -> 16-byte global stores, each thread owns one pixel row).
#include <cuda.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fbanet {

constexpr int TC_MAX_STEPS = 160;
constexpr int TC_MAX_SRC = 4;
constexpr int TC_BK = 64;            // channels per K-step (128 bytes of bf16 = one swizzle row)
constexpr int TC_A_BYTES = 128 * TC_BK * 2;
constexpr int TC_MAX_CHUNKS = 32;
constexpr int TC_HALO_TW = 8, TC_HALO_TH = 16;                       // 3x3 halo mode tile: 8 wide x 16 tall
constexpr int TC_HALO_COPY = (TC_HALO_TH + 2) * TC_HALO_TW * 128;    // one dx-shifted copy: 18 rows x 1 KB
constexpr int TC_HALO_SLOT = 3 * TC_HALO_COPY;                       // 54 KB per 64-channel chunk
// wide-box halo mode (p.halo >= 2): ONE box {64 ch, 16 px, 18 rows} at x0-4 per chunk.  An image row of the box is 2 KB = two
// swizzle atoms, so tap (ky,kx) is the same buffer read from byte offset ky*2048 + (3+kx)*128 with a 2 KB stride between the
// 8-row atoms.  TMA and tcgen05.mma both derive the 128B-swizzle XOR from shared-memory ADDRESS bits [7:9], so a start
// address that is only 128-byte aligned needs nothing else (descriptor base offset 0; mode 2, which sets it to the row
// phase, is kept only as the experiment that proved it wrong).  36 KB instead of 54 KB per chunk through TMA / L2.
constexpr int TC_HALO_WROW = 16 * 128;
constexpr int TC_HALO_WSLOT = (TC_HALO_TH + 2) * TC_HALO_WROW;

// generic mode: one A box + one B slab per K-step
struct KStep { int16_t src, c0, dx, dy; };
// halo mode (3x3 stride 1): one 64-channel chunk = 3 dx-shifted halo boxes, re-used by the 9 taps
struct Chunk { int16_t src, c0; int32_t cg; };   // cg = offset of the chunk in the concatenated channel axis

struct TcParams {
  CUtensorMap amap[TC_MAX_SRC];
  CUtensorMap bmap;
  CUtensorMap omap[4];    // TMA-store epilogue: output view(s); 4 = the (i,j) sub-pixel planes of the 2x2 scatter store
  CUtensorMap rmap;       // residual view (TMA-loaded into the staging tile)
  KStep steps[TC_MAX_STEPS];
  Chunk chunks[TC_MAX_CHUNKS];
  const float* bias;
  const float* alpha;
  const float* ln_stats;  // folded LayerNorm: per-row (mean, rstd), see fbanet_conv_params
  const float* ln_gamma;  // LayerNorm applied to the A tile in shared memory (LN_SMEM instantiation): fp32 [K]
  const float* ln_beta;
  float ln_eps;
  const bf16* residual;
  void* out;
  const float* base;
  int64_t res_img_stride, out_img_stride, base_img_stride;
  int nsteps;             // K-steps per tile (MMA groups of 4)
  int halo, nchunks, ctot;
  int N, Ho, Wo;          // tile space == output pixels
  int tw, th, tiles_x, tiles_y, m_tiles;
  int BN, n_tiles_n, Cout, Cout_store;
  int n_inner;            // N tiles one CTA computes per A tile (A-stationary; 1 = each CTA owns one N tile).  > 1: n_tiles_n is 1, all
                          // n_inner * nsteps weight slabs are resident, an A tile's slots are released after its last N tile
  int act, store_mode, res_ld, out_ld;
  int a_slots, a_slot_bytes, a_box_bytes;
  int a_mul;              // box origin = tile origin * a_mul + tap offset: 2 for the stride-2 conv read through element-strided TMA boxes, else 1
  int b_slots, b_resident;
  int tma_store;          // 1: each epilogue warp stages 32x64 bf16 sub-tiles in smem and stores them with TMA
  int stage_bufs;         // staging buffers per epilogue warp (2, or 1 when shared memory is tight)
  int store_f16;          // staged epilogue: store IEEE fp16 instead of bf16
  int gelu_h2;            // store_f16 + tanh-GELU: the activation runs on packed half2 (bias_s holds bias / 2; see the staged epilogue)
  int fold;               // tapsum mode: outputs are hi + lo halves (see fbanet_conv_params.fold_hi_lo)
  int tapsum;             // > 0: tap-stacked 3x3 conv with `tapsum` outputs per tap (see the TAPSUM epilogue); tiles step by (tw-2, th-2)
  int step_x, step_y, org;  // tile origin = tile index * step + org (tw, th, 0 except in tapsum mode: tw-2, th-2, -1)
  int nacc, nacc_shift;   // accumulator stages in TMEM (4 when 4 x BN <= 512 columns, else 2) and log2 of it
  int debug;              // FBANET_TC_DEBUG timing experiments (results are garbage): 1 = epilogue only hands the accumulator back,
                          // 2 = no tcgen05.mma issued (commits only), 4 = A producer arrives without loading
};

template <int NV>
__device__ __forceinline__ void apply_act_vec(float (&f)[NV], const int act, const float alpha) {
  if (act == FBANET_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) f[j] = fmaxf(f[j], 0.f);
  } else if (act == FBANET_ACT_PRELU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) f[j] = f[j] > 0.f ? f[j] : alpha * f[j];
  } else if (act == FBANET_ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < NV; ++j) f[j] = gelu_tanh_fast(f[j]);
  } else if (act == FBANET_ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < NV; ++j) f[j] = gelu_erf(f[j]);
  }
}

// bias + activation + residual + store of 32 (or 16) accumulator columns of one pixel row
// XTRA: the instantiations that carry the half2 GELU (see the kernel's template parameters)
template <bool XTRA>
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, const uint32_t (&v)[32], const int nc, const int col0, const int img,
                                               const int y, const int x, const float alpha, const float* bias_s,
                                               const uint4* rpre = nullptr, const bool has_ln = false, const float ln_rstd = 1.f) {
  if (XTRA && p.gelu_h2) {
    // fp16 store + tanh-GELU on packed half2: the staged epilogue's arithmetic, operation for operation (bias_s holds bias / 2), so a GEMM
    // gives the same bits whichever epilogue its shared-memory plan selects (STORE_NHWC, no residual, no folded LayerNorm: launcher)
    const uint32_t A2 = 0x3E623E62u, B2 = 0x34913491u;             // half2 (2 k0, 2 k0) and (8 k0 k1, 8 k0 k1)
    bf16* op = reinterpret_cast<bf16*>(p.out) + img * p.out_img_stride + ((int64_t)y * p.Wo + x) * p.out_ld + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (j < nc && col0 + j < p.Cout_store) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + j), b1 = *reinterpret_cast<const float4*>(bias_s + j + 4);
        const f32x2 bb[4] = {pack_f2(b0.x, b0.y), pack_f2(b0.z, b0.w), pack_f2(b1.x, b1.y), pack_f2(b1.z, b1.w)};
        uint4 h;
        uint32_t* hp = reinterpret_cast<uint32_t*>(&h);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          hp[e] = gelu_half_h2(f2_to_f16x2(fma_f2(pack_f2(__uint_as_float(v[j + 2 * e]), __uint_as_float(v[j + 2 * e + 1])), pack_f2(0.5f, 0.5f), bb[e])), A2, B2);
        *reinterpret_cast<uint4*>(op + j) = h;
      }
    }
    return;
  }
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = (j < nc) ? __uint_as_float(v[j]) : 0.f;
  if (has_ln) {   // folded LayerNorm (row-centred weights): rstd * acc + bias
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < nc) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j);
        f[j] = fmaf(f[j], ln_rstd, b4.x); f[j + 1] = fmaf(f[j + 1], ln_rstd, b4.y);
        f[j + 2] = fmaf(f[j + 2], ln_rstd, b4.z); f[j + 3] = fmaf(f[j + 3], ln_rstd, b4.w);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {   // bias of this CTA's N-tile sits in shared memory (broadcast reads)
      if (j < nc) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j);
        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
      }
    }
  }
  if (p.act == FBANET_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  } else if (p.act == FBANET_ACT_PRELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : alpha * f[j];
  } else if (p.act == FBANET_ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_tanh_fast(f[j]);
  } else if (p.act == FBANET_ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
  }
  if (p.store_mode == FBANET_STORE_NCHW_BASE) {
    // final conv: fp32 planar out + bilinear x4 (align_corners=False) of the low-res base frame
    const int Hb = p.Ho >> 2, Wb = p.Wo >> 2;
    float sy = 0.25f * (y + 0.5f) - 0.5f, sx = 0.25f * (x + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int yb = (int)sy, xb = (int)sx;
    const int y1 = yb + (yb < Hb - 1 ? 1 : 0), x1 = xb + (xb < Wb - 1 ? 1 : 0);
    const float wy = sy - yb, wx = sx - xb, hy = 1.f - wy, hx = 1.f - wx;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int col = col0 + j;
      if (col < p.Cout_store) {
        const float* bp = p.base + img * p.base_img_stride + (int64_t)col * Hb * Wb;
        const float bl = hy * (hx * __ldg(bp + yb * Wb + xb) + wx * __ldg(bp + yb * Wb + x1)) +
                         wy * (hx * __ldg(bp + y1 * Wb + xb) + wx * __ldg(bp + y1 * Wb + x1));
        reinterpret_cast<float*>(p.out)[img * p.out_img_stride + ((int64_t)col * p.Ho + y) * p.Wo + x] = f[j] + bl;
      }
    }
    return;
  }
  if (p.store_mode == FBANET_STORE_NHWC_F32) {   // narrow fp32 score map
    float* op = reinterpret_cast<float*>(p.out) + img * p.out_img_stride + ((int64_t)y * p.Wo + x) * p.out_ld + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nc && col0 + j < p.Cout_store) op[j] = f[j];
    return;
  }
  int64_t opix;
  int ocol;
  if (p.store_mode == FBANET_STORE_NHWC) {
    opix = (int64_t)y * p.Wo + x;
    ocol = col0;
    if (p.residual) {
      const bf16* rp = p.residual + img * p.res_img_stride + opix * p.res_ld + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j < nc) {
          const uint4 rv = rpre ? rpre[j >> 3] : *reinterpret_cast<const uint4*>(rp + j);   // rpre: fetched before the MMA wait
          const uint32_t ru[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { f[j + 2 * e] += __uint_as_float(ru[e] << 16); f[j + 2 * e + 1] += __uint_as_float(ru[e] & 0xffff0000u); }
        }
      }
    }
  } else {  // FBANET_STORE_CONVT2: col = (2i+j)*Co + co
    const int Co = p.Cout >> 2;
    const int qq = col0 / Co;
    ocol = col0 - qq * Co;
    opix = (int64_t)(2 * y + (qq >> 1)) * (2 * p.Wo) + (2 * x + (qq & 1));
  }
  bf16* op = reinterpret_cast<bf16*>(p.out) + img * p.out_img_stride + opix * p.out_ld + ocol;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    if (j < nc && col0 + j < p.Cout_store) {
      float t[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = f[j + e];
      if (p.store_f16) {   // IEEE fp16 instead of bf16 (same 2-byte elements)
        uint4 v;
        v.x = f2_to_f16x2(pack_f2(t[0], t[1])); v.y = f2_to_f16x2(pack_f2(t[2], t[3]));
        v.z = f2_to_f16x2(pack_f2(t[4], t[5])); v.w = f2_to_f16x2(pack_f2(t[6], t[7]));
        *reinterpret_cast<uint4*>(op + j) = v;
      } else
      store_vec<bf16, 8>(op + j, t);
    }
  }
}

// The MMA issuer's loop for resident weights, one instantiation per A layout.
// HALO: 0 = plain A tile (one K-step per slot), 1 = three dx-shifted halo copies (tap (ky,kx) = copy kx, ky rows of 1 KB down),
// 3 = one wide halo box (tap starts at byte ky*2048 + (3+kx)*128, 2 KB between 8-row atoms, descriptor base offset 0).
// Everything the loop needs is passed in registers, barriers as raw shared addresses, and the taps are a ROLLED loop whose
// operand addresses advance by constants: measured on the 64->64 body conv and the N = 16 final conv (ncu source view,
// tools/prof_bound.py), the issuing warp spent ~500 cycles per tile outside the tcgen05.mma instructions -- first in LDCU
// (kernel parameter) -> branch chains and generic->shared conversions, then, with the 36 MMAs unrolled, in ~150 uniform-register
// instructions (with spills) that ptxas hoisted in front of the first MMA to precompute all 72 descriptors.  That is longer than
// the work the tcgen05 queue holds, so the tensor pipe ran dry once per tile (43 % active on the body conv where
// tools/ubench/umma.cu gives 57-64 % as the N = 64 ceiling).
template <int TAPS, int HALO, bool NI = false>
__device__ __forceinline__ void mma_loop_resident(const int mt0, const int mt_step, const int m_tiles, const int units, const int a_slots,
                                                  const uint32_t a_slot16, const uint32_t sa16_0, const uint32_t sb16_0, const uint32_t bstep16,
                                                  const uint32_t tmem_base, const uint32_t BN, const uint32_t idesc, const uint32_t bar_af,
                                                  const uint32_t bar_ae, const uint32_t bar_tf, const uint32_t bar_te, const bool nomma, const uint32_t nacc_mask,
                                                  const uint32_t nacc_shift, const int n_inner_arg = 1) {
  const int n_inner = NI ? n_inner_arg : 1;          // (compile-time 1 without NI: the N-tile walk folds away)
  const uint64_t desc_b = make_sw128_desc(0);        // descriptor with a zero start address
  const uint64_t desc_a = HALO >= 2 ? ((desc_b & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)(TC_HALO_WROW >> 4) << 32)) : desc_b;   // wide: 2 KB between atoms
  constexpr uint32_t A_START = HALO >= 2 ? (3u * 128u) >> 4 : 0u;
  constexpr uint32_t A_KX = HALO == 1 ? (uint32_t)TC_HALO_COPY >> 4 : (HALO >= 2 ? 128u >> 4 : 0u);                 // next tap in the row
  int aslot = 0;
  uint32_t aphase = 0, sa16 = sa16_0;
  int it = 0;
  for (int mt = mt0; mt < m_tiles; mt += mt_step)
  for (int ni = 0; ni < n_inner; ++ni, ++it) {       // A-stationary: the tile's A slots serve n_inner N tiles (weight slabs ni * units ..)
    const uint32_t acc = (uint32_t)it & nacc_mask;
    mbar_wait_a(bar_te + acc * 8u, (((uint32_t)it >> nacc_shift) & 1u) ^ 1u);     // epilogue has drained this accumulator
    const uint32_t tmem_d = tmem_base + acc * BN;
    uint32_t b16 = sb16_0 + (uint32_t)(ni * units) * bstep16;
    const int aslot0 = aslot;
    const uint32_t aphase0 = aphase, sa16_t = sa16;
    const bool last_n = ni == n_inner - 1;
    for (int u = 0; u < units; ++u) {
      mbar_wait_a(bar_af + (uint32_t)aslot * 8u, aphase);
      tc_fence_after();
      uint32_t at = sa16 + A_START;
      int kx = 0;
#pragma unroll 1
      for (int t = 0; t < TAPS; ++t) {
        if (elect_one() && !nomma) {
          const uint32_t first = (u == 0 && t == 0) ? 0u : 1u;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16(tmem_d, desc_a + (uint64_t)(at + 2 * k), desc_b + (uint64_t)(b16 + 2 * k), idesc, k == 0 ? first : 1u);
        }
        __syncwarp();
        b16 += bstep16;
        if (TAPS > 1) {   // next tap: one step along the row, or back to its start and one image row down
          if (++kx < 3) at += A_KX;
          else {
            kx = 0;
            if (HALO == 1) at += (uint32_t)(1024 >> 4) - 2u * ((uint32_t)TC_HALO_COPY >> 4);
            else at += ((uint32_t)TC_HALO_WROW >> 4) - 2u * (128u >> 4);
          }
        }
      }
      if (elect_one()) {
        if (last_n) umma_commit_a(bar_ae + (uint32_t)aslot * 8u);         // frees the A slot when these MMAs retire
        if (u == units - 1) umma_commit_a(bar_tf + acc * 8u);             // accumulator ready for the epilogue
      }
      __syncwarp();
      sa16 += a_slot16;
      if (++aslot == a_slots) { aslot = 0; aphase ^= 1u; sa16 = sa16_0; }
    }
    if (!last_n) { aslot = aslot0; aphase = aphase0; sa16 = sa16_t; }    // next N tile: the same A slots (their full phase is still the completed one)
  }
}

// LayerNorm of the A tile IN SHARED MEMORY (1x1 GEMMs whose input is a LayerNorm output: norm1 -> qkv, norm2 -> fc1,
// layers/fba_net.py:196,246).  The raw token rows arrive by TMA exactly as for a plain GEMM; four extra warps -- one thread per
// tile row -- normalise them in place before the MMA warp may read them (a_norm barriers stand between a_full and the issuer), so
// the separate LayerNorm pass and the normalised tensor's trip through HBM disappear.  The arithmetic is the LayerNorm kernel's
// (layernorm_bf16_kernel), operation for operation: per 8-element chunk a serial sum, the chunks combined by the same halving
// tree its xor-shuffles form, mean = s / C, the centred second pass, rstd = rsqrt(q / C + eps), (x - mean) * rstd * gamma + beta
// rounded to bf16 -- the GEMM sees bit-identical operands.
template <int NK>
__device__ __forceinline__ void ln_smem_loop(const TcParams& p, const uint32_t sa0, uint64_t* a_full, uint64_t* a_norm, const float* ln_gb,
                                             const int mt0, const int mt_step, const int r, const int lane) {
  constexpr int C = 64 * NK, TPT = C / 8;
  const uint32_t roff = (uint32_t)r * 128u, r7 = (uint32_t)r & 7u;
  const uint32_t a_slot_bytes = (uint32_t)p.a_slot_bytes;
  const int a_slots = p.a_slots;
  const float eps = p.ln_eps;
  int aslot = 0;
  uint32_t aphase = 0;
  auto lds8 = [](uint32_t addr, float (&f)[8]) {
    uint32_t u0, u1, u2, u3;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "r"(addr));
    f[0] = __uint_as_float(u0 << 16); f[1] = __uint_as_float(u0 & 0xffff0000u);
    f[2] = __uint_as_float(u1 << 16); f[3] = __uint_as_float(u1 & 0xffff0000u);
    f[4] = __uint_as_float(u2 << 16); f[5] = __uint_as_float(u2 & 0xffff0000u);
    f[6] = __uint_as_float(u3 << 16); f[7] = __uint_as_float(u3 & 0xffff0000u);
  };
  // The K-chunk loops are ROLLED (the tile's slots are walked with a running ring index): unrolled over NK = 4 the three passes held
  // 128 loaded registers at once and spilled.  The chunk sums are combined exactly as the LayerNorm kernel's shuffle tree does --
  // cs[i] += cs[i + o] for o = TPT/2 .. 1 with cs index = 8 k + c: first across the K chunks (k with k + NK/2, then 0 with 1), then
  // across the eight 16-byte chunks c -- so sums A (k even, or NK < 4) and B (k odd, NK = 4) per c suffice.
  auto reduce = [&](float (&A)[8], float (&Bv)[8]) -> float {
    if (NK == 4) {
#pragma unroll
      for (int c = 0; c < 8; ++c) A[c] += Bv[c];
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < o; ++i) A[i] += A[i + o];
    return A[0];
  };
  for (int mt = mt0; mt < p.m_tiles; mt += mt_step) {
    const int slot0 = aslot;
#pragma unroll 1
    for (int k = 0; k < NK; ++k) {
      mbar_wait(&a_full[aslot], aphase);
      if (++aslot == a_slots) { aslot = 0; aphase ^= 1u; }
    }
    float A[8], Bv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) A[c] = Bv[c] = 0.f;
    int sl = slot0;
#pragma unroll 1
    for (int k = 0; k < NK; ++k) {
      const uint32_t base = sa0 + (uint32_t)sl * a_slot_bytes + roff;
      if (++sl == a_slots) sl = 0;
      const bool toB = NK == 4 && (k & 1), first = NK == 4 ? k < 2 : k == 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float f[8];
        lds8(base + (((uint32_t)c ^ r7) << 4), f);
        float s_ = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s_ += f[i];
        if (toB) Bv[c] = first ? s_ : Bv[c] + s_;
        else A[c] = first ? s_ : A[c] + s_;
      }
    }
    const float mean = reduce(A, Bv) * (1.0f / C);
    sl = slot0;
#pragma unroll 1
    for (int k = 0; k < NK; ++k) {
      const uint32_t base = sa0 + (uint32_t)sl * a_slot_bytes + roff;
      if (++sl == a_slots) sl = 0;
      const bool toB = NK == 4 && (k & 1), first = NK == 4 ? k < 2 : k == 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float f[8];
        lds8(base + (((uint32_t)c ^ r7) << 4), f);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = f[i] - mean; q += d * d; }
        if (toB) Bv[c] = first ? q : Bv[c] + q;
        else A[c] = first ? q : A[c] + q;
      }
    }
    const float rstd = rsqrtf(reduce(A, Bv) * (1.0f / C) + eps);
    sl = slot0;
#pragma unroll 1
    for (int k = 0; k < NK; ++k) {
      const int slk = sl;
      const uint32_t base = sa0 + (uint32_t)sl * a_slot_bytes + roff;
      if (++sl == a_slots) sl = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t addr = base + (((uint32_t)c ^ r7) << 4);
        float f[8];
        lds8(addr, f);
        const float* g = ln_gb + k * 64 + c * 8;           // broadcast reads: all lanes take the same columns
        const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(g + 256), b1 = *reinterpret_cast<const float4*>(g + 260);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t o4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float y0 = (f[2 * i] - mean) * rstd * gg[2 * i] + bb[2 * i];
          const float y1 = (f[2 * i + 1] - mean) * rstd * gg[2 * i + 1] + bb[2 * i + 1];
          __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
          o4[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o4[0]), "r"(o4[1]), "r"(o4[2]), "r"(o4[3]) : "memory");
      }
      fence_proxy_async();                                 // generic-proxy writes -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_norm[slk]);            // this warp's 32 rows of K chunk k are normalised
    }
  }
}

// (image, tile row, tile column) of the tiles a CTA owns, advanced incrementally: the role loops used two 32-bit divisions per
// tile, and with 4-MMA tiles (1x1 GEMMs with K = 64, the tap-stacked convs) every serial instruction of a role loop shows
struct TileIter {
  int img, ty, tx, d_img, d_ty, d_tx, tiles_x, tiles_y;
  __device__ __forceinline__ TileIter(int mt0, int mt_step, int tiles_x_, int tiles_y_) : tiles_x(tiles_x_), tiles_y(tiles_y_) {
    const int per_img = tiles_x * tiles_y;
    img = mt0 / per_img;
    int r = mt0 - img * per_img;
    ty = r / tiles_x; tx = r - ty * tiles_x;
    d_img = mt_step / per_img;
    r = mt_step - d_img * per_img;
    d_ty = r / tiles_x; d_tx = r - d_ty * tiles_x;
  }
  __device__ __forceinline__ void next() {
    tx += d_tx; ty += d_ty; img += d_img;
    if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    if (ty >= tiles_y) { ty -= tiles_y; ++img; }
  }
};

constexpr int TC_MAX_A_SLOTS = 8;
// warps: 0 A-producer, 1 MMA, 2 TMEM alloc, 3 B-producer, 4.. epilogue (TC_EPI_SLOTS warps per TMEM lane quarter).
// Measured in one box: 3 slots (512 threads, 128 registers) speed the staged TMA-store epilogue up (fc1 128->512 0.675 -> 0.568 ms,
// qkv 0.416 -> 0.366 ms: its GELU / pack chains are latency bound and want more warps), but cost the direct-store path (3x3 convs,
// K = 896 fusion GEMM) 10-20 % through the lower register cap -- so the slot count is a template parameter chosen per launch.
constexpr int TC_MAX_EPI_WARPS = 12;

// HAS_LN: folded-LayerNorm epilogue (a separate instantiation: compiled into the common one it costs every GEMM 12 %)
// LN_SMEM: four more warps normalise every A tile in shared memory before the MMAs read it (ln_smem_loop)
// XTRA: carries the features only a few launches use -- the half2 GELU of the fp16-store epilogue (gelu_h2) and the A-stationary N-tile
//   walk (n_inner > 1).  They are compiled out of the other instantiations: merely present in the code they cost the 64 -> 64 body convs
//   4-8 % (eager pass 3.88 -> 4.24 ms for four launches; registers / issue slots of the issuer and the direct-store epilogue).
template <bool HAS_LN, int TC_EPI_SLOTS, bool LN_SMEM = false, bool XTRA = false>
__global__ void __launch_bounds__(128 + 128 * TC_EPI_SLOTS + (LN_SMEM ? 128 : 0), 1) conv_gemm_tcgen05_kernel(const __grid_constant__ TcParams p) {
  constexpr int TC_EPI_WARPS = 4 * TC_EPI_SLOTS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[TC_MAX_A_SLOTS], a_empty[TC_MAX_A_SLOTS], tmem_full[4], tmem_empty[4];
  __shared__ __align__(8) uint64_t b_full[TC_MAX_STEPS], b_empty[TC_MAX_STEPS];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[512];                 // this CTA's N tile(s): BN <= 256 columns, or n_inner * BN <= 512
  __shared__ __align__(8) uint64_t res_bar[TC_MAX_EPI_WARPS];
  __shared__ __align__(8) uint64_t a_norm[LN_SMEM ? TC_MAX_A_SLOTS : 1];
  __shared__ __align__(16) float ln_gb[LN_SMEM ? 512 : 4];   // gamma [0, 256) | beta [256, 512)

  // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B the 128B swizzle needs
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const uint32_t b_bytes = (uint32_t)BN * TC_BK * 2;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.a_slots * p.a_slot_bytes;
  uint8_t* smem_stage = smem_b + (size_t)p.b_slots * b_bytes;   // TC_EPI_WARPS x stage_bufs x 4 KB staging sub-tiles
  const int acc_cols = p.nacc * BN;
  const uint32_t tmem_cols = (acc_cols <= 32) ? 32 : (acc_cols <= 64 ? 64 : (acc_cols <= 128 ? 128 : (acc_cols <= 256 ? 256 : 512)));
  const int nacc_mask = p.nacc - 1, nacc_shift = p.nacc_shift;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC_MAX_SRC; ++s) tma_prefetch_desc(&p.amap[s]);
    tma_prefetch_desc(&p.bmap);
    if (p.tma_store) { tma_prefetch_desc(&p.omap[0]); tma_prefetch_desc(&p.rmap); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_slots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_slots; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 4; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], TC_EPI_WARPS); }
    for (int a = 0; a < TC_EPI_WARPS; ++a) mbar_init(&res_bar[a], 1);
    if (LN_SMEM)
      for (int s = 0; s < p.a_slots; ++s) mbar_init(&a_norm[s], 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp >= 4) {   // this CTA owns N-tile blockIdx.x % n_tiles_n for its whole life: stage its bias once
    const int i = threadIdx.x - 128;
    if (XTRA) {
      for (int j = i; j < 512; j += 32 * TC_EPI_WARPS)
        bias_s[j] = ((p.bias && j < p.n_inner * BN) ? __ldg(p.bias + (blockIdx.x % p.n_tiles_n) * BN + j) : 0.f) * (p.gelu_h2 ? 0.5f : 1.f);
    } else if (i < 256) bias_s[i] = (p.bias && i < (p.tapsum ? p.Cout_store : BN)) ? __ldg(p.bias + (blockIdx.x % p.n_tiles_n) * BN + i) : 0.f;
  }
  if (LN_SMEM && warp >= 4 + TC_EPI_WARPS)
    for (int j = threadIdx.x - (128 + 32 * TC_EPI_WARPS); j < 64 * p.nsteps; j += 128) { ln_gb[j] = __ldg(p.ln_gamma + j); ln_gb[256 + j] = __ldg(p.ln_beta + j); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // static tile schedule: this CTA owns one N-tile (so resident weights stay valid) and every
  // (gridDim/n_tiles_n)-th M-tile; neighbouring CTAs share A tiles in L2.
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int nt = blockIdx.x % p.n_tiles_n;
  const int mt0 = blockIdx.x / p.n_tiles_n, mt_step = gridDim.x / p.n_tiles_n;
  const int units = p.halo ? p.nchunks : p.nsteps;   // A-slot fills per tile
  const int taps = p.halo ? 9 : 1;                   // K-steps served by one A slot

  if (warp == 0) {
    // ================= A producer (TMA); warp-uniform loop, one elected lane issues =================
    int slot = 0;
    uint32_t phase = 0;
    const int halo = p.halo, a_slots = p.a_slots, step_x = p.step_x, step_y = p.step_y, org = p.org, m_tiles = p.m_tiles, a_mul = p.a_mul;
    const uint32_t a_slot_bytes = (uint32_t)p.a_slot_bytes, a_box_bytes = (uint32_t)p.a_box_bytes;
    const bool noload = (p.debug & 4) != 0;
    const uint32_t bar_ae = smem_u32(&a_empty[0]);
    TileIter ti(mt0, mt_step, p.tiles_x, p.tiles_y);
    for (int mt = mt0; mt < m_tiles; mt += mt_step, ti.next()) {
      const int img = ti.img;
      const int y0 = ti.ty * step_y + org, x0 = ti.tx * step_x + org;
      for (int u = 0; u < units; ++u) {
        mbar_wait_a(bar_ae + (uint32_t)slot * 8u, phase ^ 1);
        uint64_t* afull = &a_full[slot];
        if (noload) {
          if (elect_one()) mbar_arrive(afull);
        } else if (elect_one()) {
          uint8_t* sa = smem_a + (size_t)slot * a_slot_bytes;
          if (halo) {
            const Chunk ch = p.chunks[u];
            if (halo == 1) {
              mbar_expect_tx(afull, 3u * (uint32_t)TC_HALO_COPY);
#pragma unroll
              for (int d = 0; d < 3; ++d) tma_load_4d(sa + d * TC_HALO_COPY, &p.amap[ch.src], afull, ch.c0, x0 + d - 1, y0 - 1, img);
            } else {
              mbar_expect_tx(afull, (uint32_t)TC_HALO_WSLOT);
              tma_load_4d(sa, &p.amap[ch.src], afull, ch.c0, x0 - 4, y0 - 1, img);
            }
          } else {
            const KStep ks = p.steps[u];
            mbar_expect_tx(afull, a_box_bytes);
            tma_load_4d(sa, &p.amap[ks.src], afull, ks.c0, x0 * a_mul + ks.dx, y0 * a_mul + ks.dy, img);
          }
        }
        __syncwarp();
        if (++slot == a_slots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ================= B producer (TMA): weight slabs [BN x 64] =================
    auto bk = [&](int s) -> int {   // K offset of K-step s in the packed weights
      if (!p.halo) return s * TC_BK;
      const int u = s / 9, t = s - u * 9;
      return t * p.ctot + p.chunks[u].cg;
    };
    if (XTRA && p.b_resident && p.n_inner > 1) {
      if (mt0 < p.m_tiles)
        for (int ni = 0; ni < p.n_inner; ++ni)
          for (int s = 0; s < p.nsteps; ++s) {
            const int sl = ni * p.nsteps + s;
            if (elect_one()) {
              mbar_expect_tx(&b_full[sl], b_bytes);
              tma_load_2d(smem_b + (size_t)sl * b_bytes, &p.bmap, &b_full[sl], bk(s), ni * BN);
            }
            __syncwarp();
          }
    } else if (p.b_resident) {
      if (mt0 < p.m_tiles)
        for (int s = 0; s < p.nsteps; ++s) {
          if (elect_one()) {
            if (p.tapsum) {   // weights [co][tap][ci] viewed as {ci, co, tap}: the box lands as rows tap * Ct + co
              mbar_expect_tx(&b_full[s], (uint32_t)(9 * p.tapsum * TC_BK * 2));
              tma_load_3d(smem_b + (size_t)s * b_bytes, &p.bmap, &b_full[s], s * TC_BK, 0, 0);
            } else {
              mbar_expect_tx(&b_full[s], b_bytes);
              tma_load_2d(smem_b + (size_t)s * b_bytes, &p.bmap, &b_full[s], bk(s), nt * BN);
            }
          }
          __syncwarp();
        }
    } else {
      const uint32_t bmask = (uint32_t)p.b_slots - 1, bshift = (uint32_t)__ffs(p.b_slots) - 1;
      uint32_t g = 0;
      for (int mt = mt0; mt < p.m_tiles; mt += mt_step)
        for (int s = 0; s < p.nsteps; ++s, ++g) {
          const uint32_t slot = g & bmask;
          mbar_wait(&b_empty[slot], ((g >> bshift) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&b_full[slot], b_bytes);
            tma_load_2d(smem_b + (size_t)slot * b_bytes, &p.bmap, &b_full[slot], bk(s), nt * BN);
          }
          __syncwarp();
        }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma =================
    // All loop state is warp-uniform (ring slots are derived from a running step counter, b_slots is a
    // power of two), so everything inside the elected block stays in uniform registers.
    const uint32_t idesc = make_idesc_bf16(BN);
    const uint64_t desc_hi = make_sw128_desc(0);       // descriptor with a zero start address
    const uint64_t desc_wide = (desc_hi & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)(TC_HALO_WROW >> 4) << 32);   // 2 KB between 8-row atoms
    const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
    const uint32_t bmask = (uint32_t)p.b_slots - 1, bshift = (uint32_t)__ffs(p.b_slots) - 1;
    int aslot = 0;
    uint32_t aphase = 0, gs = 0;                       // gs = B slabs consumed so far (ring mode)
    int it = 0;
    if (p.b_resident) {
      if (mt0 < p.m_tiles)
        for (int s = 0; s < p.nsteps * (XTRA ? p.n_inner : 1); ++s) mbar_wait(&b_full[s], 0);
      const int halo = p.halo, a_slots = p.a_slots, m_tiles = p.m_tiles;
      const uint32_t a_slot16 = (uint32_t)p.a_slot_bytes >> 4;
      const bool nomma = (p.debug & 2) != 0;
      // LN_SMEM: an A slot is ready once the LayerNorm warps have normalised it
      const uint32_t bar_af = smem_u32(LN_SMEM ? &a_norm[0] : &a_full[0]), bar_ae = smem_u32(&a_empty[0]), bar_tf = smem_u32(&tmem_full[0]), bar_te = smem_u32(&tmem_empty[0]);
#define FBANET_MMA_LOOP(T, H) \
  mma_loop_resident<T, H, XTRA && H == 0>(mt0, mt_step, m_tiles, units, a_slots, a_slot16, sa0 >> 4, sb0 >> 4, b_bytes >> 4, tmem_base, (uint32_t)BN, idesc, bar_af, bar_ae, bar_tf, bar_te, nomma, (uint32_t)nacc_mask, (uint32_t)nacc_shift, p.n_inner)
      if (halo == 0) FBANET_MMA_LOOP(1, 0);
      else if (halo == 1) FBANET_MMA_LOOP(9, 1);
      else if (halo == 3) FBANET_MMA_LOOP(9, 3);
      // (halo mode 2, descriptor base offset = row phase, was the experiment that proved the swizzle address based: the launcher rejects it)
#undef FBANET_MMA_LOOP
    } else {
    for (int mt = mt0; mt < p.m_tiles; mt += mt_step, ++it) {
      const int acc = it & nacc_mask;
      const uint32_t acc_phase = (it >> nacc_shift) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);     // epilogue has drained this accumulator
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int u = 0; u < units; ++u) {
        mbar_wait(LN_SMEM ? &a_norm[aslot] : &a_full[aslot], aphase);
        tc_fence_after();
        const uint32_t sa = sa0 + (uint32_t)aslot * (uint32_t)p.a_slot_bytes;
        if (elect_one()) {
          for (int t = 0; t < taps; ++t) {
            const int s = u * taps + t;
            const uint32_t g = gs + (uint32_t)t, slot = g & bmask;
            mbar_wait(&b_full[slot], (g >> bshift) & 1);
            tc_fence_after();
            const uint32_t sb = sb0 + slot * b_bytes;
            // halo mode: tap (ky,kx) reads the dx-shifted copy kx, starting ky rows (1 KB each) down
            uint32_t a_addr = sa;
            uint64_t dA = desc_hi;
            if (p.halo == 1) a_addr += (uint32_t)((t % 3) * TC_HALO_COPY + (t / 3) * 1024);
            else if (p.halo >= 2) {
              a_addr += (uint32_t)((t / 3) * TC_HALO_WROW + (3 + t % 3) * 128);
              dA = desc_wide + (p.halo == 2 ? ((uint64_t)((3 + t % 3) & 7) << 49) : 0ull);
            }
            const uint64_t adesc = dA + (uint64_t)(a_addr >> 4), bdesc = desc_hi + (uint64_t)(sb >> 4);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)       // advance 32 bytes (16 bf16) inside the swizzle row
              if (!(p.debug & 2)) umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((s | k) != 0));
            umma_commit(&b_empty[slot]);
          }
          umma_commit(&a_empty[aslot]);                             // frees the A slot when these MMAs retire
          if (u == units - 1) umma_commit(&tmem_full[acc]);         // accumulator ready for the epilogue
        }
        __syncwarp();
        gs += (uint32_t)taps;
        if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
      }
    }
    }
  } else if (LN_SMEM && warp >= 4 + TC_EPI_WARPS) {
    // ================= LayerNorm warps: normalise every A tile in place before the issuer reads it =================
    const int r = threadIdx.x - (128 + 32 * TC_EPI_WARPS);
    const uint32_t sa0 = smem_u32(smem_a);
    if (p.nsteps == 1) ln_smem_loop<1>(p, sa0, a_full, a_norm, ln_gb, mt0, mt_step, r, lane);
    else if (p.nsteps == 2) ln_smem_loop<2>(p, sa0, a_full, a_norm, ln_gb, mt0, mt_step, r, lane);
    else ln_smem_loop<4>(p, sa0, a_full, a_norm, ln_gb, mt0, mt_step, r, lane);
  } else if (warp >= 4 && warp < 4 + TC_EPI_WARPS) {
    // ================= epilogue: 4 TMEM lane quarters x TC_EPI_SLOTS warps; column pieces are dealt round-robin to the slots =================
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int slot = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int ly = row / p.tw, lx = row - ly * p.tw;
    const float alpha = (p.act == FBANET_ACT_PRELU) ? __ldg(p.alpha) : 0.f;
    int it = 0;
    const uint32_t bar_tf = smem_u32(&tmem_full[0]), bar_te = smem_u32(&tmem_empty[0]);   // raw shared addresses (see mma_loop_resident)
    if (p.debug & 1) {
      for (int mt = mt0; mt < p.m_tiles; mt += mt_step)
      for (int ni = 0; ni < (XTRA ? p.n_inner : 1); ++ni, ++it) {
        mbar_wait(&tmem_full[it & nacc_mask], (it >> nacc_shift) & 1);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[it & nacc_mask]);
      }
    } else if (p.tapsum) {
      // ---- TAPSUM epilogue: 3x3 convs with a handful of outputs (the FAF score conv 64 -> 2, the final conv 64 -> 3 as hi/lo
      // halves) are operand-fetch bound as implicit GEMMs: 36 MMAs per 128 pixels whatever N is (42 cycles each at N = 16,
      // tools/ubench/umma.cu).  Here the nine taps are stacked along N instead -- B rows = tap * Ct + co -- so ONE pass of
      // K/16 MMAs over a tile of 16 x 8 INPUT pixels yields P[q][tap][co] = W_tap . in[q] for every tap at once, and the
      // convolution is the shifted sum  out[p][co] = bias + sum_tap P[p + off_tap][tap][co]  over the tile's 14 x 6 interior,
      // done here through shared memory (rows padded to BN + 1 floats: the gather walks rows at a fixed column).
      // Whole tiles alternate between the two sets of four epilogue warps; the other set only hands the accumulator back.
      const int Ct = p.tapsum;
      const int pst = BN + 1;
      float* P = reinterpret_cast<float*>(smem_stage) + (size_t)slot * 128 * pst;
      const int bar_id = 1 + slot;
      TileIter ti(mt0, mt_step, p.tiles_x, p.tiles_y);
      for (int mt = mt0; mt < p.m_tiles; mt += mt_step, ++it, ti.next()) {
        const int acc = it & nacc_mask;
        const uint32_t acc_phase = (it >> nacc_shift) & 1;
        const bool mine = (it % TC_EPI_SLOTS) == slot;
        mbar_wait_a(bar_tf + (uint32_t)acc * 8u, acc_phase);
        tc_fence_after();
        if (mine) {
          const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
          const uint32_t prow = smem_u32(P) + (uint32_t)(row * pst * 4);   // explicit shared-space stores / loads below: through the generic
                                                                           // pointer the gather compiled to LD.E (long-scoreboard latency)
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            const int nc = BN - c0 >= 32 ? 32 : 16;
            if (nc == 32) tmem_ld32(taddr0 + c0, v); else tmem_ld16(taddr0 + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nc) asm volatile("st.shared.b32 [%0], %1;" ::"r"(prow + (uint32_t)((c0 + j) * 4)), "r"(v[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bar_te + (uint32_t)acc * 8u);
        if (!mine) continue;
        named_bar_sync(bar_id, 128);                       // the four lane quarters of this tile have written P
        const int img = ti.img;
        const int oy = ti.ty * p.step_y + ly - 1, ox = ti.tx * p.step_x + lx - 1;
        if (lx >= 1 && lx <= 14 && ly >= 1 && ly <= 6 && oy < p.Ho && ox < p.Wo) {
          float* op = reinterpret_cast<float*>(p.out) + img * p.out_img_stride + ((int64_t)oy * p.Wo + ox) * p.out_ld;
          const uint32_t pc_a = smem_u32(P) + (uint32_t)((((ly - 1) * 16 + (lx - 1)) * pst) * 4);   // P row of tap (0,0) for this output pixel
          auto pc = [&](int off) -> float {
            float f;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(pc_a + (uint32_t)(off * 4)));
            return f;
          };
          auto gather = [&](auto ct_tag) {
            constexpr int CT = decltype(ct_tag)::value;             // compile-time outputs per tap: all 9 * CT loads independent
            float sacc[CT];
#pragma unroll
            for (int j = 0; j < CT; ++j) sacc[j] = bias_s[j];
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
              for (int j = 0; j < CT; ++j) sacc[j] += pc(((t / 3) * 16 + (t % 3)) * pst + t * CT + j);
            if (CT == 2) *reinterpret_cast<float2*>(op) = make_float2(sacc[0], sacc[1]);
            else if (CT == 4) *reinterpret_cast<float4*>(op) = make_float4(sacc[0], sacc[1], sacc[2], sacc[3]);
            else if (CT == 8) {
              *reinterpret_cast<float4*>(op) = make_float4(sacc[0], sacc[1], sacc[2], sacc[3]);
              *reinterpret_cast<float4*>(op + 4) = make_float4(sacc[4], sacc[5], sacc[6], sacc[7]);
            } else {
#pragma unroll
              for (int j = 0; j < CT; ++j) op[j] = sacc[j];
            }
          };
          auto gather_fold = [&](auto h_tag) {                       // Ct = 2h rows: [0, h) hi, [h, 2h) lo; out[j] = hi_j + lo_j
            constexpr int HH = decltype(h_tag)::value;
            float sacc[HH];
#pragma unroll
            for (int j = 0; j < HH; ++j) sacc[j] = bias_s[j] + bias_s[HH + j];
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
              for (int j = 0; j < HH; ++j) {
                const int q = ((t / 3) * 16 + (t % 3)) * pst + t * 2 * HH + j;
                sacc[j] += pc(q) + pc(q + HH);
              }
            if constexpr (HH == 1) op[0] = sacc[0];
            else if constexpr (HH == 2) *reinterpret_cast<float2*>(op) = make_float2(sacc[0], sacc[1]);
            else if constexpr (HH == 3) *reinterpret_cast<float4*>(op) = make_float4(sacc[0], sacc[1], sacc[2], 0.f);
            else *reinterpret_cast<float4*>(op) = make_float4(sacc[0], sacc[1], sacc[2], sacc[3]);
          };
          if (p.fold) {
            switch (Ct) {
              case 2: gather_fold(std::integral_constant<int, 1>{}); break;
              case 4: gather_fold(std::integral_constant<int, 2>{}); break;
              case 6: gather_fold(std::integral_constant<int, 3>{}); break;
              default: gather_fold(std::integral_constant<int, 4>{}); break;
            }
          } else
          switch (Ct) {
            case 1: gather(std::integral_constant<int, 1>{}); break;
            case 2: gather(std::integral_constant<int, 2>{}); break;
            case 3: gather(std::integral_constant<int, 3>{}); break;
            case 4: gather(std::integral_constant<int, 4>{}); break;
            case 8: gather(std::integral_constant<int, 8>{}); break;
            default:
              for (int j = 0; j < Ct; ++j) {
                float sj = bias_s[j];
#pragma unroll
                for (int t = 0; t < 9; ++t) sj += pc(((t / 3) * 16 + (t % 3)) * pst + t * Ct + j);
                op[j] = sj;
              }
          }
        }
        named_bar_sync(bar_id, 128);                       // P is free for this set's next tile
      }
    } else if (p.tma_store) {
      // ---- per-warp staged epilogue.  A warp owns 32 pixel rows of the tile (its TMEM lane quarter), which form a
      // rectangle {bw, 32/bw} of the output image, so it can stage and TMA-store its own 32 x 64-column sub-tiles (4 KB,
      // SWIZZLE_128B) with no CTA-level barrier at all: tcgen05.ld -> bias/act(/residual) -> st.shared -> fence ->
      // __syncwarp -> one bulk store.  Two staging buffers per warp: the store of chunk k drains while chunk k+1 is computed.
      // The residual sub-tile is TMA-loaded into the staging buffer first and added in place.  64-column chunks are dealt
      // round-robin to the warps of a lane quarter, counted across tiles so any chunks-per-tile count balances.
      // (Per-thread 16-byte global stores touch 32 lines per instruction and cap an SM at ~16 B/clk.)
      const int ew = warp - 4;                             // 0 .. TC_EPI_WARPS-1
      const uint32_t nbufs = (uint32_t)p.stage_bufs;
      uint8_t* stage0 = smem_stage + ew * 4096 * nbufs;
      const uint32_t stage_u = smem_u32(stage0);
      const int r7 = lane & 7;
      const int nchunks = BN >> 6;
      const bool has_res = p.residual != nullptr;
      const bool ct2 = p.store_mode == FBANET_STORE_CONVT2;
      const int Co = ct2 ? (p.Cout >> 2) : p.Cout;
      const int wy0 = (q * 32) / p.tw, wx0 = (q * 32) % p.tw;   // this warp's rectangle inside the tile
      const int bw = p.tw < 32 ? p.tw : 32;
      const int ply = lane / bw, plx = lane - ply * bw;         // this thread's pixel inside the rectangle
      uint32_t res_phase = 0, nb = 0;
      TileIter ti(mt0, mt_step, p.tiles_x, p.tiles_y);
      const int n_inner = XTRA ? p.n_inner : 1;
      for (int mt = mt0; mt < p.m_tiles; mt += mt_step, ti.next())
      for (int ni = 0; ni < n_inner; ++ni, ++it) {         // (A-stationary mode: the tile's n_inner N tiles arrive one accumulator after the other)
        const int acc = it & nacc_mask;
        const uint32_t acc_phase = (it >> nacc_shift) & 1;
        const int img = ti.img;
        const int y0 = ti.ty * p.th + wy0, x0 = ti.tx * p.tw + wx0;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
        float ln_rstd = 0.f;   // 1/sigma of this thread's row, fetched while the MMAs still run
        if (HAS_LN && y0 + ply < p.Ho && x0 + plx < p.Wo)
          ln_rstd = __ldg(p.ln_stats + (((int64_t)img * p.Ho + (y0 + ply)) * p.Wo + (x0 + plx)) * 2 + 1);
        const f32x2 ln_r2 = pack_f2(ln_rstd, ln_rstd);
        bool waited = false;
        for (int cidx = 0; cidx < nchunks; ++cidx) {
          if (((it * nchunks + cidx) % TC_EPI_SLOTS) != slot) continue;
          const int col0 = (nt + ni) * BN + cidx * 64;   // GEMM column of the chunk
          const uint32_t boff = (nbufs == 2 ? (nb & 1u) : 0u) * 4096u;
          const uint32_t sbuf = stage_u + boff;
          uint8_t* gbuf = stage0 + boff;
          ++nb;
          if (lane == 0) {   // the store that last used this buffer has read it
            if (nbufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (has_res) {
              mbar_expect_tx(&res_bar[ew], 4096u);
              tma_load_4d(gbuf, &p.rmap, &res_bar[ew], col0, x0, y0, img);
            }
          }
          __syncwarp();
          if (!waited) { mbar_wait_a(bar_tf + (uint32_t)acc * 8u, acc_phase); tc_fence_after(); waited = true; }
          uint32_t v[64];
          tmem_ld32(taddr0 + cidx * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld32(taddr0 + cidx * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
          if (has_res) { mbar_wait(&res_bar[ew], res_phase); res_phase ^= 1; }
          const float* bs = bias_s + ni * BN + cidx * 64;
          const uint32_t row_addr = sbuf + (uint32_t)lane * 128u;
#pragma unroll
          for (int c = 0; c < 8; ++c) {                  // 8 columns = one 16-byte smem chunk at a time
            const float4 b0 = *reinterpret_cast<const float4*>(bs + c * 8), b1 = *reinterpret_cast<const float4*>(bs + c * 8 + 4);
            f32x2 f[4], bb[4] = {pack_f2(b0.x, b0.y), pack_f2(b0.z, b0.w), pack_f2(b1.x, b1.y), pack_f2(b1.z, b1.w)};
            const uint32_t saddr = row_addr + (uint32_t)((c ^ r7) << 4);   // SWIZZLE_128B position of chunk c in this row
            if (XTRA && !HAS_LN && p.gelu_h2) {
              // fp16 store + tanh-GELU (the dim-256 LeFF's fc1): z = (acc + bias) / 2 in fp32 (bias_s holds bias / 2), rounded ONCE to fp16,
              // then GELU(x) = z (1 + tanh(z (2 k0 + 8 k0 k1 z^2))) on packed half2 -- 1 FFMA2 + 1 F2FP + 4 HFMA2/HMUL2 + ONE MUFU.TANH.F16x2 per
              // pair instead of 6 packed fp32 ops + 2 MUFU.TANH + the pack: this epilogue is what bounds the GEMM (XU and issue slots).
              const uint32_t A2 = 0x3E623E62u, B2 = 0x34913491u;           // half2 (2 k0, 2 k0) and (8 k0 k1, 8 k0 k1)
              uint32_t h[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                h[e] = gelu_half_h2(f2_to_f16x2(fma_f2(pack_f2(__uint_as_float(v[c * 8 + 2 * e]), __uint_as_float(v[c * 8 + 2 * e + 1])), pack_f2(0.5f, 0.5f), bb[e])), A2, B2);
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]));
              continue;
            }
            if (HAS_LN) {   // folded LayerNorm (row-centred weights): rstd * acc + bias
#pragma unroll
              for (int e = 0; e < 4; ++e) f[e] = fma_f2(pack_f2(__uint_as_float(v[c * 8 + 2 * e]), __uint_as_float(v[c * 8 + 2 * e + 1])), ln_r2, bb[e]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) f[e] = add_f2(pack_f2(__uint_as_float(v[c * 8 + 2 * e]), __uint_as_float(v[c * 8 + 2 * e + 1])), bb[e]);
            }
            if (p.act == FBANET_ACT_GELU_TANH) {
#pragma unroll
              for (int e = 0; e < 4; ++e) f[e] = gelu_tanh_fast_f2(f[e]);
            } else if (p.act != FBANET_ACT_NONE) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float lo, hi;
                unpack_f2(f[e], lo, hi);
                if (p.act == FBANET_ACT_RELU) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                else if (p.act == FBANET_ACT_PRELU) { lo = lo > 0.f ? lo : alpha * lo; hi = hi > 0.f ? hi : alpha * hi; }
                else { lo = gelu_erf(lo); hi = gelu_erf(hi); }
                f[e] = pack_f2(lo, hi);
              }
            }
            if (has_res) {
              uint4 rv;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rv.x), "=r"(rv.y), "=r"(rv.z), "=r"(rv.w) : "r"(saddr));
              f[0] = add_f2(f[0], bf16x2_to_f2(rv.x)); f[1] = add_f2(f[1], bf16x2_to_f2(rv.y));
              f[2] = add_f2(f[2], bf16x2_to_f2(rv.z)); f[3] = add_f2(f[3], bf16x2_to_f2(rv.w));
            }
            if (p.store_f16)
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(f2_to_f16x2(f[0])), "r"(f2_to_f16x2(f[1])),
                           "r"(f2_to_f16x2(f[2])), "r"(f2_to_f16x2(f[3])));
            else
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(f2_to_bf16x2(f[0])), "r"(f2_to_bf16x2(f[1])),
                           "r"(f2_to_bf16x2(f[2])), "r"(f2_to_bf16x2(f[3])));
          }
          fence_proxy_async();                           // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            const int qq = ct2 ? col0 / Co : 0;          // 2x2 scatter store: sub-pixel plane of this chunk (0 for NHWC)
            tma_store_4d(&p.omap[qq], gbuf, col0 - qq * Co, x0, y0, img);
            bulk_commit();
          }
        }
        if (!waited) { mbar_wait_a(bar_tf + (uint32_t)acc * 8u, acc_phase); tc_fence_after(); }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bar_te + (uint32_t)acc * 8u);
      }
      if (lane == 0) bulk_wait0();                       // all stores of this warp have completed
      __syncwarp();
    } else {
      const int npieces = BN >= 32 ? BN / 32 : 1;      // 32-column pieces (one 16-column piece for BN = 16)
      TileIter ti(mt0, mt_step, p.tiles_x, p.tiles_y);
      const int n_inner = XTRA ? p.n_inner : 1;
      for (int mt = mt0; mt < p.m_tiles; mt += mt_step, ti.next())
      for (int ni = 0; ni < n_inner; ++ni, ++it) {
        const int acc = it & nacc_mask;
        const uint32_t acc_phase = (it >> nacc_shift) & 1;
        const int img = ti.img;
        const int y = ti.ty * p.th + ly, x = ti.tx * p.tw + lx;
        const bool valid = (ly < p.th) && (y < p.Ho) && (x < p.Wo);
        // this warp's first piece of the tile; its residual is fetched while the MMAs of the tile are still running
        int j0 = (slot - (it * npieces) % TC_EPI_SLOTS + TC_EPI_SLOTS) % TC_EPI_SLOTS;
        uint4 rpre[4];
        const bool pre = valid && p.residual != nullptr && BN >= 32 && j0 < npieces;
        if (pre) {
          const bf16* rp = p.residual + img * p.res_img_stride + ((int64_t)y * p.Wo + x) * p.res_ld + (nt + ni) * BN + j0 * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) rpre[j] = *reinterpret_cast<const uint4*>(rp + j * 8);
        }
        float ln_rstd = 1.f;
        if (HAS_LN && valid) ln_rstd = __ldg(p.ln_stats + (((int64_t)img * p.Ho + y) * p.Wo + x) * 2 + 1);
        mbar_wait_a(bar_tf + (uint32_t)acc * 8u, acc_phase);
        tc_fence_after();
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
        for (int j = j0; j < npieces; j += TC_EPI_SLOTS) {
          const int c0 = j * 32;
          uint32_t v[32];
          const int nc = (BN - c0 >= 32) ? 32 : 16;
          if (nc == 32) tmem_ld32(taddr0 + c0, v); else tmem_ld16(taddr0 + c0, v);
          tmem_ld_wait();
          if (valid) epilogue_chunk<XTRA>(p, v, nc, (nt + ni) * BN + c0, img, y, x, alpha, bias_s + ni * BN + c0, (pre && j == j0) ? rpre : nullptr,
                                    HAS_LN, ln_rstd);
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bar_te + (uint32_t)acc * 8u);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

static int conv_ctot(const fbanet_conv_params* p) {
  int c = 0;
  for (int s = 0; s < p->nsrc; ++s) c += p->src_s2d ? p->src[s].C / 4 : p->src[s].C;
  return c;
}

// N tile.  3x3 halo convs with a deep K (>= 128 input channels) take full 256-column tiles: the A halo is then fetched once
// per pixel tile instead of once per 128 output channels and the MMAs run at N = 256 (0.84 -> 0.66 ms on 512->256 @80x80,
// 1475 TFLOP/s); with a shallow K (the 64->256 tail convs) the longer epilogue per tile is not hidden and 128 stays better.
static int pick_bn(int cout, bool halo, int ctot, int64_t pixels) {
  int cap = (halo && ctot < 128) ? 128 : 256;
  // 1x1 GEMMs over 256 input channels: with 256-column tiles the resident weights (128 KB) leave three A slots and one staging buffer
  // per epilogue warp; 192 / 128 columns measured 5-15 % faster (profiles/r2_bn_cap_1x1.log: qkv 256->768 0.213 -> 0.196 ms,
  // proj 256->256 0.130 -> 0.110, fc1 256->1024 0.346 -> 0.326)
  // (at 40^2 x 64 bursts the narrower tiles lose instead: fc1 0.438 -> 0.479 ms -- only for >= 200 k pixels)
  if (!halo && ctot >= 256 && pixels >= 200000) cap = 192;
  static const char* bnenv = getenv("FBANET_TC_BN");   // experiment switch: cap of the N tile of 1x1 GEMMs (64 / 128 / 192 / 256)
  if (!halo && bnenv) { const int v = atoi(bnenv); if (v >= 64 && v <= 256) cap = v; }
  if (cout <= cap) return cout;
  for (int bn = cap; bn >= 64; bn -= 64)
    if (cout % bn == 0) return bn;
  return 0;
}

static void pick_tile(int H, int W, int* tw_out, int* th_out) {
  const int cands[] = {8, 16, 32, 64, 128};
  int64_t best = -1;
  auto consider = [&](int tw) {
    if (tw < 1 || tw > 128) return;
    int th = 128 / tw;
    if (th > 256) return;
    const int64_t cost = (int64_t)((W + tw - 1) / tw) * ((H + th - 1) / th);  // #tiles, each costs a full 128-row MMA
    if (best < 0 || cost < best) { best = cost; *tw_out = tw; *th_out = th; }
  };
  for (int c : cands) consider(c);
  if (W <= 128) consider(W);
}

static bool is_halo(const fbanet_conv_params* p) { return p->stride == 1 && p->KH == 3 && p->KW == 3 && p->pad == 1 && !p->src_s2d; }

// conv kinds the tensor-core path takes
static bool tc_shape_ok(const fbanet_conv_params* p) {
  if (p->dtype != FBANET_BF16) return false;
  if (p->nsrc > TC_MAX_SRC) return false;
  if (p->store_mode == FBANET_STORE_PS2) return false;
  const bool s1 = p->stride == 1 && p->KH == p->KW && (p->KH == 1 || p->KH == 3) && p->pad == p->KH / 2;
  const bool s2d = p->src_s2d && p->stride == 2 && p->KH == 4 && p->KW == 4 && p->pad == 1;
  // 4x4 stride-2 conv straight from the full-resolution source: every tap is a TMA box that steps two pixels (elementStrides {1,2,2,1})
  const bool s2 = !p->src_s2d && p->stride == 2 && p->KH == 4 && p->KW == 4 && p->pad == 1 && !(p->H & 1) && !(p->W & 1);
  if (!s1 && !s2d && !s2) return false;
  if (p->src_s2d && !s2d) return false;
  int ctot = 0;
  for (int s = 0; s < p->nsrc; ++s) {
    const fbanet_src& S = p->src[s];
    if (S.row_scale) return false;
    const int cc = s2d ? S.C / 4 : S.C;
    if (cc % TC_BK || (s2d && S.C % 4)) return false;
    if ((S.ld % 8) || (S.img_stride % 8) || ((uintptr_t)S.ptr % 16)) return false;
    ctot += cc;
  }
  const int taps = p->KH * p->KW;
  if (taps * (ctot / TC_BK) > TC_MAX_STEPS) return false;
  if (is_halo(p) && ctot / TC_BK > TC_MAX_CHUNKS) return false;
  const int bn = pick_bn(p->Cout, is_halo(p), conv_ctot(p), (int64_t)p->N * p->Ho * p->Wo);
  if (bn != 16 && bn != 64 && bn != 128 && bn != 192 && bn != 256) return false;
  if ((uintptr_t)p->weight % 16) return false;
  if (p->store_mode == FBANET_STORE_NHWC || p->store_mode == FBANET_STORE_CONVT2) {
    if ((p->out_ld % 8) || (p->out_img_stride % 8) || ((uintptr_t)p->out % 16)) return false;
    if (p->Cout_store % 8) return false;
    if (p->store_mode == FBANET_STORE_CONVT2 && ((p->Cout / 4) % 32)) return false;
    if (p->residual && ((p->res_ld % 8) || (p->res_img_stride % 8) || ((uintptr_t)p->residual % 16))) return false;
  } else if (p->store_mode == FBANET_STORE_NCHW_BASE) {
    if (bn != 16 || p->Cout_store > 16) return false;
  } else if (p->store_mode == FBANET_STORE_NHWC_F32) {
    if (p->residual || ((uintptr_t)p->out % 4)) return false;
  }
  if (p->bias && ((uintptr_t)p->bias % 16)) return false;
  if (p->ln_stats && (p->KH != 1 || p->stride != 1 || p->store_mode != FBANET_STORE_NHWC || ((uintptr_t)p->ln_stats % 8))) return false;
  if (p->store_f16 && (p->store_mode != FBANET_STORE_NHWC || p->residual)) return false;
  if (p->ln_gamma) {   // LayerNorm of the A tile in shared memory: a 1x1 GEMM over ONE source of 64 / 128 / 256 channels
    if (!p->ln_beta || p->ln_stats || p->KH != 1 || p->stride != 1 || p->nsrc != 1 || p->src_s2d) return false;
    if (ctot != 64 && ctot != 128 && ctot != 256) return false;
    if (((uintptr_t)p->ln_gamma % 4) || ((uintptr_t)p->ln_beta % 4)) return false;
  }
  return get_encode() != nullptr;
}

int conv_gemm_tc_supported(const fbanet_conv_params* p) { return tc_shape_ok(p) ? 1 : 0; }

int conv_gemm_tc_launch(const fbanet_conv_params* p, cudaStream_t stream) {
  if (!tc_shape_ok(p)) return FBANET_E_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  static thread_local TcParams tp;  // ~2.6 KB; filled per call, passed by value
  memset(&tp, 0, sizeof(tp));
  const bool s2d = p->src_s2d != 0;
  // tap-stacked mode for 3x3 convs with a handful of fp32 outputs (see the TAPSUM epilogue); FBANET_TC_TAPSUM=0 = plain halo conv
  const char* tsenv = getenv("FBANET_TC_TAPSUM");
  // (measured, batch 64: score conv 64 -> 2 @160^2 x 896 frames 1.67 -> 1.08 ms; with 8 outputs per tap -- the final conv's hi/lo
  // rows -- the shared-memory gather costs what the MMAs save, 2.0 vs 1.9 ms, so the mode is limited to <= 4 outputs)
  const bool fold = p->fold_hi_lo != 0;
  if (fold && (p->Cout_store < 2 || p->Cout_store > 8 || (p->Cout_store & 1))) return FBANET_E_BADSHAPE;
  const bool tapsum = is_halo(p) && p->store_mode == FBANET_STORE_NHWC_F32 && p->Cout_store >= 1 && p->Cout_store <= (fold ? 8 : 4) &&
                      p->Cout_store <= p->Cout && !p->residual && p->act == FBANET_ACT_NONE && !(tsenv && tsenv[0] == '0') &&
                      ((uintptr_t)p->out % 16) == 0 && (p->out_ld % ((fold ? p->Cout_store / 2 : p->Cout_store) >= 3 ? 4 : 2)) == 0 && (p->out_img_stride % 4) == 0;   // vector stores of the gather
  if (fold && !tapsum) return FBANET_E_UNSUPPORTED;
  const bool halo = is_halo(p) && !tapsum;
  // tile space = output pixels; for s2d sources the source view already has the output resolution
  const bool s2 = !s2d && p->stride == 2;   // element-strided boxes on the full-resolution source
  const int Hs = s2d ? p->H / 2 : p->H, Ws = s2d ? p->W / 2 : p->W;   // source view the tensor maps describe
  if ((s2 ? Hs / 2 : Hs) != p->Ho || (s2 ? Ws / 2 : Ws) != p->Wo) return FBANET_E_BADSHAPE;
  tp.a_mul = s2 ? 2 : 1;
  int tw = 16, th = 8;
  if (halo) { tw = TC_HALO_TW; th = TC_HALO_TH; } else if (tapsum) { tw = 16; th = 8; } else pick_tile(p->Ho, p->Wo, &tw, &th);
  tp.tw = tw; tp.th = th;
  tp.step_x = tapsum ? tw - 2 : tw; tp.step_y = tapsum ? th - 2 : th; tp.org = tapsum ? -1 : 0;
  tp.tapsum = tapsum ? p->Cout_store : 0;
  tp.fold = fold ? 1 : 0;
  tp.tiles_x = (p->Wo + tp.step_x - 1) / tp.step_x;
  tp.tiles_y = (p->Ho + tp.step_y - 1) / tp.step_y;
  tp.m_tiles = p->N * tp.tiles_x * tp.tiles_y;
  tp.N = p->N; tp.Ho = p->Ho; tp.Wo = p->Wo;
  tp.BN = tapsum ? (9 * p->Cout_store + 15) / 16 * 16 : pick_bn(p->Cout, halo, conv_ctot(p), (int64_t)p->N * p->Ho * p->Wo);
  tp.n_tiles_n = tapsum ? 1 : p->Cout / tp.BN;
  tp.Cout = p->Cout; tp.Cout_store = p->Cout_store;
  tp.a_box_bytes = tw * th * TC_BK * 2;
  // Measured (profiles/r1_notes.md): the wide box wins 4-16 % for N tiles of 128 and for N = 16, where TMA / L2 traffic
  // bounds the tile; at N = 64 it lost 7 % with the unrolled issuer loop of session 2.
  // The hardware swizzle is purely address based: base-offset 0 (mode 3) is the correct descriptor, mode 2 computes garbage.
  static const char* henv = getenv("FBANET_TC_HALO");   // experiment switch: 1 = three dx-shifted copies, 3 = one wide box
  // (with the rolled issuer loop the wide box is also the better one at N = 64: 64->64 body conv 1.80 -> 1.76 ms, A/B in one box, and
  // its 36 KB slots leave room for a third A slot there)
  const int halo_mode = halo ? (henv ? atoi(henv) : 3) : 0;
  if (halo && halo_mode != 1 && halo_mode != 3) return FBANET_E_UNSUPPORTED;
  tp.halo = halo_mode;

  int ctot = 0;
  for (int s = 0; s < p->nsrc; ++s) {
    const fbanet_src& S = p->src[s];
    const cuuint64_t dims[4] = {(cuuint64_t)S.C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)p->N};
    const cuuint64_t strides[3] = {(cuuint64_t)S.ld * 2, (cuuint64_t)S.ld * 2 * Ws, (cuuint64_t)S.img_stride * 2};
    // (with elementStrides e the box is given in UNSTRIDED elements and ceil(box / e) of them land in shared memory -- tools/ubench/tma_stride.cu)
    const cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)(halo_mode >= 2 ? 16 : (s2 ? 2 * tw : tw)), (cuuint32_t)(halo ? th + 2 : (s2 ? 2 * th : th)), 1};
    const cuuint32_t estr[4] = {1, s2 ? 2u : 1u, s2 ? 2u : 1u, 1};
    CUresult r = encode(&tp.amap[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(S.ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return FBANET_E_BADSHAPE;
    ctot += s2d ? S.C / 4 : S.C;
  }
  for (int s = p->nsrc; s < TC_MAX_SRC; ++s) tp.amap[s] = tp.amap[0];
  tp.ctot = ctot;
  const int taps = p->KH * p->KW;
  const int K = taps * ctot;
  if (tapsum) {
    // weights [Cout][tap][ctot] seen as {ci, co, tap}: one box {64, Ct, 9} lands as B rows tap * Ct + co
    const cuuint64_t dims[3] = {(cuuint64_t)ctot, (cuuint64_t)p->Cout, 9};
    const cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)ctot * 2};
    const cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)p->Cout_store, 9};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tp.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p->weight), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return FBANET_E_BADSHAPE;
  } else {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)p->Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)tp.BN};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tp.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->weight), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return FBANET_E_BADSHAPE;
  }
  int ns = 0;
  if (halo) {
    // chunk table: (source, channel offset); the 9 taps of a chunk re-use its halo boxes
    int nch = 0, cg = 0;
    for (int s = 0; s < p->nsrc; ++s)
      for (int c0 = 0; c0 < p->src[s].C; c0 += TC_BK) {
        Chunk& ch = tp.chunks[nch++];
        ch.src = (int16_t)s; ch.c0 = (int16_t)c0; ch.cg = cg;
        cg += TC_BK;
      }
    tp.nchunks = nch;
    ns = nch * 9;
  } else {
    // K-step table in weight order: k = (tap * Ctot + concat channel); tap-stacked mode: the taps live in N, one step per chunk
    for (int tap = 0; tap < (tapsum ? 1 : taps); ++tap) {
      const int ky = tap / p->KW, kx = tap % p->KW;
      for (int s = 0; s < p->nsrc; ++s) {
        const int cc = s2d ? p->src[s].C / 4 : p->src[s].C;
        for (int c0 = 0; c0 < cc; c0 += TC_BK) {
          KStep& st = tp.steps[ns++];
          st.src = (int16_t)s;
          if (s2d) {
            // input row 2y-1+ky = 2*(y+dy) + ys ; s2d channel block (ys*2+xs)*cc
            const int ry = ky - 1, rx = kx - 1;
            const int ys = ry & 1, xs = rx & 1;
            st.dy = (int16_t)((ry - ys) / 2);
            st.dx = (int16_t)((rx - xs) / 2);
            st.c0 = (int16_t)((ys * 2 + xs) * cc + c0);
          } else {
            st.dy = (int16_t)(tapsum ? 0 : ky - p->pad);
            st.dx = (int16_t)(tapsum ? 0 : kx - p->pad);
            st.c0 = (int16_t)c0;
          }
        }
      }
    }
  }
  tp.nsteps = ns;
  tp.ln_gamma = p->ln_gamma; tp.ln_beta = p->ln_beta; tp.ln_eps = p->ln_eps; tp.store_f16 = p->store_f16;
  {   // FBANET_FC1_GELU_H2=0: the fp32 GELU + one rounding (the first version of the fp16-store epilogue)
    static const char* e = getenv("FBANET_FC1_GELU_H2");
    tp.gelu_h2 = (p->store_f16 && p->act == FBANET_ACT_GELU_TANH && !(e && e[0] == '0')) ? 1 : 0;
  }
  tp.bias = p->bias; tp.alpha = p->alpha; tp.ln_stats = p->ln_stats; tp.residual = reinterpret_cast<const bf16*>(p->residual);
  tp.out = p->out; tp.base = p->base;
  tp.res_img_stride = p->res_img_stride; tp.out_img_stride = p->out_img_stride; tp.base_img_stride = p->base_img_stride;
  tp.act = p->act; tp.store_mode = p->store_mode; tp.res_ld = p->res_ld; tp.out_ld = p->out_ld;

  // TMA-store epilogue for the channels-last store modes: 64-column chunks, one 32-row rectangle {bw, 32/bw} per epilogue warp.
  // Needs a tile width that divides (or is a multiple of) 32 so a warp's 32 accumulator rows are a rectangle of the image.
  const bool rect_ok = tw == 8 || tw == 16 || tw == 32 || tw == 64 || tw == 128;
  const bool store_ok = rect_ok && (p->store_mode == FBANET_STORE_NHWC || (p->store_mode == FBANET_STORE_CONVT2 && (p->Cout / 4) % 64 == 0)) &&
                        tp.BN % 64 == 0 && p->Cout_store == p->Cout;
  static const char* force = getenv("FBANET_TC_TMA_STORE");   // experiment switch: 0 = per-thread 16-byte stores everywhere
  tp.tma_store = (store_ok && !(force && force[0] == '0')) ? 1 : 0;
  if (tp.tma_store) {
    const bool ct = p->store_mode == FBANET_STORE_CONVT2;
    const int Co = ct ? p->Cout / 4 : p->Cout_store;
    const int bw = tw < 32 ? tw : 32, bh = 32 / bw;
    const int64_t sx = ct ? 2 * (int64_t)p->out_ld : p->out_ld;                    // elements between tile-space pixels
    const int64_t sy = ct ? 4 * (int64_t)p->Wo * p->out_ld : (int64_t)p->Wo * p->out_ld;
    for (int qq = 0; qq < (ct ? 4 : 1); ++qq) {
      const int64_t off = ct ? ((int64_t)(qq >> 1) * 2 * p->Wo + (qq & 1)) * p->out_ld : 0;
      const cuuint64_t dims[4] = {(cuuint64_t)Co, (cuuint64_t)p->Wo, (cuuint64_t)p->Ho, (cuuint64_t)p->N};
      const cuuint64_t strides[3] = {(cuuint64_t)sx * 2, (cuuint64_t)sy * 2, (cuuint64_t)p->out_img_stride * 2};
      const cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&tp.omap[qq], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, reinterpret_cast<bf16*>(p->out) + off, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return FBANET_E_BADSHAPE;
    }
    for (int qq = (ct ? 4 : 1); qq < 4; ++qq) tp.omap[qq] = tp.omap[0];
    tp.rmap = tp.omap[0];
    if (p->residual) {
      const cuuint64_t dims[4] = {(cuuint64_t)p->Cout, (cuuint64_t)p->Wo, (cuuint64_t)p->Ho, (cuuint64_t)p->N};
      const cuuint64_t strides[3] = {(cuuint64_t)p->res_ld * 2, (cuuint64_t)p->Wo * p->res_ld * 2, (cuuint64_t)p->res_img_stride * 2};
      const cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&tp.rmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->residual), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return FBANET_E_BADSHAPE;
    }
  }
  // shared-memory plan: A ring + B ring (or all B slabs resident when they fit) + the epilogue staging sub-tiles.
  // The staged TMA-store epilogue is kept only where its 64 KB do not cost the mainloop anything (same weight residency,
  // same B ring, >= 3 A slots): the 3x3 halo convs and the K = 896 fusion GEMM need that memory for operands.
  const int b_bytes = tp.BN * TC_BK * 2;
  tp.a_slot_bytes = halo ? (halo_mode >= 2 ? TC_HALO_WSLOT : TC_HALO_SLOT) : TC_A_BYTES;
  const int a_min = halo ? 2 : 3;
  const int units = halo ? tp.nchunks : ns;
  const bool ln_smem = p->ln_gamma != nullptr;
  // A-stationary mode: with LayerNorm in shared memory every CTA of an N tile would repeat the normalisation of the same A tile
  // (measured a wash at 3 N tiles); where ALL the weights fit beside a whole A tile (+ look-ahead) and the staging buffers, one CTA walks
  // the N tiles of each A tile instead, so the tile is loaded and normalised once.  FBANET_TC_NINNER=0 disables, =all also for plain 1x1 GEMMs.
  tp.n_inner = 1;
  {
    static const char* e = getenv("FBANET_TC_NINNER");
    const bool want = e ? (e[0] == 'a' ? true : (e[0] != '0' && ln_smem)) : ln_smem;
    const int nt_all = tp.n_tiles_n;
    const int64_t need = (int64_t)nt_all * ns * b_bytes + (int64_t)(units + 1) * tp.a_slot_bytes + 4 * 2 * 1 * 4096;
    if (want && !halo && !tapsum && p->KH == 1 && p->stride == 1 && nt_all > 1 && nt_all * tp.BN <= 512 && nt_all * ns <= TC_MAX_STEPS && need <= 216 * 1024 &&
        units + 1 <= TC_MAX_A_SLOTS) {
      tp.n_inner = nt_all;
      tp.n_tiles_n = 1;
    }
  }
  const int nsb = ns * tp.n_inner;       // weight slabs when resident
  struct Plan { int resident, a_slots, b_slots; };
  auto plan = [&](int budget, Plan* pl) -> bool {
    if ((int64_t)nsb * b_bytes + (int64_t)a_min * tp.a_slot_bytes <= budget) {
      pl->resident = 1;
      pl->b_slots = nsb;
      const int a = (budget - nsb * b_bytes) / tp.a_slot_bytes;
      pl->a_slots = a > TC_MAX_A_SLOTS ? TC_MAX_A_SLOTS : a;
    } else if (tp.n_inner > 1) {
      return false;                      // the A-stationary mode needs every slab resident
    } else {
      pl->resident = 0;
      pl->a_slots = halo ? 2 : (ln_smem && units + 2 > 4 ? units + 2 : 4);   // LN in smem: a whole tile (units slots) + look-ahead
      const int b = (budget - pl->a_slots * tp.a_slot_bytes) / b_bytes;
      if (b < 2) return false;
      pl->b_slots = b >= 8 ? 8 : (b >= 4 ? 4 : 2);   // power of two: ring slot = step & (slots-1)
    }
    static const char* ecap = getenv("FBANET_TC_ACAP");   // experiment: 1 = old cap of two tiles of prefetch
    if (ecap && ecap[0] == '1' && pl->a_slots > units * 2 && units * 2 >= 2) pl->a_slots = units * 2;
    return true;
  };
  Plan p0, p1;
  if (!plan(216 * 1024, &p0)) return FBANET_E_UNSUPPORTED;
  int stage_bytes = 0, epi_slots = 2;
  if (tp.tma_store) {
    // staging options in order of preference: (epilogue warps per lane quarter, buffers per warp).  Three warps per quarter want
    // >= 4 A slots left (deep-K GEMMs lose more from a shallow A ring than they gain in the epilogue), two warps >= 3.
    static const int opts[4][2] = {{3, 2}, {3, 1}, {2, 2}, {2, 1}};
    tp.tma_store = 0;
    // A-stationary mode: an A tile's slots stay occupied for n_inner N tiles, so the ring should hold TWO whole tiles (the next one is
    // loaded and normalised meanwhile) even at the price of single staging buffers; second pass: one tile + one slot
    for (int pass = tp.n_inner > 1 ? 0 : 1; pass < 2 && !tp.tma_store; ++pass)
    for (int o = ln_smem ? 2 : 0; o < 4 && !tp.tma_store; ++o) {   // LN in smem: the four LayerNorm warps take the third warp set's place
      const int slots = opts[o][0], bufs = opts[o][1];
      const int sb = 4 * slots * bufs * 4096;
      const int floor_a = slots == 3 ? 4 : 3;
      const int want_a = pass == 0 ? 2 * units : ((ln_smem || tp.n_inner > 1) ? units + 1 : (p0.a_slots < floor_a ? p0.a_slots : floor_a));   // LN in smem / A-stationary: a whole tile + one slot of look-ahead
      if (plan(216 * 1024 - sb, &p1) && p1.resident == p0.resident && p1.b_slots == p0.b_slots && p1.a_slots >= want_a) {
        tp.tma_store = 1; tp.stage_bufs = bufs; stage_bytes = sb; epi_slots = slots; p0 = p1;
      }
    }
  }
  if (tapsum) {   // two P tiles (one per epilogue warp set) of 128 rows x (BN + 1) floats; weights must be resident
    stage_bytes = 2 * 128 * (tp.BN + 1) * 4;
    if (!plan(216 * 1024 - stage_bytes, &p0) || !p0.resident) return FBANET_E_UNSUPPORTED;
    epi_slots = 2;
  }
  tp.b_resident = p0.resident; tp.a_slots = p0.a_slots; tp.b_slots = p0.b_slots;
  if (tp.ln_stats) tp.gelu_h2 = 0;   // the half2 GELU is not in the folded-LayerNorm instantiation
  if ((ln_smem || tp.n_inner > 1) && tp.a_slots < units + 1) return FBANET_E_UNSUPPORTED;   // the LayerNorm warps (the N-tile walk) hold all K chunks of a tile at once
  { const char* dbg = getenv("FBANET_TC_DEBUG"); tp.debug = dbg ? atoi(dbg) : 0; }
  // accumulator stages in TMEM.  Four (N tiles up to 128) were measured against two on every layer shape and change nothing:
  // the gap between "MMA-only" (1.55 ms) + "epilogue-only" (1.28 ms) and both together (1.77 ms) on the 64->64 body conv is not
  // the two-deep accumulator hand-off (tools/prof_bound.py; the chip is power capped, every active unit costs clock).  Kept as
  // an experiment switch: FBANET_TC_NACC=4.
  { const char* na = getenv("FBANET_TC_NACC"); tp.nacc = (4 * tp.BN <= 512 && na && na[0] == '4') ? 4 : 2; tp.nacc_shift = tp.nacc == 4 ? 2 : 1; }
  const size_t smem = (size_t)tp.a_slots * tp.a_slot_bytes + (size_t)tp.b_slots * b_bytes + stage_bytes + 1024;

  static size_t smem_opted_in = 0;  // opt-in limit is per function; raise it only when a launch needs more
  if (smem > smem_opted_in) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<false, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
    smem_opted_in = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // persistent grid: a multiple of n_tiles_n (each CTA owns one N-tile), at most one CTA per SM
  int per_n = sms / tp.n_tiles_n;
  if (per_n < 1) per_n = 1;
  if (per_n > tp.m_tiles) per_n = tp.m_tiles;
  const int grid = per_n * tp.n_tiles_n;
  const bool xtra = tp.gelu_h2 || tp.n_inner > 1;   // (never together with the folded-LayerNorm epilogue: gelu_h2 is cleared, n_inner needs ln_gamma or the env switch)
  if (xtra && tp.ln_stats) return FBANET_E_UNSUPPORTED;
  if (ln_smem) {          // 8 epilogue warps + 4 LayerNorm warps
    if (xtra) conv_gemm_tcgen05_kernel<false, 2, true, true><<<grid, 512, smem, stream>>>(tp);
    else conv_gemm_tcgen05_kernel<false, 2, true><<<grid, 512, smem, stream>>>(tp);   // single N tile (dim 64): the plain LayerNorm instantiation (0.93 vs 1.03 ms per four launches)
  } else if (epi_slots == 3) {   // staged epilogue with 12 epilogue warps
    if (tp.ln_stats) conv_gemm_tcgen05_kernel<true, 3><<<grid, 512, smem, stream>>>(tp);
    else if (xtra) conv_gemm_tcgen05_kernel<false, 3, false, true><<<grid, 512, smem, stream>>>(tp);
    else conv_gemm_tcgen05_kernel<false, 3><<<grid, 512, smem, stream>>>(tp);
  } else {                // 8 epilogue warps (staged or direct stores), 168 registers
    if (tp.ln_stats) conv_gemm_tcgen05_kernel<true, 2><<<grid, 384, smem, stream>>>(tp);
    else if (xtra) conv_gemm_tcgen05_kernel<false, 2, false, true><<<grid, 384, smem, stream>>>(tp);
    else conv_gemm_tcgen05_kernel<false, 2><<<grid, 384, smem, stream>>>(tp);
  }
  return check_launch();
}

}  // namespace fbanet
