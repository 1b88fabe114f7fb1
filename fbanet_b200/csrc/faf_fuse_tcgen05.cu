// K2 in ONE pass (bf16, E = 64): Federated Affinity Fusion gate + the K = F*64 1x1 fusion conv + PReLU
// (blocks/federated_affinity_fusion.py:67-128: compute_guided_aligned_features :79-105 and the first layer of fuse_features :121-128).
//
//   s_f = wbar (*) feat_f          (3x3x64 -> 1; wbar = sum over output channels of temporal_attn1: the gate identity)
//   g_f = sigmoid(|s_f - s_0|), f >= 1;  g_0 = 1
//   z   = PReLU( sum_f g_f . (feat_f W_f^T) + b ),   W_f = columns f*64 .. f*64+63 of feature_fusion.0.weight
//
// The gate is a per-pixel scalar per frame, so it commutes with the frame's slice of the 1x1 conv: the tensor core computes the
// UNGATED product D_f = feat_f W_f^T per frame (fp32 in TMEM) and the CUDA cores accumulate acc += g_f * D_f -- the gated
// feature tensor (the largest of the model: F*64 channels) is never formed, in HBM or on chip, and the features are read from HBM
// exactly once: one TMA box per (8 x 16 pixel tile, frame) with a one-pixel halo, consumed by
//   * the score MMA: the nine taps stacked along N (rows 2 tap, 2 tap + 1 = hi / lo bf16 halves of wbar[tap]), M = the tile's 180
//     halo pixels as two 128-row tiles: P[q][tap] = wbar[tap] . feat_f[q]; the convolution is the shifted sum over P, taken by the
//     pixel's own thread from shared memory;
//   * the fusion MMA: A = the tile's interior 8 x 16 pixels of the SAME box (8-row atoms 1280 bytes apart: the 128B swizzle is a
//     function of the shared-memory address, so a 128-byte aligned start is all the descriptor needs), B = W_f resident in smem.
// Two sets of four "pixel warps" (thread = pixel = TMEM lane) take the frames alternately, each with its own TMEM buffers and
// partial sums; set 1 hands its partial sum to set 0 through TMEM at the end of a tile.
//
// Warps (384 threads, persistent, 1 CTA/SM): 0 = TMA (weights once, then the halo ring), 1 = MMA issuer, 2 = TMEM alloc,
// 4..7 = pixel set 0 (even frames, tile epilogue), 8..11 = pixel set 1 (odd frames).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fbanet {

constexpr int FF_TW = 8, FF_TH = 16, FF_HW = 10, FF_HH = 18, FF_HPX = FF_HW * FF_HH;
constexpr int FF_BOX_BYTES = FF_HPX * 128;       // 23040
constexpr int FF_SLOT = 23552;                   // halo ring slot, 1 KB multiple
constexpr int FF_SLOTS = 3;
constexpr int FF_WF_BYTES = 64 * 128;            // one frame's 64 x 64 slice of the fusion weight
constexpr int FF_WS_BYTES = 32 * 128;            // stacked score weights: 32 rows x 64
constexpr int FF_PS_STRIDE = 9;                  // floats per halo pixel in the score scratch
constexpr int FF_PS_BYTES = 13312;               // 180 x 9 x 4 = 6480 per set, two sets, padded
constexpr int FF_MAX_F = 16;

enum : uint32_t { FB_H_FULL = 0, FB_H_EMPTY = 3, FB_W_FULL = 6, FB_P_FULL = 7, FB_P_EMPTY = 9, FB_D_FULL = 11, FB_D_EMPTY = 13, FB_S0 = 15,
                  FB_PART_FULL = 16, FB_PART_EMPTY = 17, FB_COUNT = 18 };

struct FfParams {
  CUtensorMap fmap;    // feat [B*F, H, W, 64]: box {64, 10, 18, 1}, SWIZZLE_128B
  CUtensorMap wmap;    // fusion weight [64][F*64]: box {64, 64}, SWIZZLE_128B
  CUtensorMap smap;    // stacked score weight [32][64]: box {64, 32}, SWIZZLE_128B
  const float* bias;   // [64]
  const float* alpha;  // PReLU slope (scalar)
  float* gate;         // optional [B][F-1][H][W]
  bf16* out;
  int64_t out_img_stride;
  int out_ld;
  int B, F, H, W;
  int tiles_x, tiles_y, m_tiles;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, "
      "%28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(384, 1) faf_fuse_kernel(const __grid_constant__ FfParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_block[FB_COUNT];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[64];
  __shared__ float s0_s[128];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_h = smem;                                       // halo ring
  uint8_t* smem_w = smem_h + FF_SLOTS * FF_SLOT;                // F x 8 KB fusion weight slices (resident)
  uint8_t* smem_ws = smem_w + (size_t)p.F * FF_WF_BYTES;        // stacked score weights
  float* ps_s = reinterpret_cast<float*>(smem_ws + FF_WS_BYTES);  // two score scratch buffers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int F = p.F;
  constexpr uint32_t TMEM_COLS = 512;
  // TMEM columns: P[set] = set*64 + mma_tile*32 (0..127); D[set] = 128 + set*64; PART = 256
  const uint32_t bars = smem_u32(&bar_block[0]);
#define FBAR(i) (bars + 8u * (uint32_t)(i))

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.fmap); tma_prefetch_desc(&p.wmap); tma_prefetch_desc(&p.smap); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < FF_SLOTS; ++s) { mbar_init(&bar_block[FB_H_FULL + s], 1); mbar_init(&bar_block[FB_H_EMPTY + s], 1); }
    mbar_init(&bar_block[FB_W_FULL], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_block[FB_P_FULL + s], 1); mbar_init(&bar_block[FB_P_EMPTY + s], 4);
      mbar_init(&bar_block[FB_D_FULL + s], 1); mbar_init(&bar_block[FB_D_EMPTY + s], 4);
    }
    mbar_init(&bar_block[FB_S0], 4); mbar_init(&bar_block[FB_PART_FULL], 4); mbar_init(&bar_block[FB_PART_EMPTY], 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x >= 128 && threadIdx.x < 192) bias_s[threadIdx.x - 128] = p.bias ? __ldg(p.bias + threadIdx.x - 128) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================= TMA: the weights once, then one halo box per (tile, frame) =================
    if (elect_one()) {
      mbar_expect_tx(&bar_block[FB_W_FULL], (uint32_t)(F * FF_WF_BYTES + FF_WS_BYTES));
      for (int f = 0; f < F; ++f) tma_load_2d(smem_w + (size_t)f * FF_WF_BYTES, &p.wmap, &bar_block[FB_W_FULL], f * 64, 0);
      tma_load_2d(smem_ws, &p.smap, &bar_block[FB_W_FULL], 0, 0);
    }
    __syncwarp();
    uint32_t n = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
      const int img = mt / tiles_per_img, r = mt % tiles_per_img;
      const int y0 = (r / p.tiles_x) * FF_TH, x0 = (r % p.tiles_x) * FF_TW;
      for (int f = 0; f < F; ++f, ++n) {
        const uint32_t hs = n % FF_SLOTS;
        mbar_wait_a(FBAR(FB_H_EMPTY + hs), ((n / FF_SLOTS) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&bar_block[FB_H_FULL + hs], (uint32_t)FF_BOX_BYTES);
          tma_load_4d(smem_h + (size_t)hs * FF_SLOT, &p.fmap, &bar_block[FB_H_FULL + hs], 0, x0 - 1, y0 - 1, img * F + f);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc_s = make_idesc_bf16(32), idesc_f = make_idesc_bf16(64);
    const uint64_t desc0 = make_sw128_desc(0);
    const uint64_t desc_in = (desc0 & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)((FF_HW * 128) >> 4) << 32);   // interior view: 1280 B between 8-row atoms
    const uint32_t sh16 = smem_u32(smem_h) >> 4, sw16 = smem_u32(smem_w) >> 4, sws16 = smem_u32(smem_ws) >> 4;
    mbar_wait_a(FBAR(FB_W_FULL), 0);
    uint32_t n = 0, n0 = 0, n1 = 0;   // frames issued in total / for pixel set 0 / set 1
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
      for (int f = 0; f < F; ++f, ++n) {
        const uint32_t hs = n % FF_SLOTS, k = (uint32_t)f & 1u, j = k ? n1++ : n0++;
        mbar_wait_a(FBAR(FB_H_FULL + hs), (n / FF_SLOTS) & 1);
        mbar_wait_a(FBAR(FB_P_EMPTY + k), (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t h16 = sh16 + ((hs * (uint32_t)FF_SLOT) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(tmem_base + k * 64u + (uint32_t)t * 32u, desc0 + (uint64_t)(h16 + (uint32_t)t * 1024u + 2 * kk), desc0 + (uint64_t)(sws16 + 2 * kk), idesc_s,
                        (uint32_t)(kk != 0));
          umma_commit_a(FBAR(FB_P_FULL + k));
        }
        __syncwarp();
        mbar_wait_a(FBAR(FB_D_EMPTY + k), (j & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a16 = h16 + (uint32_t)(((FF_HW + 1) * 128) >> 4);   // halo pixel (1, 1) = interior pixel (0, 0)
          const uint32_t b16 = sw16 + (((uint32_t)f * (uint32_t)FF_WF_BYTES) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_base + 128u + k * 64u, desc_in + (uint64_t)(a16 + 2 * kk), desc0 + (uint64_t)(b16 + 2 * kk), idesc_f, (uint32_t)(kk != 0));
          umma_commit_a(FBAR(FB_D_FULL + k));
          umma_commit_a(FBAR(FB_H_EMPTY + hs));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ================= pixel warps: set k takes frames f = k, k+2, ... =================
    const int k = (warp - 4) >> 2, q = warp & 3;
    const int pix = q * 32 + lane;                       // this thread's pixel of the 8 x 16 tile = its TMEM lane
    const int py = pix / FF_TW, px = pix % FF_TW;
    const uint32_t ps_a = smem_u32(ps_s) + (uint32_t)(k * FF_HPX * FF_PS_STRIDE * 4);            // this set's score scratch
    const uint32_t ps_g = ps_a + (uint32_t)((py * FF_HW + px) * FF_PS_STRIDE * 4);               // ... at this pixel's tap (0, 0)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float alpha = p.alpha ? __ldg(p.alpha) : 0.f;
    const int bar_id = 1 + 2 * k;
    uint32_t j = 0, it = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++it) {
      const int img = mt / tiles_per_img, r = mt % tiles_per_img;
      const int y = (r / p.tiles_x) * FF_TH + py, x = (r % p.tiles_x) * FF_TW + px;
      const bool valid = y < p.H && x < p.W;
      float acc[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) acc[c] = 0.f;
      float s0 = 0.f;
      for (int f = k; f < F; f += 2, ++j) {
        // ---- scores: P (TMEM) -> hi + lo per tap -> scratch [halo pixel][tap]
        mbar_wait_a(FBAR(FB_P_FULL + k), j & 1);
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld32(lane_addr + (uint32_t)k * 64u, v);                   // MMA tile 0: halo pixels 0..127
          tmem_ld_wait();
          const uint32_t row = ps_a + (uint32_t)(pix * FF_PS_STRIDE * 4);   // (explicit shared-space accesses: generic ones are LD.E / ST.E)
#pragma unroll
          for (int t = 0; t < 9; ++t)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + (uint32_t)(t * 4)), "f"(__uint_as_float(v[2 * t]) + __uint_as_float(v[2 * t + 1])));
          if (q < 2) {                                                    // MMA tile 1: halo pixels 128..179 live in lane quarters 0, 1
            tmem_ld32(lane_addr + (uint32_t)k * 64u + 32u, v);
            tmem_ld_wait();
            if (128 + pix < FF_HPX) {
              const uint32_t row1 = ps_a + (uint32_t)((128 + pix) * FF_PS_STRIDE * 4);
#pragma unroll
              for (int t = 0; t < 9; ++t)
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(row1 + (uint32_t)(t * 4)), "f"(__uint_as_float(v[2 * t]) + __uint_as_float(v[2 * t + 1])));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(FBAR(FB_P_EMPTY + k));
        named_bar_sync(bar_id, 128);                                      // the set's four warps have written the scratch
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float pv;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pv) : "r"(ps_g + (uint32_t)((((t / 3) * FF_HW + t % 3) * FF_PS_STRIDE + t) * 4)));
          s += pv;
        }
        named_bar_sync(bar_id + 1, 128);                                  // ... and read it: the next frame of this set may overwrite
        float g = 1.f;
        if (f == 0) {
          s0 = s;
          s0_s[pix] = s;
          __syncwarp();
          if (lane == 0) mbar_arrive_a(FBAR(FB_S0));                      // (release) set 1 reads the base-frame scores of this tile
        } else {
          if (k == 1 && f == 1) {
            mbar_wait_a(FBAR(FB_S0), it & 1);
            s0 = s0_s[pix];
          }
          const float d = fabsf(s - s0);
          g = __fdividef(1.f, 1.f + __expf(-d));
          if (p.gate && valid) p.gate[(((int64_t)img * (F - 1) + (f - 1)) * p.H + y) * p.W + x] = g;
        }
        // ---- acc += g * D_f
        mbar_wait_a(FBAR(FB_D_FULL + k), j & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld32(lane_addr + 128u + (uint32_t)k * 64u + (uint32_t)h * 32u, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) acc[h * 32 + c] = fmaf(g, __uint_as_float(v[c]), acc[h * 32 + c]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(FBAR(FB_D_EMPTY + k));
      }
      if (k == 1) {
        // ---- hand the odd frames' partial sum to set 0 through TMEM
        mbar_wait_a(FBAR(FB_PART_EMPTY), (it & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(acc[h * 32 + c]);
          tmem_st32(lane_addr + 256u + (uint32_t)h * 32u, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(FBAR(FB_PART_FULL));
      } else {
        // ---- tile epilogue: + set 1's partial sum, + bias, PReLU, bf16 store (one 128-byte pixel row per thread)
        mbar_wait_a(FBAR(FB_PART_FULL), it & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld32(lane_addr + 256u + (uint32_t)h * 32u, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) acc[h * 32 + c] += __uint_as_float(v[c]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(FBAR(FB_PART_EMPTY));
        if (valid) {
          bf16* op = p.out + img * p.out_img_stride + ((int64_t)y * p.W + x) * p.out_ld;
#pragma unroll
          for (int c = 0; c < 64; c += 8) {
            float t[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float u = acc[c + e] + bias_s[c + e];
              t[e] = u > 0.f ? u : alpha * u;
            }
            store_vec<bf16, 8>(op + c, t);
          }
        }
      }
    }
  }
#undef FBAR

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

}  // namespace fbanet

using namespace fbanet;

extern "C" int fbanet_faf_fuse_supported(const fbanet_faf_fuse_params* p) {
  if (!p || !p->feat || !p->score_weight || !p->fuse_weight || !p->out) return 0;
  if (p->C != 64 || p->F < 2 || p->F > 14 || p->B <= 0 || p->H <= 0 || p->W <= 0) return 0;
  if (((uintptr_t)p->feat % 16) || ((uintptr_t)p->score_weight % 16) || ((uintptr_t)p->fuse_weight % 16) || ((uintptr_t)p->out % 16)) return 0;
  if ((p->out_ld % 8) || (p->out_img_stride % 8)) return 0;
  return get_encode() != nullptr;
}

extern "C" int fbanet_faf_fuse_sm100(const fbanet_faf_fuse_params* p, void* stream) {
  if (!fbanet_faf_fuse_supported(p)) return FBANET_E_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  static thread_local FfParams fp;
  memset(&fp, 0, sizeof(fp));
  {
    const cuuint64_t dims[4] = {64, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->B * p->F};
    const cuuint64_t strides[3] = {128, (cuuint64_t)128 * p->W, (cuuint64_t)128 * p->W * p->H};
    const cuuint32_t box[4] = {64, FF_HW, FF_HH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (encode(&fp.fmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->feat), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FBANET_E_BADSHAPE;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p->F * 64, 64};
    const cuuint64_t strides[1] = {(cuuint64_t)p->F * 128};
    const cuuint32_t box[2] = {64, 64};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&fp.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->fuse_weight), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FBANET_E_BADSHAPE;
  }
  {
    const cuuint64_t dims[2] = {64, 32};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 32};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&fp.smap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->score_weight), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FBANET_E_BADSHAPE;
  }
  fp.bias = p->bias; fp.alpha = p->alpha; fp.gate = p->gate; fp.out = reinterpret_cast<bf16*>(p->out);
  fp.out_img_stride = p->out_img_stride; fp.out_ld = p->out_ld;
  fp.B = p->B; fp.F = p->F; fp.H = p->H; fp.W = p->W;
  fp.tiles_x = (p->W + FF_TW - 1) / FF_TW;
  fp.tiles_y = (p->H + FF_TH - 1) / FF_TH;
  fp.m_tiles = p->B * fp.tiles_x * fp.tiles_y;
  const size_t smem = (size_t)FF_SLOTS * FF_SLOT + (size_t)p->F * FF_WF_BYTES + FF_WS_BYTES + FF_PS_BYTES + 1024;
  static size_t opted = 0;
  if (smem > opted) {
    cudaError_t e = cudaFuncSetAttribute(faf_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
    opted = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = fp.m_tiles < sms ? fp.m_tiles : sms;
  faf_fuse_kernel<<<grid, 384, smem, (cudaStream_t)stream>>>(fp);
  return check_launch();
}
