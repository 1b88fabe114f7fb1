// The whole LeFF MLP in ONE kernel (bf16 storage, fp32 accumulate), hidden tensor never in HBM:
//
//   out = Linear2( GELU( depthwise3x3( GELU( Linear1(x) ) ) ) ) + residual
//
// (layers/locally_enhanced_feed_forward.py:25-57: linear1 + GELU :27-30, reshape to the H x W map :31-35, depthwise conv 3x3 pad 1 +
// GELU :39-47, flatten, linear2 :56; the residual is layers/fba_net.py:248.)  x = LayerNorm2(x1) tokens [N,H,W,C], hidden Hd = 4C.
//
// Per 8 x 16 pixel tile and per 64-channel chunk of the hidden dimension:
//   fc1 (tcgen05): the tile's (8+2) x (16+2) = 180 HALO pixels x C input channels sit in shared memory (one TMA box per 64-channel
//        K chunk, out-of-image pixels zero-filled) as the A operand of two 128-row MMA tiles; B = 64 rows of W1; D = 2 x 64 fp32 TMEM
//        columns.  The fc1 of the halo pixels is recomputed per tile (1.41x) instead of round-tripping 4C channels through HBM.
//   P1 (CUDA cores): TMEM -> registers -> + bias -> GELU -> bf16 hidden tile in shared memory [180 px][64 ch] (pixels outside the
//        image become 0: the depthwise conv zero-pads the HIDDEN map, not x).  Twelve 32-lane x 32-column items per chunk, bound
//        to the warps of their TMEM lane quarter.
//   P2 (CUDA cores): depthwise 3x3 + bias + GELU from that tile (packed FFMA2 on channel pairs), written as bf16 into the K-major
//        SWIZZLE_128B A tile of fc2.  Sixteen 4 x 2-pixel items per chunk, CLAIMED DYNAMICALLY (shared counter): the four warps
//        whose lane quarter holds no second-tile rows have no P1 item and take more of P2.
//   The hidden tile is double buffered: a warp does P1 of chunk g+1, then P2 items of chunk g, then ONE 512-thread barrier -- the
//   MUFU-bound P1 and the FMA-bound P2 of different warps overlap instead of alternating in lockstep.
//   fc2 (tcgen05): acc[128 px, C] += A2[128, 64] . W2[C, 64-chunk]^T in TMEM; the residual initialises the accumulator through an
//        identity MMA (exact), the epilogue adds bias2 and stores bf16.
// The tensor pipe runs ahead of the CUDA cores: fc1 of chunk g+2 is issued before fc2 of chunk g; both overlap P1 / P2.
//
// Contract: w1, bias1, dw_weight, dw_bias hold HALF the layer's values (a power of two: exact in bf16 / fp32).  With z = x / 2,
//   GELU_tanh(x) = z (1 + tanh(z (2 k0 + 8 k0 k1 z^2))),  GELU_erf(x) = z (1 + erf(sqrt(2) z))  -- one multiply less per element.
//
// Warps (640 threads, persistent, 1 CTA/SM): 0 = x halo + residual TMA, 1 = MMA issuer, 2 = TMEM alloc + W2 TMA, 3 = W1 TMA,
// 4..19 = P1 / P2 / output epilogue.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fbanet {

constexpr int LM_TW = 8, LM_TH = 16;                       // output tile (pixels)
constexpr int LM_HW = LM_TW + 2, LM_HH = LM_TH + 2;        // halo tile 10 x 18
constexpr int LM_HPX = LM_HW * LM_HH;                      // 180 halo pixels
constexpr int LM_XK_BYTES = 256 * 128;                     // one 64-channel K chunk of the x tile: two 128-row MMA tiles
// C = 128: the K chunks are packed 184 rows apart (180 halo rows, rounded up to the 1 KB swizzle atom) instead of 256: the second MMA
// tile of a chunk (rows 128..255) then reads 72 rows of whatever follows -- the next chunk, or the head of the hidden tile -- as its
// accumulator rows 180..255, which nobody looks at.  The 17 KB this frees pay for a second fc2 A tile (P2 of chunk g+1 no longer
// waits for fc2 of chunk g to retire).
constexpr int LM_XK_PACKED = 184 * 128;
constexpr int LM_XBOX_BYTES = LM_HPX * 128;                // bytes TMA writes per K chunk
constexpr int LM_HID_ROW = 136;                           // bf16 hidden tile row: 64 ch x 2 B + 8 B of padding: P1's lanes (one pixel each, 8-byte
                                                           // stores of the same channels) fall into distinct banks; P2's warps read whole rows
constexpr int LM_HID_BUF = 24576;                          // one hidden tile: 180 px x 136 B = 24480, padded to 1 KB
constexpr int LM_HID_BYTES = 2 * LM_HID_BUF;               // double buffered: P1 of chunk g+1 and P2 of chunk g run in the same interval
constexpr int LM_A2_BYTES = 128 * 128;                     // fc2 A tile: 128 px x 64 ch bf16
constexpr int LM_I_BYTES = 64 * 128;                       // 64 x 64 bf16 identity
constexpr int LM_W1K_BYTES = 64 * 128;                     // 64 hidden rows x 64 K
constexpr int LM_CW = 16;                                  // P2 items per chunk (= compute warps of the first version)
constexpr int LM_P1W = 6;                                  // of which P1 (fc1 epilogue) warps; the other 10 run P2

struct LmParams {
  CUtensorMap xmap;    // x [N,H,W,C]: box {64, 10, 18, 1}, SWIZZLE_128B (A operand of fc1)
  CUtensorMap w1map;   // W1 [Hd][C]: box {64, 64}, SWIZZLE_128B
  CUtensorMap w2map;   // W2 [C][Hd]: box {64, C}, SWIZZLE_128B
  CUtensorMap rmap;    // residual [N,H,W,C]: box {64, 8, 16, 1}, SWIZZLE_128B (A operand of the identity MMAs)
  const float* b1;     // [Hd] (halved)
  const float* dw_w;   // [9][Hd] (halved)
  const float* dw_b;   // [Hd] (halved)
  const float* b2;     // [C]
  const bf16* residual;
  bf16* out;
  int64_t out_img_stride;
  int out_ld;
  int N, H, W, C, Hd, act;
  int tiles_x, tiles_y, m_tiles, nchunks, nk;
  int x_slots, w_slots, a2_slots;   // powers of two
  int poly;                         // P1: second MMA tile's GELU on the FMA pipe (polynomial) instead of MUFU.TANH
  int fixed_order;                  // issuer: fc1(g+2), fc2(g) in program order instead of readiness driven
};

// z = x / 2  ->  GELU_tanh(x)
__device__ __forceinline__ f32x2 gelu_half_f2(f32x2 z) {
  const float A = 2.f * 0.7978845608028654f, B = 8.f * 0.7978845608028654f * 0.044715f;
  const f32x2 u = mul_f2(fma_f2(mul_f2(z, z), pack_f2(B, B), pack_f2(A, A)), z);
  float u0, u1, t0, t1;
  unpack_f2(u, u0, u1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  return fma_f2(z, pack_f2(t0, t1), z);
}
__device__ __forceinline__ f32x2 gelu_half_erf_f2(f32x2 z) {
  float z0, z1;
  unpack_f2(z, z0, z1);
  return pack_f2(z0 * (1.f + erff(z0 * 1.4142135623730951f)), z1 * (1.f + erff(z1 * 1.4142135623730951f)));
}
// The same function without the XU pipe: GELU(x) = relu(x) - q(|x|), q(a) = a / (1 + exp(2 k0 (a + k1 a^3))) is a smooth bump
// that is < 7e-5 beyond a = 4; degree-8 polynomial in t = a/2 - 1 on [0,4] (Chebyshev fit, max abs error 1.2e-4: below the
// MUFU.TANH path's own 0.5 |x| 5e-4).  z = x / 2, so a/2 = |z|.
__device__ __forceinline__ f32x2 gelu_half_poly_f2(f32x2 z) {
  float z0, z1;
  unpack_f2(z, z0, z1);
  const float a0 = fminf(fabsf(z0), 2.f) - 1.f, a1 = fminf(fabsf(z1), 2.f) - 1.f;   // t in [-1, 1]
  const f32x2 t = pack_f2(a0, a1);
  // q(2(t+1)) monomial coefficients in t, highest first (tools/fit_gelu_poly.py)
  constexpr float c8 = 1.86211600e-02f, c7 = -8.16851076e-02f, c6 = 2.23242581e-02f, c5 = 2.36755375e-01f, c4 = -3.03877162e-01f,
                  c3 = 1.78552066e-02f, c2 = 2.17632151e-01f, c1 = -1.72978379e-01f, c0 = 4.53658677e-02f;
  f32x2 r = fma_f2(pack_f2(c8, c8), t, pack_f2(c7, c7));
  r = fma_f2(r, t, pack_f2(c6, c6));
  r = fma_f2(r, t, pack_f2(c5, c5));
  r = fma_f2(r, t, pack_f2(c4, c4));
  r = fma_f2(r, t, pack_f2(c3, c3));
  r = fma_f2(r, t, pack_f2(c2, c2));
  r = fma_f2(r, t, pack_f2(c1, c1));
  r = fma_f2(r, t, pack_f2(c0, c0));
  float q0, q1;
  unpack_f2(r, q0, q1);
  return pack_f2(fmaxf(z0 + z0, 0.f) - q0, fmaxf(z1 + z1, 0.f) - q1);
}

// ---- fp16 hidden path (H16): the on-chip hidden tile and the fc2 A tile are fp16 (11 significant bits: finer than the bf16 they replace),
// the depthwise 3x3 and its GELU run on packed HALF2 -- HFMA2 takes the tile's words as they lie (no bf16 -> fp32 unpacking: 48 of a P2
// item's 236 instructions), and one MUFU.TANH.F16x2 serves two values.  fp16 range: |x| < 65504 for the hidden activations (GELU outputs of
// LayerNorm-ed inputs; a value beyond it would have been 3 decimal digits in bf16 as well).
// barrier block: one shared array, addressed as base + constant (a pointer per barrier costs an S2R + LEA chain at every use)
enum : uint32_t {
  LB_X_FULL = 0, LB_X_EMPTY = 2, LB_W1_FULL = 4, LB_W1_EMPTY = 6, LB_W2_FULL = 8, LB_W2_EMPTY = 10, LB_A2_FULL = 12, LB_A2_EMPTY = 14,
  LB_TM1_FULL = 16, LB_TM1_EMPTY = 18, LB_TM2_FULL = 20, LB_TM2_EMPTY = 22, LB_R_FULL = 24, LB_R_EMPTY = 25, LB_DW_FULL = 26, LB_DW_EMPTY = 28,
  LB_HID_FULL = 30, LB_HID_EMPTY = 32, LB_COUNT = 34
};
constexpr int LM_DW_SLOT = 10 * 64 * 4;                    // depthwise taps [9][64] + bias [64] of one hidden chunk, fp32

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// NK = C / 64 (K chunks of fc1 = 64-column slices of the output); POLY: second MMA tile's GELU on the FMA pipe
// CW = compute warps: 16 (96 registers per thread; default) or 24 (72 registers: six P1 + eighteen P2 warps -- measured slower)
// H16: fp16 hidden tile / fc2 A tile with HALF2 depthwise + GELU (tanh flavour only)
template <int NK, bool ERF, bool POLY, int CW, bool H16>
__global__ void __launch_bounds__(128 + 32 * CW, 1) leff_mlp_kernel(const __grid_constant__ LmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_block[LB_COUNT];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias2_s[128];
  __shared__ __align__(16) float bias1_s[512];
  __shared__ __align__(16) float dw_s[2 * 10 * 64];          // two-slot ring of depthwise taps + bias
  __shared__ uint32_t p2_counter;                            // running index of the next unclaimed P2 item (16 per chunk)

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int C = NK * 64;
  const int nchunks = p.nchunks;
  constexpr uint32_t XK = NK == 1 ? (uint32_t)LM_XK_BYTES : (uint32_t)LM_XK_PACKED;   // bytes between the K chunks of an x tile
  constexpr uint32_t x_slot_bytes = (uint32_t)NK * XK;
  constexpr uint32_t w1_slot_bytes = (uint32_t)NK * LM_W1K_BYTES;
  constexpr uint32_t w2_slot_bytes = (uint32_t)C * 128u;
  constexpr uint32_t X_SLOTS = NK == 1 ? 2 : 1, A2_SLOTS = 2, W_SLOTS = 2;
  uint8_t* smem_x = smem;
  uint8_t* smem_hid = smem_x + X_SLOTS * x_slot_bytes;
  uint8_t* smem_w1 = smem_hid + LM_HID_BYTES;
  uint8_t* smem_w2 = smem_w1 + W_SLOTS * w1_slot_bytes;
  uint8_t* smem_a2 = smem_w2 + W_SLOTS * w2_slot_bytes;
  uint8_t* smem_r = smem_a2 + A2_SLOTS * LM_A2_BYTES;
  uint8_t* smem_i = smem_r + LM_A2_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_res = p.residual != nullptr;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t tm1_col0 = 2u * (uint32_t)C;          // [0, 2C): two fc2 accumulators; then two fc1 buffers of 2 x 64 columns
  const uint32_t bars = smem_u32(&bar_block[0]);
#define LBAR(i) (bars + 8u * (uint32_t)(i))

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.xmap); tma_prefetch_desc(&p.w1map); tma_prefetch_desc(&p.w2map);
    if (has_res) tma_prefetch_desc(&p.rmap);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_block[LB_X_FULL + s], 1); mbar_init(&bar_block[LB_X_EMPTY + s], 1);
      mbar_init(&bar_block[LB_W1_FULL + s], 1); mbar_init(&bar_block[LB_W1_EMPTY + s], 1);
      mbar_init(&bar_block[LB_W2_FULL + s], 1); mbar_init(&bar_block[LB_W2_EMPTY + s], 1);
      mbar_init(&bar_block[LB_A2_FULL + s], LM_CW); mbar_init(&bar_block[LB_A2_EMPTY + s], 1);
      mbar_init(&bar_block[LB_TM1_FULL + s], 1); mbar_init(&bar_block[LB_TM1_EMPTY + s], LM_P1W);
      mbar_init(&bar_block[LB_TM2_FULL + s], 1); mbar_init(&bar_block[LB_TM2_EMPTY + s], 4);
      mbar_init(&bar_block[LB_DW_FULL + s], 1); mbar_init(&bar_block[LB_DW_EMPTY + s], CW - LM_P1W);
      mbar_init(&bar_block[LB_HID_FULL + s], LM_P1W); mbar_init(&bar_block[LB_HID_EMPTY + s], CW - LM_P1W);
    }
    mbar_init(&bar_block[LB_R_FULL], 1); mbar_init(&bar_block[LB_R_EMPTY], 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 0) p2_counter = 0;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp >= 4) {
    const int i = threadIdx.x - 128;
    if (i < 128) bias2_s[i] = (p.b2 && i < C) ? __ldg(p.b2 + i) : 0.f;
    if (i < 512) bias1_s[i] = (i < p.Hd) ? __ldg(p.b1 + i) : 0.f;
    if (i < 64) {   // identity row i in the K-major SWIZZLE_128B layout: 16-byte chunk c of row n sits at chunk c ^ (n & 7)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c == (i >> 3)) {
          const uint32_t one = (i & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 in element (i & 7) of the chunk
          const int wsel = (i & 7) >> 1;
          v.x = wsel == 0 ? one : 0u; v.y = wsel == 1 ? one : 0u; v.z = wsel == 2 ? one : 0u; v.w = wsel == 3 ? one : 0u;
        }
        *reinterpret_cast<uint4*>(smem_i + i * 128 + ((c ^ (i & 7)) << 4)) = v;
      }
      fence_proxy_async();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  constexpr uint32_t xmask = X_SLOTS - 1, xshift = X_SLOTS == 2 ? 1 : 0;
  constexpr uint32_t amask = A2_SLOTS - 1, ashift = A2_SLOTS == 2 ? 1 : 0;

  if (warp == 0) {
    // ================= x halo tiles (+ residual slices) =================
    uint32_t it = 0, rg = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++it) {
      const int img = mt / tiles_per_img, r = mt % tiles_per_img;
      const int y0 = (r / p.tiles_x) * LM_TH, x0 = (r % p.tiles_x) * LM_TW;
      const uint32_t xs = it & xmask;
      mbar_wait_a(LBAR(LB_X_EMPTY + xs), ((it >> xshift) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&bar_block[LB_X_FULL + xs], (uint32_t)NK * LM_XBOX_BYTES);
        for (int kc = 0; kc < NK; ++kc)
          tma_load_4d(smem_x + (size_t)xs * x_slot_bytes + (size_t)kc * XK, &p.xmap, &bar_block[LB_X_FULL + xs], kc * 64, x0 - 1, y0 - 1, img);
      }
      __syncwarp();
      if (has_res) {
        for (int kc = 0; kc < NK; ++kc, ++rg) {
          mbar_wait_a(LBAR(LB_R_EMPTY), (rg & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&bar_block[LB_R_FULL], (uint32_t)LM_A2_BYTES);
            tma_load_4d(smem_r, &p.rmap, &bar_block[LB_R_FULL], kc * 64, x0, y0, img);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    // ================= W1 chunks (64 hidden rows x C) + the chunk's depthwise taps / bias =================
    uint32_t g = 0;
    const uint32_t dw0 = smem_u32(dw_s);
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x)
      for (int c = 0; c < nchunks; ++c, ++g) {
        const uint32_t slot = g & 1;
        mbar_wait_a(LBAR(LB_W1_EMPTY + slot), ((g >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&bar_block[LB_W1_FULL + slot], w1_slot_bytes);
          for (int kc = 0; kc < NK; ++kc)
            tma_load_2d(smem_w1 + (size_t)slot * w1_slot_bytes + (size_t)kc * LM_W1K_BYTES, &p.w1map, &bar_block[LB_W1_FULL + slot], kc * 64, c * 64);
        }
        __syncwarp();
        mbar_wait_a(LBAR(LB_DW_EMPTY + slot), ((g >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx_a(LBAR(LB_DW_FULL + slot), (uint32_t)LM_DW_SLOT);
#pragma unroll 1
          for (int t = 0; t < 9; ++t) bulk_load(dw0 + slot * LM_DW_SLOT + (uint32_t)t * 256u, p.dw_w + (size_t)t * p.Hd + c * 64, 256u, LBAR(LB_DW_FULL + slot));
          bulk_load(dw0 + slot * LM_DW_SLOT + 9u * 256u, p.dw_b + c * 64, 256u, LBAR(LB_DW_FULL + slot));
        }
        __syncwarp();
      }
  } else if (warp == 2) {
    // ================= W2 chunks: C rows x 64 hidden channels =================
    uint32_t g = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x)
      for (int c = 0; c < nchunks; ++c, ++g) {
        const uint32_t slot = g & 1;
        mbar_wait_a(LBAR(LB_W2_EMPTY + slot), ((g >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&bar_block[LB_W2_FULL + slot], w2_slot_bytes);
          tma_load_2d(smem_w2 + (size_t)slot * w2_slot_bytes, &p.w2map, &bar_block[LB_W2_FULL + slot], c * 64, 0);
        }
        __syncwarp();
      }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc1 = make_idesc_bf16(64), idesc2 = H16 ? make_idesc_f16(C) : make_idesc_bf16(C);
    const uint64_t desc0 = make_sw128_desc(0);
    const uint32_t sx16 = smem_u32(smem_x) >> 4, sw1_16 = smem_u32(smem_w1) >> 4, sw2_16 = smem_u32(smem_w2) >> 4;
    const uint32_t sa2_16 = smem_u32(smem_a2) >> 4, sr16 = smem_u32(smem_r) >> 4, si16 = smem_u32(smem_i) >> 4;
    int ntiles = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) ++ntiles;
    const uint32_t G = (uint32_t)ntiles * (uint32_t)nchunks;
    uint32_t g1 = 0, it1 = 0, rg = 0;   // fc1 stream: chunk counter, tile counter
    int c1 = 0;
    auto issue_fc1 = [&]() {
      const uint32_t xs = it1 & xmask, b = g1 & 1, ws = g1 & 1;
      if (c1 == 0) mbar_wait_a(LBAR(LB_X_FULL + xs), (it1 >> xshift) & 1);
      mbar_wait_a(LBAR(LB_TM1_EMPTY + b), ((g1 >> 1) & 1) ^ 1);
      mbar_wait_a(LBAR(LB_W1_FULL + ws), (g1 >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t tm = tmem_base + tm1_col0 + b * 128u + (uint32_t)t * 64u;
#pragma unroll
          for (int kc = 0; kc < NK; ++kc) {
            const uint32_t a16 = sx16 + ((xs * x_slot_bytes + (uint32_t)kc * XK + (uint32_t)t * 16384u) >> 4);
            const uint32_t b16 = sw1_16 + ((ws * w1_slot_bytes + (uint32_t)kc * LM_W1K_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tm, desc0 + (uint64_t)(a16 + 2 * k), desc0 + (uint64_t)(b16 + 2 * k), idesc1, (uint32_t)((kc | k) != 0));
          }
        }
        umma_commit_a(LBAR(LB_W1_EMPTY + ws));
        umma_commit_a(LBAR(LB_TM1_FULL + b));
        if (c1 == nchunks - 1) umma_commit_a(LBAR(LB_X_EMPTY + xs));   // the tile's last fc1: its x tile is free once these MMAs retire
      }
      __syncwarp();
      ++g1;
      if (++c1 == nchunks) { c1 = 0; ++it1; }
    };
    uint32_t g2 = 0, it2 = 0;
    int c2 = 0;
    auto issue_fc2 = [&]() {
      const uint32_t acc = it2 & 1, as = g2 & amask, ws = g2 & 1;
      const uint32_t tm = tmem_base + acc * (uint32_t)C;
      if (c2 == 0) {
        mbar_wait_a(LBAR(LB_TM2_EMPTY + acc), ((it2 >> 1) & 1) ^ 1);
        if (has_res) {   // acc[:, 64 kc .. 64 kc + 63] = residual slice . I
          for (int kc = 0; kc < NK; ++kc, ++rg) {
            mbar_wait_a(LBAR(LB_R_FULL), rg & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tm + (uint32_t)(kc * 64), desc0 + (uint64_t)(sr16 + 2 * k), desc0 + (uint64_t)(si16 + 2 * k), idesc1, (uint32_t)(k != 0));
              umma_commit_a(LBAR(LB_R_EMPTY));
            }
            __syncwarp();
          }
        }
      }
      mbar_wait_a(LBAR(LB_A2_FULL + as), (g2 >> ashift) & 1);
      mbar_wait_a(LBAR(LB_W2_FULL + ws), (g2 >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a16 = sa2_16 + ((as * (uint32_t)LM_A2_BYTES) >> 4), b16 = sw2_16 + ((ws * w2_slot_bytes) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tm, desc0 + (uint64_t)(a16 + 2 * k), desc0 + (uint64_t)(b16 + 2 * k), idesc2, (uint32_t)(has_res || (c2 | k) != 0));
        umma_commit_a(LBAR(LB_A2_EMPTY + as));
        umma_commit_a(LBAR(LB_W2_EMPTY + ws));
        if (c2 == nchunks - 1) umma_commit_a(LBAR(LB_TM2_FULL + acc));
      }
      __syncwarp();
      ++g2;
      if (++c2 == nchunks) { c2 = 0; ++it2; }
    };
    // Issue order is READINESS driven (FBANET_LEFF_ORDER=0: the fixed order fc1(g+2), fc2(g) of the first version).  In program order the
    // issuer sat in the wait for P2(g-1)'s A tile before it could issue fc1(g+2), although that needs only P1(g)'s TMEM buffer and the W1
    // slab: P1 then waited for fc1 (11 % of its time, ncu) and P2 for P1 (8 %).  Now whichever of the two next operations has all its
    // barriers complete is issued first, fc1 preferred (it is at most two chunks ahead: two TMEM buffers).
    if (p.fixed_order) {
      if (G > 0) issue_fc1();
      if (G > 1) issue_fc1();
      for (uint32_t g = 0; g < G; ++g) {
        if (g + 2 < G) issue_fc1();
        issue_fc2();
      }
    } else {
      auto ready_fc1 = [&]() -> bool {
        const uint32_t xs = it1 & xmask, b = g1 & 1;
        if (c1 == 0 && !mbar_test_all_a(LBAR(LB_X_FULL + xs), (it1 >> xshift) & 1)) return false;
        return mbar_test_all_a(LBAR(LB_TM1_EMPTY + b), ((g1 >> 1) & 1) ^ 1) && mbar_test_all_a(LBAR(LB_W1_FULL + b), (g1 >> 1) & 1);
      };
      auto ready_fc2 = [&]() -> bool {
        const uint32_t acc = it2 & 1, as = g2 & amask, ws = g2 & 1;
        if (c2 == 0 && !mbar_test_all_a(LBAR(LB_TM2_EMPTY + acc), ((it2 >> 1) & 1) ^ 1)) return false;
        return mbar_test_all_a(LBAR(LB_A2_FULL + as), (g2 >> ashift) & 1) && mbar_test_all_a(LBAR(LB_W2_FULL + ws), (g2 >> 1) & 1);
      };
      while (g2 < G) {
        if (g1 < G && g1 < g2 + 3 && ready_fc1()) issue_fc1();
        else if (ready_fc2()) issue_fc2();
      }
    }
  } else {
    // ================= compute warps, specialised =================
    // P1 warps (6): quarters 0/1 hold rows of both MMA tiles -> two warps each (tile 0, tile 1); quarters 2/3 only tile 0 -> one
    // warp each.  Every P1 warp converts two 32-lane x 32-column items per chunk.  P2 warps (10) claim the chunk's sixteen
    // 4 x 2-pixel items dynamically, so the SM sub-partitions with less P1 work take more of P2.  The two groups meet only through
    // mbarriers on the double-buffered hidden tile (HID_FULL: 6 producers, HID_EMPTY: 10 consumers) and drift up to two chunks apart.
    const int w = warp - 4;
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int slot = w >> 2;
    const bool is_p1 = slot < (q < 2 ? 2 : 1);
    const uint32_t hid0 = smem_u32(smem_hid);
    if (is_p1) {
      const int p1_p = slot * 128 + q * 32 + lane;                 // this thread's halo pixel (MMA tile `slot`)
      const bool p1_valid = p1_p < LM_HPX;
      const uint32_t p1_taddr = tmem_base + ((uint32_t)(q * 32) << 16) + tm1_col0 + (uint32_t)(slot * 64);
      const uint32_t p1_row = hid0 + (uint32_t)p1_p * (uint32_t)LM_HID_ROW;
      uint32_t g = 0;
#pragma unroll 1
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
        bool inimg;
        {
          const int r = mt % tiles_per_img;
          const int hy = p1_p / LM_HW, hx = p1_p - hy * LM_HW;
          const int gy = (r / p.tiles_x) * LM_TH - 1 + hy, gx = (r % p.tiles_x) * LM_TW - 1 + hx;
          inimg = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
        }
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c, ++g) {
          const uint32_t b = g & 1;
          mbar_wait_a(LBAR(LB_TM1_FULL + b), (g >> 1) & 1);
          mbar_wait_a(LBAR(LB_HID_EMPTY + b), ((g >> 1) & 1) ^ 1);   // the P2 warps are done with chunk g-2 (same buffer)
          tc_fence_after();
          const uint32_t row = p1_row + b * (uint32_t)LM_HID_BUF;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {                              // two items: column halves of this warp's 32 rows
            uint32_t v[32];
            tmem_ld32(p1_taddr + b * 128u + (uint32_t)(h * 32), v);
            tmem_ld_wait();
            if (h == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_a(LBAR(LB_TM1_EMPTY + b));   // registers hold the accumulator: fc1 of chunk g+2 may overwrite it
            }
            if (p1_valid) {
              if (inimg) {
                const float* bs = bias1_s + c * 64 + h * 32;
                float4 bn = *reinterpret_cast<const float4*>(bs);
#pragma unroll
                for (int j = 0; j < 8; ++j) {            // 4 channels = one 8-byte store at a time
                  const float4 b4 = bn;
                  if (j < 7) bn = *reinterpret_cast<const float4*>(bs + (j + 1) * 4);
                  f32x2 f0 = add_f2(pack_f2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), pack_f2(b4.x, b4.y));
                  f32x2 f1 = add_f2(pack_f2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack_f2(b4.z, b4.w));
                  if (ERF) { f0 = gelu_half_erf_f2(f0); f1 = gelu_half_erf_f2(f1); }
                  else if (POLY && h == 1) { f0 = gelu_half_poly_f2(f0); f1 = gelu_half_poly_f2(f1); }
                  else { f0 = gelu_half_f2(f0); f1 = gelu_half_f2(f1); }
                  // (H16: P1 keeps its GELU in fp32 -- the half2 form was measured no faster here, 1.406 vs 1.390 ms -- and rounds once, to fp16)
                  if (H16) asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(row + (uint32_t)(h * 64 + j * 8)), "r"(f2_to_f16x2(f0)), "r"(f2_to_f16x2(f1)));
                  else asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(row + (uint32_t)(h * 64 + j * 8)), "r"(f2_to_bf16x2(f0)), "r"(f2_to_bf16x2(f1)));
                }
              } else {                                   // outside the image: the depthwise conv pads the hidden map with zeros
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("st.shared.v2.b32 [%0], {%1,%1};" ::"r"(row + (uint32_t)(h * 64 + j * 8)), "r"(0u));
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_a(LBAR(LB_HID_FULL + b));       // release: this warp's rows of chunk g are in the hidden tile
        }
      }
    } else {
      const uint32_t a2_0 = smem_u32(smem_a2);
      const uint32_t dwl = smem_u32(dw_s) + (uint32_t)lane * 8u;
      const uint32_t ctr_addr = smem_u32(&p2_counter);
      const bool epi = slot == 3;                                    // one P2 warp per lane quarter also stores finished tiles

      auto out_epilogue = [&](int mt, uint32_t it) {                // this quarter's 32 rows x C columns of tile `mt`
        const uint32_t acc = it & 1;
        const int img = mt / tiles_per_img, r = mt % tiles_per_img;
        const int row = q * 32 + lane;
        const int y = (r / p.tiles_x) * LM_TH + row / LM_TW, x = (r % p.tiles_x) * LM_TW + row % LM_TW;
        const bool valid = y < p.H && x < p.W;
        mbar_wait_a(LBAR(LB_TM2_FULL + acc), (it >> 1) & 1);
        tc_fence_after();
        bf16* op = p.out + img * p.out_img_stride + ((int64_t)y * p.W + x) * p.out_ld;
#pragma unroll 1
        for (int cb = 0; cb < C; cb += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)C + (uint32_t)cb, v);
          tmem_ld_wait();
          if (cb + 32 == C) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(LBAR(LB_TM2_EMPTY + acc));
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
              const float4 b0 = *reinterpret_cast<const float4*>(bias2_s + cb + j), b1 = *reinterpret_cast<const float4*>(bias2_s + cb + j + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]) + bb[e];
              store_vec<bf16, 8>(op + cb + j, f);
            }
          }
        }
      };

      uint32_t next;                                                 // the P2 item this warp holds (running index, 16 per chunk)
      {
        uint32_t t = 0;
        if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(t) : "r"(ctr_addr) : "memory");
        next = __shfl_sync(0xffffffffu, t, 0);
      }
      uint32_t g = 0, it = 0;
#pragma unroll 1
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++it) {
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c, ++g) {
          const uint32_t b = g & 1, as = g & amask;
          const uint32_t lim = 16u * (g + 1u);
          if (next < lim) {
            mbar_wait_a(LBAR(LB_HID_FULL + b), (g >> 1) & 1);        // acquire: all six P1 warps have written chunk g
            // the chunk's depthwise taps / bias (shared-memory ring filled by warp 3): this lane's channel pair
            mbar_wait_a(LBAR(LB_DW_FULL + b), (g >> 1) & 1);
            f32x2 wd[9], bd;
#pragma unroll
            for (int t = 0; t < 9; ++t) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(wd[t]) : "r"(dwl + b * (uint32_t)LM_DW_SLOT + (uint32_t)(t * 256)));
            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(bd) : "r"(dwl + b * (uint32_t)LM_DW_SLOT + 9u * 256u));
            uint32_t wh[9], bh = 0, A2 = 0, B2 = 0;          // H16: the taps / bias of this lane's channel pair as half2
            if (H16) {
#pragma unroll
              for (int t = 0; t < 9; ++t) wh[t] = f2_to_f16x2(wd[t]);
              bh = f2_to_f16x2(bd);
              A2 = f2_to_f16x2(pack_f2(2.f * 0.7978845608028654f, 2.f * 0.7978845608028654f));
              B2 = f2_to_f16x2(pack_f2(8.f * 0.7978845608028654f * 0.044715f, 8.f * 0.7978845608028654f * 0.044715f));
            }
            mbar_wait_a(LBAR(LB_A2_EMPTY + as), ((g >> ashift) & 1) ^ 1);   // fc2 of the chunk that used this A tile last has retired
#pragma unroll 1
            while (next < lim) {
              const uint32_t m = next & 15u, cp2 = (m & 3u) * 2u, r4 = (m >> 2) * 4u;   // output columns cp2, cp2+1; rows r4 .. r4+3
              const uint32_t hp = hid0 + b * (uint32_t)LM_HID_BUF + (r4 * LM_HW + cp2) * (uint32_t)LM_HID_ROW + (uint32_t)lane * 4u;
              // A-tile row of output (r4 + o, cp2 + oc) = (r4 + o) * 8 + cp2 + oc: the swizzle phase is row & 7 = cp2 + oc
              const uint32_t ab = a2_0 + as * (uint32_t)LM_A2_BYTES + (r4 * LM_TW + cp2) * 128u + ((uint32_t)lane & 3u) * 4u;
              const uint32_t a0 = ab + ((((uint32_t)lane >> 2) ^ cp2) << 4), a1 = ab + 128u + ((((uint32_t)lane >> 2) ^ (cp2 + 1u)) << 4);
              if (H16) {
                uint32_t hacc[4][2];
#pragma unroll
                for (int o = 0; o < 4; ++o) { hacc[o][0] = bh; hacc[o][1] = bh; }
#pragma unroll
                for (int hr = 0; hr < 6; ++hr) {           // halo rows r4 + hr feed output rows hr - ky
                  uint32_t hv[4];
#pragma unroll
                  for (int hc = 0; hc < 4; ++hc) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hv[hc]) : "r"(hp + (uint32_t)((hr * LM_HW + hc) * LM_HID_ROW)));
#pragma unroll
                  for (int ky = 0; ky < 3; ++ky) {
                    const int o = hr - ky;
                    if (o >= 0 && o < 4) {
#pragma unroll
                      for (int kx = 0; kx < 3; ++kx) {
                        hacc[o][0] = hfma2_(hv[kx], wh[ky * 3 + kx], hacc[o][0]);
                        hacc[o][1] = hfma2_(hv[kx + 1], wh[ky * 3 + kx], hacc[o][1]);
                      }
                    }
                  }
                }
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0 + (uint32_t)(o * LM_TW * 128)), "r"(gelu_half_h2(hacc[o][0], A2, B2)));
                  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1 + (uint32_t)(o * LM_TW * 128)), "r"(gelu_half_h2(hacc[o][1], A2, B2)));
                }
              } else {
              f32x2 acc[4][2];
#pragma unroll
              for (int o = 0; o < 4; ++o) { acc[o][0] = bd; acc[o][1] = bd; }
#pragma unroll
              for (int hr = 0; hr < 6; ++hr) {             // halo rows r4 + hr feed output rows hr - ky
                f32x2 hv[4];
#pragma unroll
                for (int hc = 0; hc < 4; ++hc) {
                  uint32_t u;
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(hp + (uint32_t)((hr * LM_HW + hc) * LM_HID_ROW)));
                  hv[hc] = bf16x2_to_f2(u);
                }
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                  const int o = hr - ky;
                  if (o >= 0 && o < 4) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                      acc[o][0] = fma_f2(hv[kx], wd[ky * 3 + kx], acc[o][0]);
                      acc[o][1] = fma_f2(hv[kx + 1], wd[ky * 3 + kx], acc[o][1]);
                    }
                  }
                }
              }
#pragma unroll
              for (int o = 0; o < 4; ++o) {
                const f32x2 y0 = ERF ? gelu_half_erf_f2(acc[o][0]) : gelu_half_f2(acc[o][0]);
                const f32x2 y1 = ERF ? gelu_half_erf_f2(acc[o][1]) : gelu_half_f2(acc[o][1]);
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0 + (uint32_t)(o * LM_TW * 128)), "r"(f2_to_bf16x2(y0)));
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1 + (uint32_t)(o * LM_TW * 128)), "r"(f2_to_bf16x2(y1)));
              }
              }
              fence_proxy_async();                         // A tile written through the generic proxy -> visible to the tensor core
              __syncwarp();
              uint32_t t = 0;
              if (lane == 0) {
                mbar_arrive_a(LBAR(LB_A2_FULL + as));      // one arrival per item: 16 complete the chunk's A tile
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(t) : "r"(ctr_addr) : "memory");
              }
              next = __shfl_sync(0xffffffffu, t, 0);
            }
          }
          // this warp holds an item of a LATER chunk: every item of chunk g it claimed is done -- release the chunk's taps and hidden buffer
          __syncwarp();
          if (lane == 0) { mbar_arrive_a(LBAR(LB_DW_EMPTY + b)); mbar_arrive_a(LBAR(LB_HID_EMPTY + b)); }
          // a finished tile is stored after its successor's first chunk went through P2 (its last fc2 was issued long before)
          if (epi && c == 0 && it >= 1) out_epilogue(mt - (int)gridDim.x, it - 1);
        }
      }
      if (epi && it >= 1) out_epilogue((int)blockIdx.x + (int)(it - 1) * (int)gridDim.x, it - 1);
    }
  }
#undef LBAR

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

}  // namespace fbanet

using namespace fbanet;

extern "C" int fbanet_leff_mlp_supported(const fbanet_leff_mlp_params* p) {
  if (!p || !p->x || !p->w1 || !p->bias1 || !p->dw_weight || !p->dw_bias || !p->w2 || !p->out) return 0;
  if (p->C != 64 && p->C != 128) return 0;
  if (p->Hd % 64 || p->Hd < 64 || p->Hd > 512 || p->N <= 0 || p->H <= 0 || p->W <= 0) return 0;
  if (((uintptr_t)p->x % 16) || ((uintptr_t)p->w1 % 16) || ((uintptr_t)p->w2 % 16) || ((uintptr_t)p->out % 16) || ((uintptr_t)p->dw_weight % 16) ||
      ((uintptr_t)p->dw_bias % 16) || ((uintptr_t)p->bias1 % 4))
    return 0;
  if ((p->x_ld % 8) || (p->x_img_stride % 8) || (p->out_ld % 8) || (p->out_img_stride % 8)) return 0;
  if (p->residual && (((uintptr_t)p->residual % 16) || (p->res_ld % 8) || (p->res_img_stride % 8))) return 0;
  if (p->act != FBANET_ACT_GELU_TANH && p->act != FBANET_ACT_GELU_ERF) return 0;
  if (p->w2_f16 && p->act != FBANET_ACT_GELU_TANH) return 0;
  return get_encode() != nullptr;
}

extern "C" int fbanet_leff_mlp_sm100(const fbanet_leff_mlp_params* p, void* stream) {
  if (!fbanet_leff_mlp_supported(p)) return FBANET_E_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  static thread_local LmParams lp;
  memset(&lp, 0, sizeof(lp));
  const cuuint32_t estr4[4] = {1, 1, 1, 1};
  {
    const cuuint64_t dims[4] = {(cuuint64_t)p->C, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
    const cuuint64_t strides[3] = {(cuuint64_t)p->x_ld * 2, (cuuint64_t)p->x_ld * 2 * p->W, (cuuint64_t)p->x_img_stride * 2};
    const cuuint32_t box[4] = {64, LM_HW, LM_HH, 1};
    if (encode(&lp.xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->x), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FBANET_E_BADSHAPE;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p->C, (cuuint64_t)p->Hd};
    const cuuint64_t strides[1] = {(cuuint64_t)p->C * 2};
    const cuuint32_t box[2] = {64, 64};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&lp.w1map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->w1), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FBANET_E_BADSHAPE;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p->Hd, (cuuint64_t)p->C};
    const cuuint64_t strides[1] = {(cuuint64_t)p->Hd * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)p->C};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&lp.w2map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->w2), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FBANET_E_BADSHAPE;
  }
  if (p->residual) {
    const cuuint64_t dims[4] = {(cuuint64_t)p->C, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
    const cuuint64_t strides[3] = {(cuuint64_t)p->res_ld * 2, (cuuint64_t)p->res_ld * 2 * p->W, (cuuint64_t)p->res_img_stride * 2};
    const cuuint32_t box[4] = {64, LM_TW, LM_TH, 1};
    if (encode(&lp.rmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->residual), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FBANET_E_BADSHAPE;
  }
  lp.b1 = p->bias1; lp.dw_w = p->dw_weight; lp.dw_b = p->dw_bias; lp.b2 = p->bias2;
  lp.residual = reinterpret_cast<const bf16*>(p->residual); lp.out = reinterpret_cast<bf16*>(p->out);
  lp.out_img_stride = p->out_img_stride; lp.out_ld = p->out_ld;
  lp.N = p->N; lp.H = p->H; lp.W = p->W; lp.C = p->C; lp.Hd = p->Hd; lp.act = p->act;
  lp.tiles_x = (p->W + LM_TW - 1) / LM_TW;
  lp.tiles_y = (p->H + LM_TH - 1) / LM_TH;
  lp.m_tiles = p->N * lp.tiles_x * lp.tiles_y;
  lp.nchunks = p->Hd / 64;
  lp.nk = p->C / 64;
  // shared-memory plan (bytes): x tile(s) | fp32 hidden tile | W1 ring | W2 ring | fc2 A tile(s) | residual slice | identity
  lp.x_slots = p->C == 64 ? 2 : 1;      // C = 128: one 46 KB x tile (the next tile's load waits for the last fc1 of this one)
  lp.w_slots = 2;
  lp.a2_slots = 2;
  { static const char* e = getenv("FBANET_LEFF_POLY"); lp.poly = (e && e[0] == '1') ? 1 : 0; }
  { static const char* e = getenv("FBANET_LEFF_ORDER"); lp.fixed_order = (e && e[0] == '0') ? 1 : 0; }
  const size_t smem = (size_t)lp.x_slots * lp.nk * (p->C == 64 ? LM_XK_BYTES : LM_XK_PACKED) + LM_HID_BYTES + (size_t)lp.w_slots * lp.nk * LM_W1K_BYTES +
                      (size_t)lp.w_slots * p->C * 128 + (size_t)lp.a2_slots * LM_A2_BYTES + LM_A2_BYTES + LM_I_BYTES + 1024;
  typedef void (*KernelFn)(const LmParams);
  const bool erf = p->act == FBANET_ACT_GELU_ERF;
  // experiment switch FBANET_LEFF_WARPS=24: measured slower (dec1 1.579 vs 1.483 ms, profiles/r2_cw_leff_warps_ab.log): the kernel is not
  // occupancy bound, and at 72 registers the compiler has less room to interleave the GELU / FMA chains
  static const int cw = [] { const char* e = getenv("FBANET_LEFF_WARPS"); return (e && atoi(e) == 24) ? 24 : 16; }();
  // w2_f16 (fc2 weights passed as fp16): fp16 hidden tile with HALF2 depthwise arithmetic; else bf16 hidden tile, fp32 (FFMA2) arithmetic
  const bool h16 = p->w2_f16 != 0;
  if (h16 && (erf || lp.poly || cw != 16)) return FBANET_E_UNSUPPORTED;
  KernelFn fn;
  if (h16) fn = p->C == 64 ? leff_mlp_kernel<1, false, false, 16, true> : leff_mlp_kernel<2, false, false, 16, true>;
  else if (cw == 16) {
    if (p->C == 64) fn = erf ? leff_mlp_kernel<1, true, false, 16, false> : (lp.poly ? leff_mlp_kernel<1, false, true, 16, false> : leff_mlp_kernel<1, false, false, 16, false>);
    else fn = erf ? leff_mlp_kernel<2, true, false, 16, false> : (lp.poly ? leff_mlp_kernel<2, false, true, 16, false> : leff_mlp_kernel<2, false, false, 16, false>);
  } else {
    if (p->C == 64) fn = erf ? leff_mlp_kernel<1, true, false, 24, false> : (lp.poly ? leff_mlp_kernel<1, false, true, 24, false> : leff_mlp_kernel<1, false, false, 24, false>);
    else fn = erf ? leff_mlp_kernel<2, true, false, 24, false> : (lp.poly ? leff_mlp_kernel<2, false, true, 24, false> : leff_mlp_kernel<2, false, false, 24, false>);
  }
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = lp.m_tiles < sms ? lp.m_tiles : sms;
  fn<<<grid, 128 + 32 * cw, smem, (cudaStream_t)stream>>>(lp);
  return check_launch();
}
