// K6 on the 5th-gen tensor cores: windowed self-attention for the d_h = 64 stages (first two encoder stages of both hourglasses:
// dim 64 x 1 head @ S^2, dim 128 x 2 heads @ (S/2)^2), window 10 x 10 = 100 tokens.
// (layers/window_attention.py:173-243 with the roll / partition / shift mask of layers/fba_net.py:149-238.)
//
// One item = one (window, head).  Per item:
//   TMA   : the window's q | k | v rows -- a {64 channels, 10 x 10 tokens} box of the [B,H,W,3C] token map per operand (2 or 4 boxes
//           when the cyclically shifted window wraps around the image edge; rows then sit in box order, see ROW ORDER) -- land in
//           shared memory as K-major SWIZZLE_128B tiles, three items deep, so the HBM stream never waits for the math;
//   tcgen05: S[128 x 112] = Q K^T (K = 64: four MMAs) into TMEM;
//   softmax: two threads per query row (the row = their TMEM lane; each owns half of the key columns): S + bias (+ shift mask), row
//           max, exp2, sum in registers, one shared-memory exchange per row for max and sum -- and P as bf16 into a K-major
//           SWIZZLE_128B A tile;
//   tcgen05: O[128 x 64] = P V (K = 112: seven MMAs; V is used as it lies, [key][d] = MN-major B operand) into TMEM;
//   epilogue: the row's thread scales by 1 / sum and stores its token's 64 channels (128 bytes).
// Two sets of four softmax warps take the items alternately (S / O double buffered in TMEM), the MMA warp issues S of item n+1
// before O of item n.  The q columns already carry scale * log2(e) (folded into the q projection at pack time) and the bias table
// is dense and in log2 units (fbanet_attn_params.bias_expanded), so the scores come out of the MMA ready for exp2.
//
// ROW ORDER.  A window that wraps is fetched as 2 (one axis) or 4 (both) boxes; the rows of a box are contiguous, so the tile holds
// the tokens in box order: y-split keeps the natural order, an x-split puts columns 0..4 of every row first.  Attention is
// equivariant under a permutation applied to q, k, v alike; the bias of a wrapping window comes from a table prepared for its wrap
// type (fbanet_attn_params.bias_wrap: rows and columns permuted alike) into which the Swin mask is folded: it is block structure
// in this order -- tokens of different boxes are exactly the tokens of different shift regions (-100 in the reference,
// layers/fba_net.py:151-184).
//
// Warps (640 threads, persistent, 1 CTA/SM, head = blockIdx.x % heads): 0 = TMA, 1 = MMA issuer, 2 = TMEM alloc, 4..11 / 12..19 =
// softmax + epilogue sets (two warps per TMEM lane quarter: a query row is shared by two threads).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fbanet {

constexpr int WA_N = 100, WA_NP = 112, WA_WIN = 10, WA_DH = 64;
// one input slot = q | k | v tiles of 64 bf16 per row.  TMA writes 100 rows of each; the K and V tiles are read to row 111 (N = K = 112:
// rows 100..111 are zeroed once and never written), the Q tile to row 127 (M = 128: its rows 104.. run into the K tile -- rows of S
// nobody reads), so the tiles are 13 / 14 / 14 KB instead of 16 each and four items fit where three did
constexpr int WA_QT = 104 * 128, WA_KT = 112 * 128;
constexpr int WA_SLOT = WA_QT + 2 * WA_KT;          // 41 KB
constexpr int WA_SLOTS = 4;
constexpr int WA_SBUFS = 3;                         // S / P buffers in TMEM, rotating over the items (112 columns each)

constexpr int WA_BIAS_LD = 116;                     // floats per bias row: 16-byte aligned, conflict-free float4 reads across rows
constexpr int WA_BIAS_BYTES = 46592;                // 100 x 116 x 4 = 46400, padded to 512

enum : uint32_t { WB_IN_FULL = 0, WB_IN_EMPTY = 4, WB_S_FULL = 8, WB_O_FULL = 11, WB_O_EMPTY = 13, WB_P_FULL = 15, WB_COUNT = 17 };

struct WaParams {
  CUtensorMap map[4];   // [B,H,W,3C] token map, boxes {64,10,10,1}, {64,10,5,1} (y split), {64,5,10,1} (x split), {64,5,5,1}
  const float* bias;    // [heads][100][112] log2 units
  const float* bias_wrap;  // [heads][3][100][112]: the same table for the wrap types 1..3 (box row order, shift mask folded in)
  bf16* out;
  int out_ld;
  int B, H, W, C, heads, shift;
  int nwx, nwy, total_windows;
};

__device__ __forceinline__ void wa_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// token (iy, ix) of tile row r for wrap type T (0 none, 1 y, 2 x, 3 both)
template <int T>
__host__ __device__ constexpr int wa_token(int r) {
  if (T <= 1) return r;
  if (T == 2) return ((r % 50) / 5) * 10 + (r / 50) * 5 + (r % 50) % 5;
  return ((r / 50) * 5 + (r % 25) / 5) * 10 + ((r / 25) & 1) * 5 + (r % 25) % 5;
}
__device__ __forceinline__ int wa_token_rt(int type, int r) {
  if (type <= 1) return r;
  if (type == 2) return ((r % 50) / 5) * 10 + (r / 50) * 5 + (r % 50) % 5;
  return ((r / 50) * 5 + (r % 25) / 5) * 10 + ((r / 25) & 1) * 5 + (r % 25) % 5;
}

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
__device__ __forceinline__ void wa_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void wa_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (bf16 pairs, one 32-bit column per two K elements, lane = row) is read from TMEM
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float wa_max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__global__ void __launch_bounds__(384, 1) window_attention_tcgen05_kernel(const __grid_constant__ WaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_block[WB_COUNT];
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_in = smem;                                   // 3 slots x (q | k | v)
  float* bias_s = reinterpret_cast<float*>(smem_in + WA_SLOTS * WA_SLOT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.heads;
  const int cta = blockIdx.x / p.heads, ncta = gridDim.x / p.heads;
  constexpr uint32_t TMEM_COLS = 512;                        // S / P buffer i at 112 i (i < 3), O[set] at 336 + 64 set
  const uint32_t bars = smem_u32(&bar_block[0]);
#define WBAR(i) (bars + 8u * (uint32_t)(i))

  if (warp == 0 && lane == 0)
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.map[i]);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WA_SLOTS; ++s) { mbar_init(&bar_block[WB_IN_FULL + s], 1); mbar_init(&bar_block[WB_IN_EMPTY + s], 1); }
    for (int s = 0; s < WA_SBUFS; ++s) mbar_init(&bar_block[WB_S_FULL + s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_block[WB_P_FULL + s], 4);
      mbar_init(&bar_block[WB_O_FULL + s], 1); mbar_init(&bar_block[WB_O_EMPTY + s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // rows 100..127 of every operand tile are never written by TMA: zero them once (finite K / V padding), and stage this head's bias
  for (int i = threadIdx.x; i < WA_SLOTS * 2 * 12 * 8; i += blockDim.x) {
    const int tile = i / (12 * 8), rem = i % (12 * 8);          // tile = slot * 2 + (0: K, 1: V)
    *reinterpret_cast<uint4*>(smem_in + (size_t)(tile >> 1) * WA_SLOT + WA_QT + (size_t)(tile & 1) * WA_KT + (100 + rem / 8) * 128 + (rem % 8) * 16) =
        make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < WA_N * (WA_NP / 4); i += blockDim.x) {
    const int r = i / (WA_NP / 4), c = (i % (WA_NP / 4)) * 4;
    *reinterpret_cast<float4*>(bias_s + r * WA_BIAS_LD + c) = __ldg(reinterpret_cast<const float4*>(p.bias + ((size_t)head * WA_N + r) * WA_NP + c));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int nw_img = p.nwx * p.nwy;
  const int qcol = head * WA_DH, kcol = p.C + head * WA_DH, vcol = 2 * p.C + head * WA_DH;

  // window geometry: image coordinates of the window origin (shifted grid), wrap type
  auto geometry = [&](int wid, int& b, int& y0, int& x0, int& type) {
    b = wid / nw_img;
    const int wl = wid - b * nw_img;
    const int wy = wl / p.nwx, wx = wl - wy * p.nwx;
    y0 = wy * WA_WIN + p.shift; x0 = wx * WA_WIN + p.shift;
    type = (p.shift > 0 && wy == p.nwy - 1 ? 1 : 0) + (p.shift > 0 && wx == p.nwx - 1 ? 2 : 0);
  };

  if (warp == 0) {
    // ================= TMA producer =================
    uint32_t n = 0;
    for (int wid = cta; wid < p.total_windows; wid += ncta, ++n) {
      int b, y0, x0, type;
      geometry(wid, b, y0, x0, type);
      const uint32_t slot = n % WA_SLOTS;
      mbar_wait_a(WBAR(WB_IN_EMPTY + slot), ((n / WA_SLOTS) & 1) ^ 1);
      if (elect_one()) {
        uint64_t* full = &bar_block[WB_IN_FULL + slot];
        uint8_t* base = smem_in + (size_t)slot * WA_SLOT;
        mbar_expect_tx(full, 3u * WA_N * 128u);
        const int cols[3] = {qcol, kcol, vcol};
#pragma unroll 1
        for (int o = 0; o < 3; ++o) {
          uint8_t* t = base + (o == 0 ? 0 : (o == 1 ? WA_QT : WA_QT + WA_KT));
          if (type == 0) tma_load_4d(t, &p.map[0], full, cols[o], x0, y0, b);
          else if (type == 1) {                       // rows y0 .. y0+4 at the bottom edge, then rows 0 .. 4
            tma_load_4d(t, &p.map[1], full, cols[o], x0, y0, b);
            tma_load_4d(t + 50 * 128, &p.map[1], full, cols[o], x0, 0, b);
          } else if (type == 2) {
            tma_load_4d(t, &p.map[2], full, cols[o], x0, y0, b);
            tma_load_4d(t + 50 * 128, &p.map[2], full, cols[o], 0, y0, b);
          } else {
            tma_load_4d(t, &p.map[3], full, cols[o], x0, y0, b);
            tma_load_4d(t + 25 * 128, &p.map[3], full, cols[o], 0, y0, b);
            tma_load_4d(t + 50 * 128, &p.map[3], full, cols[o], x0, 0, b);
            tma_load_4d(t + 75 * 128, &p.map[3], full, cols[o], 0, 0, b);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer: S(n+1) before O(n) =================
    const uint32_t idesc_s = make_idesc_bf16(WA_NP);
    const uint32_t idesc_o = make_idesc_bf16(WA_DH) | (1u << 16);      // B (= V) is MN-major
    const uint64_t desc0 = make_sw128_desc(0);
    const uint32_t sin16 = smem_u32(smem_in) >> 4;
    int nitems = 0;
    for (int wid = cta; wid < p.total_windows; wid += ncta) ++nitems;
    auto issue_s = [&](uint32_t n) {
      // S of item n goes to buffer n % 3.  That buffer last held S / P of item n-3, whose P V MMAs were issued earlier by this thread
      // and execute in order before this one: no barrier is needed to reuse it
      const uint32_t slot = n % WA_SLOTS, sb = n % WA_SBUFS;
      mbar_wait_a(WBAR(WB_IN_FULL + slot), (n / WA_SLOTS) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t q16 = sin16 + ((slot * (uint32_t)WA_SLOT) >> 4), k16 = q16 + ((uint32_t)WA_QT >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base + sb * (uint32_t)WA_NP, desc0 + (uint64_t)(q16 + 2 * kk), desc0 + (uint64_t)(k16 + 2 * kk), idesc_s, (uint32_t)(kk != 0));
        umma_commit_a(WBAR(WB_S_FULL + sb));
      }
      __syncwarp();
    };
    auto issue_o = [&](uint32_t n) {
      const uint32_t slot = n % WA_SLOTS, k = n & 1, sb = n % WA_SBUFS;
      mbar_wait_a(WBAR(WB_P_FULL + k), (n >> 1) & 1);
      mbar_wait_a(WBAR(WB_O_EMPTY + k), ((n >> 1) & 1) ^ 1);     // the set has read O of its previous item
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v16 = sin16 + ((slot * (uint32_t)WA_SLOT + (uint32_t)(WA_QT + WA_KT)) >> 4);
#pragma unroll
        for (int ks = 0; ks < WA_NP / 16; ++ks)     // keys 16 ks .. 16 ks + 15: 8 TMEM columns of P per step; V rows advance by 2 KB
          umma_bf16_ts(tmem_base + (uint32_t)(WA_SBUFS * WA_NP) + k * 64u, tmem_base + sb * (uint32_t)WA_NP + (uint32_t)(ks * 8),
                       desc0 + (uint64_t)(v16 + (uint32_t)(ks * 128)), idesc_o, (uint32_t)(ks != 0));
        umma_commit_a(WBAR(WB_O_FULL + k));
        umma_commit_a(WBAR(WB_IN_EMPTY + slot));
      }
      __syncwarp();
    };
    if (nitems > 0) issue_s(0);
    if (nitems > 1) issue_s(1);
    for (int n = 0; n < nitems; ++n) {
      if (n + 2 < nitems) issue_s((uint32_t)n + 2);
      issue_o((uint32_t)n);
    }
  } else if (warp >= 4) {
    // ================= softmax + epilogue: set k takes items n = k, k+2, ...; one thread per query row =================
    const int k = (warp - 4) >> 2, q = warp & 3;
    const int r = q * 32 + lane;                         // query row = TMEM lane
    const uint32_t bias_a = smem_u32(bias_s);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t o_addr = lane_addr + (uint32_t)(WA_SBUFS * WA_NP) + (uint32_t)k * 64u;
    uint32_t n = (uint32_t)k;
#pragma unroll 1
    for (int wid = cta + k * ncta; wid < p.total_windows; wid += 2 * ncta, n += 2) {
      // ---- S -> registers (packed pairs), then release the lanes for ... nothing yet: P is written over S below
      const uint32_t sb = n % WA_SBUFS;
      const uint32_t s_addr = lane_addr + sb * (uint32_t)WA_NP;      // S (fp32) and, over it, P (bf16 pairs)
      mbar_wait_a(WBAR(WB_S_FULL + sb), (n / WA_SBUFS) & 1);
      tc_fence_after();
      f32x2 s[WA_NP / 2];
      {
        uint32_t su[WA_NP];
        tmem_ld32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&su[0]));
        tmem_ld32(s_addr + 32u, *reinterpret_cast<uint32_t(*)[32]>(&su[32]));
        tmem_ld32(s_addr + 64u, *reinterpret_cast<uint32_t(*)[32]>(&su[64]));
        wa_tmem_ld16(s_addr + 96u, *reinterpret_cast<uint32_t(*)[16]>(&su[96]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < WA_NP / 2; ++j) s[j] = pack_f2(__uint_as_float(su[2 * j]), __uint_as_float(su[2 * j + 1]));
      }
      // ---- bias: this head's table from shared memory; a wrapping window reads ITS table (rows and columns in box order, the
      // shift mask between boxes folded in) from global memory -- same float4 pattern, no per-column index arithmetic
      {
        int b, y0, x0, type;
        geometry(wid, b, y0, x0, type);
        if (type == 0) {
          // (explicit shared-space loads: through the generic pointer these were LD.E.128 with long-scoreboard stalls -- half of
          // the softmax warps' time in ncu)
          const uint32_t brow = bias_a + (uint32_t)((r < WA_N ? r : 0) * WA_BIAS_LD * 4);
#pragma unroll
          for (int j = 0; j < WA_NP; j += 4) {
            float4 bv;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bv.x), "=f"(bv.y), "=f"(bv.z), "=f"(bv.w) : "r"(brow + (uint32_t)(j * 4)));
            s[j / 2] = add_f2(s[j / 2], pack_f2(bv.x, bv.y));
            s[j / 2 + 1] = add_f2(s[j / 2 + 1], pack_f2(bv.z, bv.w));
          }
        } else {
          const float* grow = p.bias_wrap + ((size_t)(head * 3 + type - 1) * WA_N + (r < WA_N ? r : 0)) * WA_NP;
#pragma unroll
          for (int j = 0; j < WA_NP; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(grow + j));
            s[j / 2] = add_f2(s[j / 2], pack_f2(bv.x, bv.y));
            s[j / 2 + 1] = add_f2(s[j / 2 + 1], pack_f2(bv.z, bv.w));
          }
        }
      }
      // ---- row maximum over the valid keys: four independent FMNMX3 chains
      float m;
      {
        float mm[4] = {-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
        for (int j = 0; j < WA_N / 2; ++j) {
          float a0, a1;
          unpack_f2(s[j], a0, a1);
          mm[j & 3] = wa_max3(mm[j & 3], a0, a1);
        }
        m = fmaxf(wa_max3(mm[0], mm[1], mm[2]), mm[3]);
      }
      // ---- exp2, row sum (two accumulators), P as bf16 pairs: 56 words per row
      const f32x2 negm = pack_f2(-m, -m);
      f32x2 ls0 = pack_f2(0.f, 0.f), ls1 = pack_f2(0.f, 0.f);
      uint32_t pb[64];
#pragma unroll
      for (int j = 0; j < WA_NP / 2; ++j) {
        float d0, d1, e0, e1;
        unpack_f2(add_f2(s[j], negm), d0, d1);
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(d0));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(d1));
        const f32x2 e = pack_f2(e0, e1);
        if (j & 1) ls1 = add_f2(ls1, e); else ls0 = add_f2(ls0, e);
        pb[j] = f2_to_bf16x2(e);
      }
      float inv;
      {
        float l0, l1;
        unpack_f2(add_f2(ls0, ls1), l0, l1);
        inv = 1.f / (l0 + l1);
      }
      // ---- P -> tensor memory, over the first 56 columns of this set's S (every lane's S is in registers): the A operand of the
      // P V MMAs is read from TMEM, so P never goes through shared memory
      tmem_st32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&pb[0]));
      wa_tmem_st16(s_addr + 32u, *reinterpret_cast<uint32_t(*)[16]>(&pb[32]));
      wa_tmem_st8(s_addr + 48u, *reinterpret_cast<uint32_t(*)[8]>(&pb[48]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(WBAR(WB_P_FULL + k));
      // ---- O / l -> this token's 64 channels
      bf16* op = nullptr;
      if (r < WA_N) {
        int b, y0, x0, type;
        geometry(wid, b, y0, x0, type);
        const int tok = wa_token_rt(type, r);
        int y = y0 + tok / WA_WIN, x = x0 + tok % WA_WIN;
        if (y >= p.H) y -= p.H;
        if (x >= p.W) x -= p.W;
        op = p.out + ((int64_t)(b * p.H + y) * p.W + x) * p.out_ld + head * WA_DH;
      }
      mbar_wait_a(WBAR(WB_O_FULL + k), (n >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        uint32_t o[32];
        tmem_ld32(o_addr + (uint32_t)h * 32u, o);
        tmem_ld_wait();
        if (h == 1) {                                    // S / P and O of this set are free: the next S MMA of the set may be issued
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(WBAR(WB_O_EMPTY + k));
        }
        if (op) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            float t[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] = __uint_as_float(o[c + e]) * inv;
            store_vec<bf16, 8>(op + h * 32 + c, t);
          }
        }
      }
    }
  }
#undef WBAR

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

int window_attention_tcgen05_supported(const fbanet_attn_params* p) {
  if (p->dtype != FBANET_BF16 || p->win != WA_WIN || p->C != p->heads * WA_DH) return 0;
  if (!p->bias_expanded || !p->q_prescaled) return 0;
  if (p->shift && (!p->bias_wrap || ((uintptr_t)p->bias_wrap % 16))) return 0;
  if ((p->shift != 0 && p->shift != 5) || (p->H % WA_WIN) || (p->W % WA_WIN)) return 0;
  if (p->shift && (p->H < 2 * WA_WIN || p->W < 2 * WA_WIN)) return 0;      // a wrapping window must not also be the first one
  if ((p->qkv_ld % 8) || (p->out_ld % 8) || ((uintptr_t)p->qkv % 16) || ((uintptr_t)p->out % 16) || ((uintptr_t)p->bias_expanded % 16)) return 0;
  static const char* off = getenv("FBANET_ATTN_TCGEN05");                  // experiment switch: 0 = mma.sync kernel for d_h = 64 too
  if (off && off[0] == '0') return 0;
  return get_encode() != nullptr;
}

int window_attention_tcgen05_launch(const fbanet_attn_params* p, cudaStream_t s) {
  if (!window_attention_tcgen05_supported(p)) return FBANET_E_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  static thread_local WaParams wp;
  memset(&wp, 0, sizeof(wp));
  const cuuint64_t dims[4] = {(cuuint64_t)3 * p->C, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->B};
  const cuuint64_t strides[3] = {(cuuint64_t)p->qkv_ld * 2, (cuuint64_t)p->qkv_ld * 2 * p->W, (cuuint64_t)p->qkv_ld * 2 * p->W * p->H};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint32_t boxes[4][4] = {{64, 10, 10, 1}, {64, 10, 5, 1}, {64, 5, 10, 1}, {64, 5, 5, 1}};
  for (int i = 0; i < 4; ++i)
    if (encode(&wp.map[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->qkv), dims, strides, boxes[i], estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FBANET_E_BADSHAPE;
  wp.bias = p->bias_expanded; wp.bias_wrap = p->bias_wrap; wp.out = reinterpret_cast<bf16*>(p->out); wp.out_ld = p->out_ld;
  wp.B = p->B; wp.H = p->H; wp.W = p->W; wp.C = p->C; wp.heads = p->heads; wp.shift = p->shift;
  wp.nwx = p->W / WA_WIN; wp.nwy = p->H / WA_WIN; wp.total_windows = p->B * wp.nwx * wp.nwy;
  const size_t smem = (size_t)WA_SLOTS * WA_SLOT + WA_BIAS_BYTES + 1024;
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
    opted = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_head = sms / p->heads;
  if (per_head > wp.total_windows) per_head = wp.total_windows;
  if (per_head < 1) per_head = 1;
  window_attention_tcgen05_kernel<<<per_head * p->heads, 384, smem, s>>>(wp);
  return check_launch();
}

}  // namespace fbanet
