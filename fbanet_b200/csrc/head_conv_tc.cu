// K0 on the tensor cores (bf16 output): head conv 3x3, C_in (3 or 4) -> 64, straight from the planar fp32 burst.
//
// The op is store bound (64 bf16 channels written per 3-4 floats read), but on the CUDA cores its 27 x 64 FMAs per pixel
// cost 2.75 ms per 64-burst step against a 0.45 ms HBM floor.  Here four producer warps build the im2col rows directly in
// the K-major SWIZZLE_128B layout tcgen05.mma reads -- one thread per pixel gathers its 9*C_in neighbours, splits every
// fp32 sample into hi + lo bf16 halves (so the image is NOT quantised to bf16: K = 2*9*C_in (+pad) columns, the weights repeated
// for the lo half); two producer groups take alternate tiles -- one elected thread issues M=128, N=64 MMAs into double-buffered TMEM accumulators, and eight epilogue
// warps add the bias and stream 32-row x 128-byte sub-tiles out with per-warp TMA stores.  Weights are converted to bf16
// and laid out as the B operand in shared memory once per CTA.
//
// Producers, second form: a group first stages the tile's 10 x 18 HALO of samples in shared memory (one or two pixels per
// thread, the next tile's loads already in flight while this tile's im2col rows are built), then every thread gathers its 9 * C_in
// neighbours from there: 180 * C_in global loads per tile instead of 128 * 9 * C_in.  With `M` given, a halo sample is the
// homography-warped burst pixel (K1 fused: bilinear taps at the fp64 dst->src coordinates, same arithmetic as warp_kernel, base
// frame copied), so the warped burst of homography_alignment.py:46-55 never exists in HBM.
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fbanet {

namespace {

constexpr int HT_TW = 8, HT_TH = 16;         // pixel tile (one swizzle atom per image row of the tile)
constexpr int HT_A_SLOTS = 4;
constexpr int HT_THREADS = 640;              // warps: 0 idle, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 producers (2 groups), 12..19 epilogue

struct HeadTcParams {
  CUtensorMap omap;      // dst [frames,H,W,64] bf16: box {64, 8, 4, 1}, SWIZZLE_128B
  const float* src;      // [frames][C][H][W]
  const float* weight;   // [9C][64] fp32
  const float* bias;     // [64]
  const double* M;       // [frames][3][3] dst->src homographies (WARP instantiation)
  int frames, frames_per_burst, H, W;
  int tiles_x, tiles_y, m_tiles;
};

constexpr int HT_HW = HT_TW + 2, HT_HH = HT_TH + 2, HT_HPX = HT_HW * HT_HH;   // halo tile 10 x 18 = 180 samples per channel
constexpr int HT_HROW = 12;                                                   // floats per halo row in shared memory

template <int CIN, bool WARP>
__global__ void __launch_bounds__(HT_THREADS, 1) head_conv_tc_kernel(const __grid_constant__ HeadTcParams p) {
  constexpr int K1 = 9 * CIN;                 // real taps
  constexpr int K1P = (K1 + 1) & ~1;          // padded to a whole bf16 pair
  constexpr int KTOT = 2 * K1P;               // hi | lo
  constexpr int NK16 = (KTOT + 15) / 16;      // MMA K-steps
  constexpr int NK64 = (KTOT + 63) / 64;      // 128-byte K chunks per row
  constexpr int NCH = (KTOT + 7) / 8;         // 16-byte chunks a producer writes per row
  constexpr int A_SLOT = NK64 * 16384;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[HT_A_SLOTS], a_empty[HT_A_SLOTS], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[64];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                    // HT_A_SLOTS x NK64 x [128 rows][128 B]
  uint8_t* smem_b = smem_a + HT_A_SLOTS * A_SLOT;            // NK64 x [64 rows][128 B]
  uint8_t* smem_stage = smem_b + NK64 * 8192;                // 8 warps x 2 x 4 KB
  float* smem_halo = reinterpret_cast<float*>(smem_stage + 8 * 8192);   // 2 producer groups x [CIN][18][12] fp32
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&p.omap);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HT_A_SLOTS; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = p.bias ? __ldg(p.bias + threadIdx.x) : 0.f;
  // B operand: row n = output channel, K-major, 16-byte chunk c of row n at chunk c ^ (n & 7); columns [0,K1) = weights,
  // [K1P, K1P + K1) = the same weights again (they multiply the lo halves of the samples), the rest zero.
  for (int i = threadIdx.x; i < NK64 * 64 * 64; i += HT_THREADS) {
    const int k = i >> 6, n = i & 63;   // n fastest: the fp32 weight rows [k][0..63] are read coalesced
    const int kk = k < K1P ? k : k - K1P;
    const float w = (k < KTOT && kk < K1) ? __ldg(p.weight + kk * 64 + n) : 0.f;
    const int c64 = k >> 6, kc = k & 63;
    *reinterpret_cast<bf16*>(smem_b + c64 * 8192 + n * 128 + (((kc >> 3) ^ (n & 7)) << 4) + (kc & 7) * 2) = __float2bfloat16_rn(w);
  }
  // A slots start all-zero: the K padding columns are never written afterwards
  for (int i = threadIdx.x; i < HT_A_SLOTS * A_SLOT / 16; i += HT_THREADS) reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc_bf16(64);
    const uint64_t desc_hi = make_sw128_desc(0);
    const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
    int it = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++it) {
      const int acc = it & 1, slot = it & (HT_A_SLOTS - 1);
      mbar_wait(&tmem_empty[acc], (((uint32_t)it >> 1) & 1) ^ 1);
      mbar_wait(&a_full[slot], ((uint32_t)it / HT_A_SLOTS) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 64);
#pragma unroll
        for (int j = 0; j < NK16; ++j) {
          const uint32_t a_lo = (sa0 + slot * A_SLOT + (j >> 2) * 16384) >> 4, b_lo = (sb0 + (j >> 2) * 8192) >> 4;
          umma_bf16(tmem_d, desc_hi + (uint64_t)(a_lo + 2 * (j & 3)), desc_hi + (uint64_t)(b_lo + 2 * (j & 3)), idesc, (uint32_t)(j != 0));
        }
        umma_commit(&a_empty[slot]);
        umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 12) {
    // ================= im2col producers: two groups of 4 warps take alternate tiles; one thread = one pixel row of the A tile =================
    const int grp = (warp - 4) >> 2;
    const int r = (threadIdx.x - 128) & 127;
    const int ly = r / HT_TW, lx = r % HT_TW;
    const int hw = p.H * p.W;
    const uint32_t row_off = (uint32_t)r * 128u;
    const int r7 = r & 7;
    float* halo = smem_halo + grp * (CIN * HT_HH * HT_HROW);
    const uint32_t halo_u = smem_u32(halo);
    constexpr int NT = WARP ? 4 : 1;          // taps per halo sample
    // this thread's halo samples of ONE tile: h = r and r + 128 (< 180).  Two such sets are in flight (the loads of the tile after next are
    // issued as soon as a set has been written to shared memory): with one set the producers sat on the global-load latency at the top
    // of every tile (ncu: 13 % of all samples on the first halo store).
    struct Samp { float tp[2][CIN][NT], wt[2][4]; bool okt[2][4]; };
    Samp sA, sB;
    auto issue = [&](Samp& S, int mt) {        // loads of tile mt's halo samples (not waited for)
      float (&tp)[2][CIN][NT] = S.tp;
      float (&wt)[2][4] = S.wt;
      bool (&okt)[2][4] = S.okt;
      const int f = mt / tiles_per_img, rr = mt % tiles_per_img;
      const int ty0 = (rr / p.tiles_x) * HT_TH - 1, tx0 = (rr % p.tiles_x) * HT_TW - 1;
      const float* sf = p.src + (int64_t)f * CIN * hw;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int h = r + 128 * e;
        const int hy = h / HT_HW, hx = h - hy * HT_HW;
        const int gy = ty0 + hy, gx = tx0 + hx;
        const bool in = h < HT_HPX && (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
        if constexpr (!WARP) {
#pragma unroll
          for (int c = 0; c < CIN; ++c) tp[e][c][0] = in ? __ldg(sf + c * hw + gy * p.W + gx) : 0.f;
        } else {
          bool ident = (f % p.frames_per_burst) == 0;   // base frame: identity (homography_alignment.py:168,179)
          double sx = gx, sy = gy;
          if (in && !ident) {                  // same arithmetic as warp_kernel (bit-identical samples)
            const double* M = p.M + (int64_t)f * 9;
            const double X = gx, Y = gy;
            const double u = fma(__ldg(M + 0), X, fma(__ldg(M + 1), Y, __ldg(M + 2)));
            const double v = fma(__ldg(M + 3), X, fma(__ldg(M + 4), Y, __ldg(M + 5)));
            const double w = fma(__ldg(M + 6), X, fma(__ldg(M + 7), Y, __ldg(M + 8)));
            const double iw = 1.0 / w;
            sx = u * iw; sy = v * iw;
          }
          const double fx = floor(sx), fy = floor(sy);
          const float ax = (float)(sx - fx), ay = (float)(sy - fy);
          const int x0 = (int)fmin(fmax(fx, -2.0), (double)p.W + 1.0);
          const int y0 = (int)fmin(fmax(fy, -2.0), (double)p.H + 1.0);
          const bool okx0 = x0 >= 0 && x0 < p.W, okx1 = x0 + 1 >= 0 && x0 + 1 < p.W;
          const bool oky0 = y0 >= 0 && y0 < p.H, oky1 = y0 + 1 >= 0 && y0 + 1 < p.H;
          okt[e][0] = in && oky0 && okx0; okt[e][1] = in && oky0 && okx1; okt[e][2] = in && oky1 && okx0; okt[e][3] = in && oky1 && okx1;
          wt[e][0] = (1.f - ay) * (1.f - ax); wt[e][1] = (1.f - ay) * ax; wt[e][2] = ay * (1.f - ax); wt[e][3] = ay * ax;
          const float* q = sf + y0 * p.W + x0;
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            tp[e][c][0] = okt[e][0] ? __ldg(q + c * hw) : 0.f;
            tp[e][c][1] = okt[e][1] ? __ldg(q + c * hw + 1) : 0.f;
            tp[e][c][2] = okt[e][2] ? __ldg(q + c * hw + p.W) : 0.f;
            tp[e][c][3] = okt[e][3] ? __ldg(q + c * hw + p.W + 1) : 0.f;
          }
        }
      }
    };
    const int stride = 2 * (int)gridDim.x;                 // this group's tiles: every second one of the CTA
    constexpr int PF = WARP ? 1 : 2;                       // sample sets in flight (the fused-warp variant's set is 4x larger: one)
    auto body = [&](Samp& S, const int mt, const int it) {
      float (&tp)[2][CIN][NT] = S.tp;
      float (&wt)[2][4] = S.wt;
      bool (&okt)[2][4] = S.okt;
      const int slot = it & (HT_A_SLOTS - 1);
      // ---- halo samples -> shared memory ----
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int h = r + 128 * e;
        if (h < HT_HPX) {
          const int hy = h / HT_HW, hx = h - hy * HT_HW;
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            float val;
            if constexpr (!WARP) val = tp[e][c][0];
            else {   // same summation order as warp_kernel
              float acc = 0.f;
              if (okt[e][0]) acc += wt[e][0] * tp[e][c][0];
              if (okt[e][1]) acc += wt[e][1] * tp[e][c][1];
              if (okt[e][2]) acc += wt[e][2] * tp[e][c][2];
              if (okt[e][3]) acc += wt[e][3] * tp[e][c][3];
              val = acc;
            }
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(halo_u + (uint32_t)(((c * HT_HH + hy) * HT_HROW + hx) * 4)), "f"(val));
          }
        }
      }
      named_bar_sync(1 + grp, 128);                          // the group's halo tile is complete
      if (mt + PF * stride < p.m_tiles) issue(S, mt + PF * stride);   // this set's next tile: its loads fly while the tile(s) in between are built
      float v[K1P];
      if (K1P > K1) v[K1P - 1] = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < CIN; ++c)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[t * CIN + c]) : "r"(halo_u + (uint32_t)(((c * HT_HH + ly + t / 3) * HT_HROW + lx + t % 3) * 4)));
      // sample = hi + lo (bf16 each): columns [0,K1P) hi, [K1P, 2 K1P) lo, zero padding to a whole 16-byte chunk
      uint32_t e[NCH * 4];
#pragma unroll
      for (int j = 0; j < NCH * 4; ++j) e[j] = 0u;
#pragma unroll
      for (int j = 0; j < K1P / 2; ++j) {
        const f32x2 vv = pack_f2(v[2 * j], v[2 * j + 1]);
        const uint32_t hi = f2_to_bf16x2(vv);
        e[j] = hi;
        e[K1P / 2 + j] = f2_to_bf16x2(fma_f2(bf16x2_to_f2(hi), pack_f2(-1.f, -1.f), vv));
      }
      mbar_wait(&a_empty[slot], (((uint32_t)it / HT_A_SLOTS) & 1) ^ 1);
      const uint32_t abase = smem_u32(smem_a) + slot * A_SLOT + row_off;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const uint32_t addr = abase + (uint32_t)((c >> 3) * 16384) + (uint32_t)((((c & 7) ^ r7)) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(e[4 * c]), "r"(e[4 * c + 1]), "r"(e[4 * c + 2]), "r"(e[4 * c + 3]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[slot]);
      named_bar_sync(1 + grp, 128);                          // everyone has read the halo tile: it may be overwritten
    };
    int it = grp, mt = blockIdx.x + grp * (int)gridDim.x;
    if (mt < p.m_tiles) issue(sA, mt);
    if (PF == 2) {
      if (mt + stride < p.m_tiles) issue(sB, mt + stride);
      for (; mt < p.m_tiles; mt += 2 * stride, it += 4) {
        body(sA, mt, it);
        if (mt + stride < p.m_tiles) body(sB, mt + stride, it + 2);
      }
    } else {
      for (; mt < p.m_tiles; mt += stride, it += 2) body(sA, mt, it);
    }
  } else if (warp >= 12) {
    // ================= epilogue: warps 12..15 drain accumulator 0 (even tiles), 16..19 accumulator 1 (odd tiles) =================
    const int q = warp & 3, half = (warp - 12) >> 2, ew = warp - 12;
    uint8_t* stage0 = smem_stage + ew * 8192;
    const uint32_t stage_u = smem_u32(stage0);
    const int r7 = lane & 7;
    uint32_t nb = 0;
    int it = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++it) {
      if ((it & 1) != half) continue;
      const int f = mt / tiles_per_img, rr = mt % tiles_per_img;
      const int y0 = (rr / p.tiles_x) * HT_TH + q * 4, x0 = (rr % p.tiles_x) * HT_TW;   // this warp's 8 x 4 pixel rectangle
      const uint32_t boff = (nb & 1u) * 4096u;
      ++nb;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last used this buffer has read it
      __syncwarp();
      mbar_wait(&tmem_full[half], ((uint32_t)it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
      uint32_t v[64];
      tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[half]);   // accumulator is in registers: the next MMA may overwrite it
      const uint32_t row_addr = stage_u + boff + (uint32_t)lane * 128u;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c * 8), b1 = *reinterpret_cast<const float4*>(bias_s + c * 8 + 4);
        const uint32_t o0 = f2_to_bf16x2(add_f2(pack_f2(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1])), pack_f2(b0.x, b0.y)));
        const uint32_t o1 = f2_to_bf16x2(add_f2(pack_f2(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3])), pack_f2(b0.z, b0.w)));
        const uint32_t o2 = f2_to_bf16x2(add_f2(pack_f2(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5])), pack_f2(b1.x, b1.y)));
        const uint32_t o3 = f2_to_bf16x2(add_f2(pack_f2(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7])), pack_f2(b1.z, b1.w)));
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(row_addr + (uint32_t)((c ^ r7) << 4)), "r"(o0), "r"(o1), "r"(o2), "r"(o3));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&p.omap, stage0 + boff, 0, x0, y0, f);
        bulk_commit();
      }
    }
    if (lane == 0) bulk_wait0();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
  }
}

template <int CIN, bool WARP>
int launch_head_tc(const fbanet_head_conv_params* p, cudaStream_t stream) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return FBANET_E_UNSUPPORTED;
  static thread_local HeadTcParams hp;
  memset(&hp, 0, sizeof(hp));
  const cuuint64_t dims[4] = {64, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->frames};
  const cuuint64_t strides[3] = {128, (cuuint64_t)128 * p->W, (cuuint64_t)128 * p->W * p->H};
  const cuuint32_t box[4] = {64, HT_TW, 4, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if (encode(&hp.omap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, p->dst, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return FBANET_E_BADSHAPE;
  hp.src = p->src; hp.weight = p->weight; hp.bias = p->bias; hp.M = p->M;
  hp.frames = p->frames; hp.frames_per_burst = p->frames_per_burst > 0 ? p->frames_per_burst : 1; hp.H = p->H; hp.W = p->W;
  hp.tiles_x = (p->W + HT_TW - 1) / HT_TW;
  hp.tiles_y = (p->H + HT_TH - 1) / HT_TH;
  hp.m_tiles = p->frames * hp.tiles_x * hp.tiles_y;
  constexpr int NK64 = (2 * ((9 * CIN + 1) & ~1) + 63) / 64;
  constexpr size_t smem = (size_t)HT_A_SLOTS * NK64 * 16384 + NK64 * 8192 + 8 * 8192 + 2 * CIN * HT_HH * HT_HROW * 4 + 1024;
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(head_conv_tc_kernel<CIN, WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
    opted = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = hp.m_tiles < sms ? hp.m_tiles : sms;
  head_conv_tc_kernel<CIN, WARP><<<grid, HT_THREADS, smem, stream>>>(hp);
  return check_launch();
}

}  // namespace

// bf16 destination, 64 output channels, 3 or 4 input channels, 16-byte aligned dst
int head_conv_tc_supported(const fbanet_head_conv_params* p) {
  return p->dtype == FBANET_BF16 && p->Cout == 64 && (p->C == 3 || p->C == 4) && ((uintptr_t)p->dst % 16) == 0 && get_encode() != nullptr &&
         (int64_t)p->C * p->H * p->W < ((int64_t)1 << 31) && (!p->M || (p->frames_per_burst > 0 && ((uintptr_t)p->M % 8) == 0));
}

int head_conv_tc_launch(const fbanet_head_conv_params* p, cudaStream_t stream) {
  if (p->C == 3) return p->M ? launch_head_tc<3, true>(p, stream) : launch_head_tc<3, false>(p, stream);
  if (p->C == 4) return p->M ? launch_head_tc<4, true>(p, stream) : launch_head_tc<4, false>(p, stream);
  return FBANET_E_UNSUPPORTED;
}

}  // namespace fbanet
