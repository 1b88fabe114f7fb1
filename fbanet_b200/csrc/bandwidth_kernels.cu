// HBM-bound kernels of the BaseModel forward: homography warp (K1), layout conversion, LayerNorm (K8),
// LeFF depthwise 3x3 + GELU (K7), FAF gate (K2a), full-size tile divide/merge (8f-1).
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace fbanet {

// ------------------------------------------------------------------------------------------------
// K1  homography warp, bilinear, BORDER_CONSTANT 0, dst->src matrix (cv2 WARP_INVERSE_MAP).
// One thread per destination pixel; coordinates in fp64 (fp32 cannot hold 1e-5 px at x~1920).
// ------------------------------------------------------------------------------------------------
// CT = compile-time channel count (0: runtime loop).  One fp64 reciprocal per pixel instead of two divisions; the 4*C taps
// are fetched before any arithmetic on them so every thread keeps them all in flight.
template <int CT>
__global__ void __launch_bounds__(256) warp_kernel(const fbanet_warp_params p) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;   // pixel inside the frame; frame = blockIdx.y (no 64-bit divisions)
  if (r >= p.H * p.W) return;
  const int C = CT ? CT : p.C;
  const int y = r / p.W;
  const int x = r - y * p.W;
  const int f = blockIdx.y;
  const int64_t idx = (int64_t)f * p.H * p.W + r;
  const float* s = p.src + (int64_t)f * p.s_frame;
  float* d = p.dst + (int64_t)f * p.d_frame + (int64_t)y * p.d_y + (int64_t)x * p.d_x;
  if (f % p.frames_per_burst == 0) {  // base frame: identity (homography_alignment.py:168,179)
    const float* s0 = s + (int64_t)y * p.s_y + (int64_t)x * p.s_x;
#pragma unroll
    for (int c = 0; c < C; ++c) d[(int64_t)c * p.d_c] = __ldg(s0 + (int64_t)c * p.s_c);
    if (p.coords) { double* co = p.coords + idx * 2; co[0] = x; co[1] = y; }
    return;
  }
  const double* M = p.M + (int64_t)f * 9;
  const double X = x, Y = y;
  const double u = fma(__ldg(M + 0), X, fma(__ldg(M + 1), Y, __ldg(M + 2)));
  const double v = fma(__ldg(M + 3), X, fma(__ldg(M + 4), Y, __ldg(M + 5)));
  const double w = fma(__ldg(M + 6), X, fma(__ldg(M + 7), Y, __ldg(M + 8)));
  double sx, sy;
  if (p.coords) {   // parity / test path: the two correctly rounded quotients
    sx = u / w; sy = v / w;
    double* co = p.coords + idx * 2; co[0] = sx; co[1] = sy;
  } else {          // one reciprocal (0.5 ulp) and two products: relative error < 2^-51, i.e. < 1e-12 px
    const double iw = 1.0 / w;
    sx = u * iw; sy = v * iw;
  }
  const double fx = floor(sx), fy = floor(sy);
  const float ax = (float)(sx - fx), ay = (float)(sy - fy);
  // clamp before the int conversion so wild homographies cannot overflow
  const int x0 = (int)fmin(fmax(fx, -2.0), (double)p.W + 1.0);
  const int y0 = (int)fmin(fmax(fy, -2.0), (double)p.H + 1.0);
  const bool okx0 = x0 >= 0 && x0 < p.W, okx1 = x0 + 1 >= 0 && x0 + 1 < p.W;
  const bool oky0 = y0 >= 0 && y0 < p.H, oky1 = y0 + 1 >= 0 && y0 + 1 < p.H;
  const float w00 = (1.f - ay) * (1.f - ax), w01 = (1.f - ay) * ax, w10 = ay * (1.f - ax), w11 = ay * ax;
  const float* r0 = s + (int64_t)y0 * p.s_y + (int64_t)x0 * p.s_x;
  const float* r1 = r0 + p.s_y;
  if (CT) {
    float t[CT ? CT : 1][4];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const int64_t oc = (int64_t)c * p.s_c;
      t[c][0] = (oky0 && okx0) ? __ldg(r0 + oc) : 0.f;
      t[c][1] = (oky0 && okx1) ? __ldg(r0 + p.s_x + oc) : 0.f;
      t[c][2] = (oky1 && okx0) ? __ldg(r1 + oc) : 0.f;
      t[c][3] = (oky1 && okx1) ? __ldg(r1 + p.s_x + oc) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      // same summation order as the runtime-C loop below (bit-identical results)
      float acc = 0.f;
      if (oky0 && okx0) acc += w00 * t[c][0];
      if (oky0 && okx1) acc += w01 * t[c][1];
      if (oky1 && okx0) acc += w10 * t[c][2];
      if (oky1 && okx1) acc += w11 * t[c][3];
      d[(int64_t)c * p.d_c] = acc;
    }
  } else {
    for (int c = 0; c < C; ++c) {
      const int64_t oc = (int64_t)c * p.s_c;
      float acc = 0.f;
      if (oky0 && okx0) acc += w00 * __ldg(r0 + oc);
      if (oky0 && okx1) acc += w01 * __ldg(r0 + p.s_x + oc);
      if (oky1 && okx0) acc += w10 * __ldg(r1 + oc);
      if (oky1 && okx1) acc += w11 * __ldg(r1 + p.s_x + oc);
      d[(int64_t)c * p.d_c] = acc;
    }
  }
}

// Planar layouts (x stride 1 on both sides, W % 4 == 0, 16-byte aligned rows): one thread = 4 consecutive destination pixels.
// The 16*CT taps of the four pixels are all in flight before the first blend (the one-pixel kernel is latency bound at 1.9
// TB/s), the three homogeneous coordinates advance by one fp64 add per pixel, and every channel row is written as one float4.
template <int CT>
__global__ void __launch_bounds__(128) warp_planar4_kernel(const fbanet_warp_params p) {
  // 32-bit index math throughout (one frame spans < 2^31 elements, checked by the launcher): the first version spent 450 of its 820
  // instructions per thread on 64-bit tap addresses and was issue bound (ncu: 63 % issue-active, 2.7 TB/s).
  const unsigned W4 = (unsigned)p.W >> 2;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;   // < 2^31 (launcher): 32-bit divisions
  if (idx >= (unsigned)p.frames * (unsigned)p.H * W4) return;
  const unsigned rowi = idx / W4;
  const int x = (int)(idx - rowi * W4) * 4;
  const int f = (int)(rowi / (unsigned)p.H);
  const int y = (int)(rowi - (unsigned)f * (unsigned)p.H);
  const float* s = p.src + (int64_t)f * p.s_frame;
  float* d = p.dst + (int64_t)f * p.d_frame;
  asm volatile("" : "+l"(s), "+l"(d));   // keep the frame bases in registers: ptxas otherwise re-derives src + f * s_frame for every tap
  const unsigned sy_ = (unsigned)p.s_y, sc_ = (unsigned)p.s_c, dc_ = (unsigned)p.d_c;
  const unsigned dofs = (unsigned)y * (unsigned)p.d_y + (unsigned)x;
  if (f % p.frames_per_burst == 0) {  // base frame: identity (homography_alignment.py:168,179)
    const unsigned sofs = (unsigned)y * sy_ + (unsigned)x;
#pragma unroll
    for (int c = 0; c < CT; ++c) *reinterpret_cast<float4*>(d + (dofs + c * dc_)) = __ldg(reinterpret_cast<const float4*>(s + (sofs + c * sc_)));
    return;
  }
  const double* M = p.M + (int64_t)f * 9;
  const double m0 = __ldg(M + 0), m3 = __ldg(M + 3), m6 = __ldg(M + 6);
  const double X = x, Y = y;
  double u = fma(m0, X, fma(__ldg(M + 1), Y, __ldg(M + 2)));
  double v = fma(m3, X, fma(__ldg(M + 4), Y, __ldg(M + 5)));
  double w = fma(m6, X, fma(__ldg(M + 7), Y, __ldg(M + 8)));
  // Out-of-image taps: index clamped to the edge pixel, weight set to zero -- every load is unconditional (16 * CT per thread with no
  // predicate to carry: the predicated form spilled its 16 flags through P2R) and a zero weight adds +-0 exactly as the skipped
  // term did, so results stay bit-identical to warp_kernel for finite images.
  float wt[4][4];
  unsigned o00[4], o01[4], o10[4], o11[4];
  // the four reciprocals from ONE division (Montgomery's trick): 1 / (w0 w1 w2 w3), then products -- a few ulp each (< 1e-12 px)
  double iwk[4];
  {
    const double w0 = w, w1 = w + m6, w2 = w1 + m6, w3 = w2 + m6;
    const double p01 = w0 * w1, p23 = w2 * w3;
    const double ip = 1.0 / (p01 * p23);
    const double i01 = ip * p23, i23 = ip * p01;
    iwk[0] = i01 * w1; iwk[1] = i01 * w0; iwk[2] = i23 * w3; iwk[3] = i23 * w2;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double iw = iwk[k];
    const double sx = u * iw, sy = v * iw;
    u += m0; v += m3;
    // floor in the integer domain: cvt.rmi saturates (and maps NaN to 0), so a wild homography cannot overflow; a saturated
    // coordinate has no valid tap and its weights are zeroed
    const int xf = __double2int_rd(sx), yf = __double2int_rd(sy);
    const float ax = (float)(sx - (double)xf), ay = (float)(sy - (double)yf);
    const bool okx0 = (unsigned)xf < (unsigned)p.W, okx1 = (unsigned)xf + 1u < (unsigned)p.W;   // unsigned: INT_MAX + 1 does not wrap into range
    const bool oky0 = (unsigned)yf < (unsigned)p.H, oky1 = (unsigned)yf + 1u < (unsigned)p.H;
    const int x0 = min(max(xf, -1), p.W), y0 = min(max(yf, -1), p.H);
    const float wx0 = okx0 ? 1.f - ax : 0.f, wx1 = okx1 ? ax : 0.f, wy0 = oky0 ? 1.f - ay : 0.f, wy1 = oky1 ? ay : 0.f;
    // (1 - ay)(1 - ax) etc. are the same products as before; a zeroed factor gives the exact zero weight
    wt[k][0] = wy0 * wx0; wt[k][1] = wy0 * wx1; wt[k][2] = wy1 * wx0; wt[k][3] = wy1 * wx1;
    const unsigned xa = (unsigned)min(max(x0, 0), p.W - 1), xb = (unsigned)min(max(x0 + 1, 0), p.W - 1);
    const unsigned ya = (unsigned)min(max(y0, 0), p.H - 1) * sy_, yb = (unsigned)min(max(y0 + 1, 0), p.H - 1) * sy_;
    o00[k] = ya + xa; o01[k] = ya + xb; o10[k] = yb + xa; o11[k] = yb + xb;
  }
  float t[CT][4][4];
#pragma unroll
  for (int c = 0; c < CT; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned oc = c * sc_;
      t[c][k][0] = __ldg(s + (o00[k] + oc));
      t[c][k][1] = __ldg(s + (o01[k] + oc));
      t[c][k][2] = __ldg(s + (o10[k] + oc));
      t[c][k][3] = __ldg(s + (o11[k] + oc));
    }
#pragma unroll
  for (int c = 0; c < CT; ++c) {
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // same summation order as warp_kernel
      float acc = 0.f;
      acc += wt[k][0] * t[c][k][0];
      acc += wt[k][1] * t[c][k][1];
      acc += wt[k][2] * t[c][k][2];
      acc += wt[k][3] * t[c][k][3];
      o[k] = acc;
    }
    *reinterpret_cast<float4*>(d + (dofs + c * dc_)) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Planar layouts, source STAGED IN SHARED MEMORY: one CTA = one 32 x 16 tile of destination pixels.  The projective image of the
// tile is a convex quadrilateral spanned by the images of its corners (w keeps one sign on the tile, checked), so the taps of all
// its pixels lie in the corners' bounding box: that box (+ a safety pixel each side, x origin rounded down to a multiple of 4) is
// copied once with coalesced 16-byte loads, zero-filled outside the image (BORDER_CONSTANT 0), and every pixel takes its four taps
// per channel from shared memory at IMMEDIATE offsets (+1, +pitch, channel planes): one address per pixel instead of one 64-bit
// address per tap, no validity flags, no clamps.  warp_planar4_kernel spends ~205 instructions per pixel, most of them on the 16 * CT
// global tap addresses, and is bound by instruction issue (3.5 TB/s); this form needs ~90.  A lane owns one tile column, four rows:
// consecutive lanes read consecutive shared-memory words (no bank conflicts) and write full 128-byte lines.
// A tile whose box does not fit (strong zoom / rotation, w changing sign) gathers from global memory as before.
// WT_TW = 32 (tile 32 x 16) or 16 (tile 16 x 32, two tile rows per warp step: frames whose width wastes less that way, e.g. W = 80)
template <int CT, int WT_TW>
__global__ void __launch_bounds__(128) warp_tile_kernel(const fbanet_warp_params p, const int tiles_x, const int tiles_y) {
  constexpr int WT_TH = 512 / WT_TW;                       // destination tile rows
  constexpr int WT_BW = WT_TW + 16, WT_BH = WT_TH + 12;    // staged box (floats per row = pitch, rows)
  constexpr int RPW = 32 / WT_TW;                          // tile rows one warp covers per step (1 or 2)
  __shared__ __align__(16) float box[CT][WT_BH][WT_BW];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const unsigned tpf = (unsigned)tiles_x * (unsigned)tiles_y;
  const unsigned f = blockIdx.x / tpf, tr = blockIdx.x - f * tpf;
  const int ty = (int)(tr / (unsigned)tiles_x), tx = (int)(tr - (unsigned)ty * (unsigned)tiles_x);
  const int x0 = tx * WT_TW, y0 = ty * WT_TH;
  const float* s = p.src + (int64_t)f * p.s_frame;
  float* d = p.dst + (int64_t)f * p.d_frame;
  const unsigned sy_ = (unsigned)p.s_y, sc_ = (unsigned)p.s_c, dc_ = (unsigned)p.d_c, dy_ = (unsigned)p.d_y;
  const int x = x0 + (lane & (WT_TW - 1));                  // this lane's column; rows yb + RPW * k
  const int yb = y0 + 4 * RPW * wrp + lane / WT_TW;
  if (f % (unsigned)p.frames_per_burst == 0) {  // base frame: identity (homography_alignment.py:168,179)
    if (x < p.W) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int y = yb + RPW * k;
        if (y >= p.H) break;
#pragma unroll
        for (int c = 0; c < CT; ++c) d[(unsigned)y * dy_ + (unsigned)x + c * dc_] = __ldg(s + ((unsigned)y * sy_ + (unsigned)x + c * sc_));
      }
    }
    return;
  }
  const double* M = p.M + (int64_t)f * 9;
  const double m0 = __ldg(M + 0), m1 = __ldg(M + 1), m2 = __ldg(M + 2), m3 = __ldg(M + 3), m4 = __ldg(M + 4), m5 = __ldg(M + 5);
  const double m6 = __ldg(M + 6), m7 = __ldg(M + 7), m8 = __ldg(M + 8);
  // ---- the staged box, derived by EVERY warp from the tile's four corners (lanes 0..3; exact quotients): no barrier, no broadcast ----
  int minx4, miny, ncol4, nrow;
  bool staged;
  {
    const int cx = (lane & 1) ? min(x0 + WT_TW, p.W) - 1 : x0, cy = (lane & 2) ? min(y0 + WT_TH, p.H) - 1 : y0;
    const double X = cx, Y = cy;
    const double w = fma(m6, X, fma(m7, Y, m8));
    const double iw = 1.0 / w;
    double sx = fma(m0, X, fma(m1, Y, m2)) * iw, sy = fma(m3, X, fma(m4, Y, m5)) * iw;
    const bool fin = fabs(sx) < 1e9 && fabs(sy) < 1e9 && fabs(w) > 1e-300;     // NaN / huge -> not staged
    const unsigned finm = __ballot_sync(0xffffffffu, fin) & 0xfu, pos = __ballot_sync(0xffffffffu, w > 0.0) & 0xfu;
    if (!fin) { sx = 0.0; sy = 0.0; }
    int lox = __double2int_rd(sx), loy = __double2int_rd(sy), hix = lox, hiy = loy;
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      lox = min(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = max(hix, __shfl_xor_sync(0xffffffffu, hix, o));
      loy = min(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = max(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
    }
    lox = __shfl_sync(0xffffffffu, lox, 0); hix = __shfl_sync(0xffffffffu, hix, 0);
    loy = __shfl_sync(0xffffffffu, loy, 0); hiy = __shfl_sync(0xffffffffu, hiy, 0);
    lox -= 1; loy -= 1; hix += 2; hiy += 2;                 // the second tap, and a pixel of slack for the last-ulp differences of the shared division
    minx4 = lox & ~3;                                       // (two's complement: rounds towards -inf)
    miny = loy;
    ncol4 = ((hix - minx4) >> 2) + 1; nrow = hiy - loy + 1;
    staged = finm == 0xfu && (pos == 0xfu || pos == 0u) && ncol4 * 4 <= WT_BW && nrow <= WT_BH;
  }
  if (staged) {
    // 8 rows x 16 vector slots per pass (a box row has at most WT_BW / 4 <= 12 vectors); asynchronous 16-byte copies, zero-filled outside the image
    const int j = tid & 15, gx = minx4 + 4 * j;
    const bool inx = (unsigned)gx < (unsigned)p.W;        // W % 4 == 0 and gx % 4 == 0: a vector is inside or outside as a whole
    if (j < ncol4) {
      const unsigned sox = (unsigned)min(max(gx, 0), p.W - 4);
#pragma unroll 1
      for (int r = tid >> 4; r < nrow; r += 8) {
        const int gy = miny + r;
        const unsigned nbytes = (inx && (unsigned)gy < (unsigned)p.H) ? 16u : 0u;
        const unsigned so = (unsigned)min(max(gy, 0), p.H - 1) * sy_ + sox;
        const uint32_t dsts = (uint32_t)__cvta_generic_to_shared(&box[0][r][4 * j]);
#pragma unroll
        for (int c = 0; c < CT; ++c)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dsts + (uint32_t)(c * WT_BH * WT_BW * 4)), "l"(s + (so + (unsigned)c * sc_)), "r"(nbytes)
                       : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const bool live = x < p.W && yb < p.H;
  // this lane's four pixels (x, yb + k), computed while the copies are in flight: homogeneous coordinates advance by one fp64 add per
  // row, one division for the four (Montgomery's trick)
  const double X = min(x, p.W - 1), Y = min(yb, p.H - 1);
  const double du = RPW * m1, dv = RPW * m4, dw = RPW * m7;   // (exact: RPW is 1 or 2)
  double u = fma(m0, X, fma(m1, Y, m2)), v = fma(m3, X, fma(m4, Y, m5));
  double iwk[4];
  {
    const double w0 = fma(m6, X, fma(m7, Y, m8)), w1 = w0 + dw, w2 = w1 + dw, w3 = w2 + dw;
    const double p01 = w0 * w1, p23 = w2 * w3;
    const double ip = 1.0 / (p01 * p23);
    const double i01 = ip * p23, i23 = ip * p01;
    iwk[0] = i01 * w1; iwk[1] = i01 * w0; iwk[2] = i23 * w3; iwk[3] = i23 * w2;
  }
  if (staged) {
    const uint32_t box0 = (uint32_t)__cvta_generic_to_shared(&box[0][0][0]);
    uint32_t a[4];
    float wq[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double sx = u * iwk[k], sy = v * iwk[k];
      u += du; v += dv;
      const int xf = __double2int_rd(sx), yf = __double2int_rd(sy);
      const float ax = (float)(sx - (double)xf), ay = (float)(sy - (double)yf);
      const float wx0 = 1.f - ax, wy0 = 1.f - ay;
      wq[k][0] = wy0 * wx0; wq[k][1] = wy0 * ax; wq[k][2] = ay * wx0; wq[k][3] = ay * ax;
      // inside the box by construction; the clamp only keeps a pathological matrix from reading outside the array
      const int lx = min(max(xf - minx4, 0), WT_BW - 2), ly = min(max(yf - miny, 0), WT_BH - 2);
      a[k] = box0 + (uint32_t)(ly * WT_BW + lx) * 4u;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (!live) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (yb + RPW * k >= p.H) break;
      const unsigned dofs = (unsigned)(yb + RPW * k) * dy_ + (unsigned)x;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        float t00, t01, t10, t11;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t00) : "r"(a[k] + (uint32_t)(c * WT_BH * WT_BW * 4)));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t01) : "r"(a[k] + (uint32_t)(c * WT_BH * WT_BW * 4 + 4)));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t10) : "r"(a[k] + (uint32_t)(c * WT_BH * WT_BW * 4 + WT_BW * 4)));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t11) : "r"(a[k] + (uint32_t)(c * WT_BH * WT_BW * 4 + WT_BW * 4 + 4)));
        float acc = 0.f;                         // same summation order as warp_kernel (an outside tap is a staged zero)
        acc += wq[k][0] * t00;
        acc += wq[k][1] * t01;
        acc += wq[k][2] * t10;
        acc += wq[k][3] * t11;
        d[dofs + c * dc_] = acc;
      }
    }
  } else {
    if (!live) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double sx = u * iwk[k], sy = v * iwk[k];
      u += du; v += dv;
      if (yb + RPW * k >= p.H) break;
      const int xf = __double2int_rd(sx), yf = __double2int_rd(sy);
      const float ax = (float)(sx - (double)xf), ay = (float)(sy - (double)yf);
      const bool okx0 = (unsigned)xf < (unsigned)p.W, okx1 = (unsigned)xf + 1u < (unsigned)p.W;
      const bool oky0 = (unsigned)yf < (unsigned)p.H, oky1 = (unsigned)yf + 1u < (unsigned)p.H;
      const int xc = min(max(xf, -1), p.W), yc = min(max(yf, -1), p.H);
      const float wx0 = okx0 ? 1.f - ax : 0.f, wx1 = okx1 ? ax : 0.f, wy0 = oky0 ? 1.f - ay : 0.f, wy1 = oky1 ? ay : 0.f;
      const unsigned xa = (unsigned)min(max(xc, 0), p.W - 1), xb = (unsigned)min(max(xc + 1, 0), p.W - 1);
      const unsigned ya = (unsigned)min(max(yc, 0), p.H - 1) * sy_, yb2 = (unsigned)min(max(yc + 1, 0), p.H - 1) * sy_;
      const unsigned dofs = (unsigned)(yb + RPW * k) * dy_ + (unsigned)x;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const unsigned oc = c * sc_;
        float acc = 0.f;
        acc += (wy0 * wx0) * __ldg(s + (ya + xa + oc));
        acc += (wy0 * wx1) * __ldg(s + (ya + xb + oc));
        acc += (wy1 * wx0) * __ldg(s + (yb2 + xa + oc));
        acc += (wy1 * wx1) * __ldg(s + (yb2 + xb + oc));
        d[dofs + c * dc_] = acc;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// planar fp32 [frames][C][H][W] -> channels-last [frames][H][W][Cp]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) to_nhwc_kernel(const fbanet_to_nhwc_params p) {
  const int64_t total = (int64_t)p.frames * p.H * p.W;
  T* dst = reinterpret_cast<T*>(p.dst);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t hw = (int64_t)p.H * p.W;
    const int64_t f = idx / hw, pix = idx % hw;
    const float* s = p.src + f * p.C * hw + pix;
    T* d = dst + idx * p.Cp;
    if (!p.im2col3x3) {
      for (int c = 0; c < p.Cp; ++c) d[c] = from_f32<T>(c < p.C ? __ldg(s + (int64_t)c * hw) : 0.f);
    } else {
      const int y = (int)(pix / p.W), x = (int)(pix % p.W);
      for (int c = 0; c < p.Cp; ++c) {
        float v = 0.f;
        if (c < 9 * p.C) {
          const int tap = c / p.C, ch = c - tap * p.C;
          const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
          if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) v = __ldg(p.src + (f * p.C + ch) * hw + (int64_t)yy * p.W + xx);
        }
        d[c] = from_f32<T>(v);
      }
    }
  }
}

// head-conv im2col, bf16: 8 threads per pixel, each builds one 16-byte vector (8 of the Cp channels)
__global__ void __launch_bounds__(256) im2col3x3_bf16_kernel(const fbanet_to_nhwc_params p) {
  const int vpp = p.Cp / 8;  // vectors per pixel
  const int64_t hw = (int64_t)p.H * p.W;
  const int64_t total = (int64_t)p.frames * hw * vpp;
  bf16* dst = reinterpret_cast<bf16*>(p.dst);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % vpp);
    const int64_t gp = idx / vpp;
    const int64_t f = gp / hw, pix = gp % hw;
    const int y = (int)(pix / p.W), x = (int)(pix % p.W);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      float val = 0.f;
      if (c < 9 * p.C) {
        const int tap = c / p.C, ch = c - tap * p.C;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) val = __ldg(p.src + (f * p.C + ch) * hw + (int64_t)yy * p.W + xx);
      }
      o[j] = val;
    }
    store_vec<bf16, 8>(dst + idx * 8, o);
  }
}

// Head conv: one thread = two horizontally adjacent pixels x 64 output channels.  The 3 x 4 x C input
// patch sits in registers, the [9C][64] weights in shared memory (read as broadcast float4).
template <typename T, int CIN>
__global__ void __launch_bounds__(128) head_conv_kernel(const fbanet_head_conv_params p) {
  constexpr int CO = 64, K = 9 * CIN;
  __shared__ __align__(16) float ws[K * CO];
  __shared__ __align__(16) float bs[CO];
  for (int i = threadIdx.x; i < K * CO; i += blockDim.x) ws[i] = __ldg(p.weight + i);
  if (threadIdx.x < CO) bs[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  __syncthreads();
  const int wp = (p.W + 1) / 2;  // pixel pairs per row
  const int64_t hw = (int64_t)p.H * p.W;
  const int64_t total = (int64_t)p.frames * p.H * wp;
  const int64_t idx0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t idx = idx0 < total ? idx0 : total - 1;   // dead lanes recompute the last item (kept for the warp-wide store)
  const int xp = (int)(idx % wp);
  const int y = (int)((idx / wp) % p.H);
  const int64_t f = idx / ((int64_t)wp * p.H);
  const int x0 = xp * 2;
  float in[3][4][CIN];  // rows y-1..y+1, columns x0-1..x0+2
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int yy = y + r - 1, xx = x0 + c - 1;
      const bool ok = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) in[r][c][ch] = ok ? __ldg(p.src + (f * CIN + ch) * hw + (int64_t)yy * p.W + xx) : 0.f;
    }
  float a0[CO], a1[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) { a0[j] = bs[j]; a1[j] = bs[j]; }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) {
        const float v0 = in[r][c][ch], v1 = in[r][c + 1][ch];
        const float4* wr = reinterpret_cast<const float4*>(ws + ((r * 3 + c) * CIN + ch) * CO);
#pragma unroll
        for (int j = 0; j < CO / 4; ++j) {
          const float4 w4 = wr[j];
          a0[4 * j] = fmaf(v0, w4.x, a0[4 * j]); a0[4 * j + 1] = fmaf(v0, w4.y, a0[4 * j + 1]);
          a0[4 * j + 2] = fmaf(v0, w4.z, a0[4 * j + 2]); a0[4 * j + 3] = fmaf(v0, w4.w, a0[4 * j + 3]);
          a1[4 * j] = fmaf(v1, w4.x, a1[4 * j]); a1[4 * j + 1] = fmaf(v1, w4.y, a1[4 * j + 1]);
          a1[4 * j + 2] = fmaf(v1, w4.z, a1[4 * j + 2]); a1[4 * j + 3] = fmaf(v1, w4.w, a1[4 * j + 3]);
        }
      }
  if constexpr (sizeof(T) == 2) {
    // Each thread owns 2 adjacent pixels x 64 channels = 256 contiguous bytes (bf16) of the output row.  A
    // per-thread 16-byte store would touch 32 different lines per instruction, so the warp's 32 x 2 pixels are
    // first transposed through shared memory and then written as contiguous 512-byte runs.
    constexpr int V = Vec16<T>::N;
    constexpr int ROW = 2 * CO;                       // elements per thread
    __shared__ __align__(16) T stage[4][32][ROW + V];  // +V: rows 16 B apart in bank space
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* my = stage[warp][lane];
  #pragma unroll
    for (int j = 0; j < CO; j += V) {
      float t[V];
  #pragma unroll
      for (int e = 0; e < V; ++e) t[e] = a0[j + e];
      store_vec<T, V>(my + j, t);
  #pragma unroll
      for (int e = 0; e < V; ++e) t[e] = a1[j + e];
      store_vec<T, V>(my + CO + j, t);
    }
    __syncwarp();
    // the warp's threads are consecutive pixel pairs of one row segment (idx is linear in (f, y, xp)), except across
    // row ends; write thread by thread, 32 lanes x 16 B per instruction
    constexpr int VPT = ROW / V;                      // 16-byte vectors per thread row
    const int64_t warp_idx0 = (int64_t)blockIdx.x * blockDim.x + warp * 32;
    if ((p.W & 1) == 0) {
      // even width: (frame, row, pixel pair) is linear in memory, so the warp's 32 x 2 pixels are ONE contiguous run
      T* wbase = reinterpret_cast<T*>(p.dst) + warp_idx0 * ROW;
      const int64_t lim = (total - warp_idx0) * (ROW / V);   // vectors that exist (last warp may be partial)
#pragma unroll 4
      for (int i = lane; i < 32 * VPT; i += 32) {
        if (i >= lim) break;
        const int tsrc = i / VPT, vsrc = i - tsrc * VPT;
        *reinterpret_cast<uint4*>(wbase + (int64_t)i * V) = *reinterpret_cast<const uint4*>(stage[warp][tsrc] + vsrc * V);
      }
    } else {
      for (int i = lane; i < 32 * VPT; i += 32) {
        const int tsrc = i / VPT, vsrc = i - tsrc * VPT;
        const int64_t tidx = warp_idx0 + tsrc;
        if (tidx >= total) continue;
        const int txp = (int)(tidx % wp);
        const int ty = (int)((tidx / wp) % p.H);
        const int64_t tf = tidx / ((int64_t)wp * p.H);
        const int tx0 = txp * 2;
        const int el = vsrc * V;                        // element offset inside the 2-pixel run
        if (el >= CO && tx0 + 1 >= p.W) continue;       // odd width: second pixel does not exist
        T* dstp = reinterpret_cast<T*>(p.dst) + ((tf * p.H + ty) * p.W + tx0) * CO + el;
        *reinterpret_cast<uint4*>(dstp) = *reinterpret_cast<const uint4*>(stage[warp][tsrc] + el);
      }
    }
  } else {
    if (idx0 >= total) return;
    constexpr int V4 = Vec16<T>::N;
    T* o = reinterpret_cast<T*>(p.dst) + ((f * p.H + y) * p.W + x0) * CO;
#pragma unroll
    for (int j = 0; j < CO; j += V4) {
      float t[V4];
#pragma unroll
      for (int e = 0; e < V4; ++e) t[e] = a0[j + e];
      store_vec<T, V4>(o + j, t);
    }
    if (x0 + 1 < p.W) {
#pragma unroll
      for (int j = 0; j < CO; j += V4) {
        float t[V4];
#pragma unroll
        for (int e = 0; e < V4; ++e) t[e] = a1[j + e];
        store_vec<T, V4>(o + CO + j, t);
      }
    }
  }
}

// final assembly: channels-last SR + bilinear x4 of the base frame -> planar fp32.  One thread per output pixel; lanes
// run along X so every planar store is a coalesced 128-byte run.
template <typename T>
__global__ void __launch_bounds__(256) assemble_kernel(const fbanet_assemble_params p) {
  const int64_t total = (int64_t)p.N * p.H * p.W;
  const int Hb = p.H >> 2, Wb = p.W >> 2;
  const T* sr = reinterpret_cast<const T*>(p.sr);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % p.W);
    const int y = (int)((idx / p.W) % p.H);
    const int64_t n = idx / ((int64_t)p.W * p.H);
    float sy = 0.25f * (y + 0.5f) - 0.5f, sx = 0.25f * (x + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int yb = (int)sy, xb = (int)sx;
    const int y1 = yb + (yb < Hb - 1 ? 1 : 0), x1 = xb + (xb < Wb - 1 ? 1 : 0);
    const float wy = sy - yb, wx = sx - xb, hy = 1.f - wy, hx = 1.f - wx;
    const T* s = sr + idx * p.Cp;
    for (int c = 0; c < p.C; ++c) {
      const float* bp = p.base + n * p.base_img_stride + (int64_t)c * Hb * Wb;
      const float bl = hy * (hx * __ldg(bp + yb * Wb + xb) + wx * __ldg(bp + yb * Wb + x1)) +
                       wy * (hx * __ldg(bp + y1 * Wb + xb) + wx * __ldg(bp + y1 * Wb + x1));
      const float v = to_f32<T>(s[c]) + (p.lo_offset > 0 ? to_f32<T>(s[c + p.lo_offset]) : 0.f);
      p.out[((n * p.C + c) * p.H + y) * p.W + x] = v + bl;
    }
  }
}

// The bf16-path shape of the same assembly: fp32 SR rows of 8 columns (0..3 hi-weight part, 4..7 lo-weight part), C <= 4.
// grid = (x blocks, rows, images): no 64-bit divisions, the pixel's 32 bytes arrive as two 16-byte loads (the generic kernel's six
// scalar loads and three 64-bit divisions per pixel held it at 2.5 TB/s), same arithmetic -> bit-identical output.
template <int CP>   // 8: hi columns 0..3 + lo columns 4..7; 4: already summed (fold_hi_lo conv)
__global__ void __launch_bounds__(256) assemble_f32x8_kernel(const fbanet_assemble_params p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= p.W) return;
  const int y = blockIdx.y;
  const int Hb = p.H >> 2, Wb = p.W >> 2;
  float sy = 0.25f * (y + 0.5f) - 0.5f, sx = 0.25f * (x + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int yb = (int)sy, xb = (int)sx;
  const int y1 = yb + (yb < Hb - 1 ? 1 : 0), x1 = xb + (xb < Wb - 1 ? 1 : 0);
  const float wy = sy - yb, wx = sx - xb, hy = 1.f - wy, hx = 1.f - wx;
  for (int n = blockIdx.z; n < p.N; n += gridDim.z) {
    const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.sr) + (((int64_t)n * p.H + y) * p.W + x) * CP);
    const float4 hi = __ldg(s), lo = CP == 8 ? __ldg(s + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < p.C) {
        const float* bp = p.base + (int64_t)n * p.base_img_stride + (int64_t)c * Hb * Wb;
        const float bl = hy * (hx * __ldg(bp + yb * Wb + xb) + wx * __ldg(bp + yb * Wb + x1)) +
                         wy * (hx * __ldg(bp + y1 * Wb + xb) + wx * __ldg(bp + y1 * Wb + x1));
        p.out[(((int64_t)n * p.C + c) * p.H + y) * p.W + x] = (CP == 8 ? hv[c] + lv[c] : hv[c]) + bl;
      }
    }
  }
}

// Four consecutive output pixels per thread (W % 4 == 0): at x4 upsampling the pixels x = 4k .. 4k+3 read base columns k-1, k, k+1 only, so
// an interior thread loads 3 columns x 2 rows per channel for four outputs (18 cached loads instead of 48), takes the SR columns as
// 16-byte loads and writes one float4 per channel row; the bilinear weights 0.625 / 0.875 / 0.125 / 0.375 are exact, and the blend keeps
// assemble_f32x8_kernel's operation order (bit-identical results).  The two edge threads of a row use the per-pixel clamped form.
// (the one-pixel kernel ran at 2.7-3.0 TB/s: ~65 instructions per pixel, issue bound)
template <int CP>
__global__ void __launch_bounds__(160) assemble_f32x8_quad_kernel(const fbanet_assemble_params p) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;   // base column = output quad index
  const int Hb = p.H >> 2, Wb = p.W >> 2;
  if (k >= Wb) return;
  const int y = blockIdx.y, x = 4 * k;
  float sy = 0.25f * (y + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  const int yb = (int)sy;
  const int y1 = yb + (yb < Hb - 1 ? 1 : 0);
  const float wy = sy - yb, hy = 1.f - wy;
  const bool interior = k >= 1 && k <= Wb - 2;
  for (int n = blockIdx.z; n < p.N; n += gridDim.z) {
    const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.sr) + (((int64_t)n * p.H + y) * p.W + x) * CP);
    float sr[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 hi = __ldg(s + j * (CP / 4));
      if (CP == 8) {
        const float4 lo = __ldg(s + j * 2 + 1);
        sr[j][0] = hi.x + lo.x; sr[j][1] = hi.y + lo.y; sr[j][2] = hi.z + lo.z; sr[j][3] = hi.w + lo.w;
      } else { sr[j][0] = hi.x; sr[j][1] = hi.y; sr[j][2] = hi.z; sr[j][3] = hi.w; }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < p.C) {
        const float* bp = p.base + (int64_t)n * p.base_img_stride + (int64_t)c * Hb * Wb;
        float o[4];
        if (interior) {
          const float* r0 = bp + yb * Wb + k - 1;
          const float* r1 = bp + y1 * Wb + k - 1;
          const float a0 = __ldg(r0), a1 = __ldg(r0 + 1), a2 = __ldg(r0 + 2), b0 = __ldg(r1), b1 = __ldg(r1 + 1), b2 = __ldg(r1 + 2);
          const float wx[4] = {0.625f, 0.875f, 0.125f, 0.375f};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float hx = 1.f - wx[j];
            const float tl = j < 2 ? a0 : a1, tr = j < 2 ? a1 : a2, bl_ = j < 2 ? b0 : b1, br = j < 2 ? b1 : b2;
            o[j] = sr[j][c] + (hy * (hx * tl + wx[j] * tr) + wy * (hx * bl_ + wx[j] * br));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float sx = 0.25f * (x + j + 0.5f) - 0.5f;
            sx = sx < 0.f ? 0.f : sx;
            const int xb = (int)sx;
            const int x1 = xb + (xb < Wb - 1 ? 1 : 0);
            const float wxj = sx - xb, hx = 1.f - wxj;
            o[j] = sr[j][c] + (hy * (hx * __ldg(bp + yb * Wb + xb) + wxj * __ldg(bp + yb * Wb + x1)) +
                               wy * (hx * __ldg(bp + y1 * Wb + xb) + wxj * __ldg(bp + y1 * Wb + x1)));
          }
        }
        *reinterpret_cast<float4*>(p.out + (((int64_t)n * p.C + c) * p.H + y) * p.W + x) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// channels-last view -> space-to-depth(2), one thread per 16-byte vector of the destination
template <typename T>
__global__ void __launch_bounds__(256) s2d_kernel(const fbanet_s2d_params p) {
  // grid = (vectors of one output row, output rows, images): 32-bit index math only (the flat 64-bit version spent five 64-bit
  // divisions per 16-byte vector)
  constexpr int V = Vec16<T>::N;
  const int cg = p.C / V, Ho = p.H / 2, Wo = p.W / 2;
  const int row_vecs = Wo * 4 * cg;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row_vecs) return;
  const int c0 = (i % cg) * V;
  const int r = i / cg;
  const int sub = r & 3, x = r >> 2;
  const int y = blockIdx.y;
  const T* src = reinterpret_cast<const T*>(p.src);
  T* dst = reinterpret_cast<T*>(p.dst);
  for (int n = blockIdx.z; n < p.N; n += gridDim.z) {
    const T* s = src + (int64_t)n * p.img_stride + ((int64_t)(2 * y + (sub >> 1)) * p.W + (2 * x + (sub & 1))) * p.ld + c0;
    *reinterpret_cast<uint4*>(dst + (((int64_t)n * Ho + y) * row_vecs + i) * V) = *reinterpret_cast<const uint4*>(s);
  }
}

// ------------------------------------------------------------------------------------------------
// K8  LayerNorm: one warp per token, two-pass fp32 statistics, biased variance (torch / equinox).
// ------------------------------------------------------------------------------------------------
template <typename T, int C>
__global__ void __launch_bounds__(256) layernorm_kernel(const fbanet_layernorm_params p) {
  constexpr int PER = C / 32;  // contiguous channels per lane
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const T* x = reinterpret_cast<const T*>(p.x) + row * p.x_ld + lane * PER;
  float v[PER];
  if constexpr (PER % Vec16<T>::N == 0) {
#pragma unroll
    for (int i = 0; i < PER; i += Vec16<T>::N) {
      float t[Vec16<T>::N];
      load_vec<T, Vec16<T>::N>(x + i, t);
#pragma unroll
      for (int j = 0; j < Vec16<T>::N; ++j) v[i + j] = t[j];
    }
  } else {
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = to_f32<T>(x[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + p.eps);
  T* y = reinterpret_cast<T*>(p.y) + row * p.y_ld + lane * PER;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane * PER + i;
    y[i] = from_f32<T>((v[i] - mean) * rstd * __ldg(p.gamma + c) + __ldg(p.beta + c));
  }
}

// K8 (bf16 fast path): C/8 threads per token, RPT tokens per thread group (their 16-byte loads are all issued before the first
// reduction: one load in flight per thread left the kernel at 4.6 TB/s), xor-shuffle reductions, one 16-byte store per token.
template <int C, int RPT>
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const fbanet_layernorm_params p) {
  constexpr int TPT = C / 8;  // threads per token (8, 16 or 32)
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t grp = gt / TPT;
  const int sub = (int)(gt % TPT);
  float v[RPT][8];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int64_t row = grp * RPT + r;   // rows padded to a whole group: keep all lanes alive for the shuffles
    if (row < p.rows) load_vec<bf16, 8>(reinterpret_cast<const bf16*>(p.x) + row * p.x_ld + sub * 8, v[r]);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[r][i] = 0.f;
    }
  }
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + sub * 8)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + sub * 8 + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + sub * 8)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + sub * 8 + 4));
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[r][i];
#pragma unroll
    for (int o = TPT / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[r][i] - mean; q += d * d; }
#pragma unroll
    for (int o = TPT / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / C) + p.eps);
    const int64_t row = grp * RPT + r;
    if (row >= p.rows) continue;
    if (p.stats && sub == 0) *reinterpret_cast<float2*>(p.stats + row * 2) = make_float2(mean, rstd);
    float o8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o8[i] = (v[r][i] - mean) * rstd * gg[i] + bb[i];
    store_vec<bf16, 8>(reinterpret_cast<bf16*>(p.y) + row * p.y_ld + sub * 8, o8);
  }
}

// LayerNorm statistics only (bf16): C/8 threads per row, RPT rows per thread group in flight (a read-only kernel needs
// several independent 16-byte loads per thread to reach the HBM rate); writes (mean, rstd) per row.
template <int C, int RPT>
__global__ void __launch_bounds__(256) rowstats_bf16_kernel(const fbanet_layernorm_params p) {
  constexpr int TPT = C / 8;
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t grp = gt / TPT;
  const int sub = (int)(gt % TPT);
  float v[RPT][8];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int64_t row = grp * RPT + r;
    if (row < p.rows) load_vec<bf16, 8>(reinterpret_cast<const bf16*>(p.x) + row * p.x_ld + sub * 8, v[r]);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[r][i] = 0.f;
    }
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[r][i];
#pragma unroll
    for (int o = TPT / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[r][i] - mean; q += d * d; }
#pragma unroll
    for (int o = TPT / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const int64_t row = grp * RPT + r;
    if (sub == 0 && row < p.rows) *reinterpret_cast<float2*>(p.stats + row * 2) = make_float2(mean, rsqrtf(q * (1.0f / C) + p.eps));
  }
}

// ------------------------------------------------------------------------------------------------
// K7  depthwise 3x3 pad 1 + bias + activation, channels-last.  One thread = one pixel x VEC channels.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const fbanet_dwconv_params p) {
  constexpr int V = Vec16<T>::N;
  const int cg = p.C / V;
  const int64_t total = (int64_t)p.N * p.H * p.W * cg;
  const T* X = reinterpret_cast<const T*>(p.x);
  T* Yo = reinterpret_cast<T*>(p.y);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(idx % cg) * V;
    const int64_t pix = idx / cg;
    const int x = (int)(pix % p.W);
    const int y = (int)((pix / p.W) % p.H);
    const int64_t n = pix / ((int64_t)p.W * p.H);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = __ldg(p.bias + c0 + j);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= p.H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= p.W) continue;
        float v[V];
        load_vec<T, V>(X + ((n * p.H + yy) * p.W + xx) * p.C + c0, v);
        const float* wt = p.weight + (ky * 3 + kx) * p.C + c0;
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = fmaf(v[j], __ldg(wt + j), acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = apply_act(acc[j], p.act, 0.f);
    store_vec<T, V>(Yo + pix * p.C + c0, acc);
  }
}

// K7 (bf16 fast path): one thread = 8 channels x a horizontal run of SEG pixels.  The 9x8 weights live in
// registers for the whole run and a 3x3 window of fp32 columns slides along x (each loaded vector is
// unpacked once and used for three outputs), so every output costs three 16-byte loads instead of nine
// loads plus 80 scalar weight loads.  The x loop is unrolled by 3 so the window rotates by renaming.
__device__ __forceinline__ void unpack8(const uint4& v, float (&o)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(u[i] << 16); o[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
}

template <int V> struct PackV;
template <> struct PackV<8> { typedef uint4 type; };
template <> struct PackV<4> { typedef uint2 type; };
__device__ __forceinline__ void unpackv(const uint2& v, float (&o)[4]) {
  o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
  o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
}
__device__ __forceinline__ void unpackv(const uint4& v, float (&o)[8]) { unpack8(v, o); }

// V = channels per thread (4: ~110 registers, 16+ resident warps/SM; the loads of the next TWO columns
// are in flight while a column is computed)
template <int SEG, int V>
__global__ void __launch_bounds__(128) dwconv3x3_bf16_run_kernel(const fbanet_dwconv_params p) {
  typedef typename PackV<V>::type PV;
  const int cg = p.C / V;
  const int segs = (p.W + SEG - 1) / SEG;
  const int64_t total = (int64_t)p.N * p.H * segs * cg;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (int)(idx % cg) * V;
  int64_t r = idx / cg;
  const int seg = (int)(r % segs); r /= segs;
  const int y = (int)(r % p.H);
  const int64_t n = r / p.H;
  const int x_begin = seg * SEG, x_end = min(x_begin + SEG, p.W);
  const bf16* X = reinterpret_cast<const bf16*>(p.x) + n * (int64_t)p.H * p.W * p.C + c0;
  bf16* Y = reinterpret_cast<bf16*>(p.y) + (n * (int64_t)p.H + y) * p.W * p.C + c0;

  float w[9][V], bias[V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < V; j += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.weight + t * p.C + c0 + j));
      w[t][j] = a.x; w[t][j + 1] = a.y; w[t][j + 2] = a.z; w[t][j + 3] = a.w;
    }
#pragma unroll
  for (int j = 0; j < V; j += 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
    bias[j] = a.x; bias[j + 1] = a.y; bias[j + 2] = a.z; bias[j + 3] = a.w;
  }
  const bool rok[3] = {y - 1 >= 0, true, y + 1 < p.H};
  const bf16* rows[3] = {X + (int64_t)(y - 1) * p.W * p.C, X + (int64_t)y * p.W * p.C, X + (int64_t)(y + 1) * p.W * p.C};
  auto ld = [&](int rr, int xx) -> PV {
    PV z;
    memset(&z, 0, sizeof(z));
    if (!rok[rr] || xx < 0 || xx >= p.W) return z;
    return *reinterpret_cast<const PV*>(rows[rr] + (int64_t)xx * p.C);
  };
  float col[3][3][V];   // [column slot][row][channel]
  PV nxa[3], nxb[3];    // columns x+1 and x+2 in flight
#pragma unroll
  for (int rr = 0; rr < 3; ++rr) { unpackv(ld(rr, x_begin - 1), col[0][rr]); unpackv(ld(rr, x_begin), col[1][rr]); }
#pragma unroll
  for (int rr = 0; rr < 3; ++rr) { nxa[rr] = ld(rr, x_begin + 1); nxb[rr] = ld(rr, x_begin + 2); }

  auto step = [&](int x, float (&cl)[3][V], float (&cm)[3][V], float (&cr)[3][V]) {
    // cl = column x-1, cm = column x, cr <- column x+1 (arrives from the prefetch registers)
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) { unpackv(nxa[rr], cr[rr]); nxa[rr] = nxb[rr]; }
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) nxb[rr] = ld(rr, x + 3);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = bias[j];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
      for (int j = 0; j < V; ++j) {
        acc[j] = fmaf(cl[rr][j], w[rr * 3 + 0][j], acc[j]);
        acc[j] = fmaf(cm[rr][j], w[rr * 3 + 1][j], acc[j]);
        acc[j] = fmaf(cr[rr][j], w[rr * 3 + 2][j], acc[j]);
      }
    if (p.act == FBANET_ACT_GELU_TANH) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = gelu_tanh_fast(acc[j]);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = apply_act(acc[j], p.act, 0.f);
    }
    store_vec<bf16, V>(Y + (int64_t)x * p.C, acc);
  };
  int x = x_begin;
  for (; x + 2 < x_end; x += 3) {
    step(x, col[0], col[1], col[2]);
    step(x + 1, col[1], col[2], col[0]);
    step(x + 2, col[2], col[0], col[1]);
  }
  if (x < x_end) { step(x, col[0], col[1], col[2]); ++x; }
  if (x < x_end) { step(x, col[1], col[2], col[0]); }
}

// ------------------------------------------------------------------------------------------------
// K2a  FAF gate.  One warp per (burst, pixel): the 3x3xC neighbourhood of the base frame is loaded
// once into registers and re-used for the F-1 other frames; lanes split the channels.
// ------------------------------------------------------------------------------------------------
template <typename T, int C>
__global__ void __launch_bounds__(256) faf_gate_kernel(const fbanet_faf_gate_params p) {
  constexpr int PER = C / 32;
  const int lane = threadIdx.x & 31;
  const int64_t hw = (int64_t)p.H * p.W;
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (int64_t)p.B * hw) return;
  const int64_t b = gw / hw, pix = gw % hw;
  const int y = (int)(pix / p.W), x = (int)(pix % p.W);
  const T* feat = reinterpret_cast<const T*>(p.feat);
  float ref[9][PER], wt[9][PER];
  bool ok[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    ok[t] = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      wt[t][i] = __ldg(p.wsum + t * C + lane * PER + i);
      ref[t][i] = ok[t] ? to_f32<T>(feat[((b * p.F) * hw + (int64_t)yy * p.W + xx) * C + lane * PER + i]) : 0.f;
    }
  }
  for (int f = 1; f < p.F; ++f) {
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (!ok[t]) continue;
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      const T* q = feat + ((b * p.F + f) * hw + (int64_t)yy * p.W + xx) * C + lane * PER;
#pragma unroll
      for (int i = 0; i < PER; ++i) acc = fmaf(wt[t][i], to_f32<T>(q[i]) - ref[t][i], acc);
    }
    acc = warp_sum(acc);
    const float g = 1.0f / (1.0f + expf(-fabsf(acc)));
    if (lane == 0 && p.gate) p.gate[(b * (p.F - 1) + (f - 1)) * hw + pix] = g;
    if (p.gated) {  // centre tap of frame f, scaled (federated_affinity_fusion.py:102-105); pixel-major [B][H][W][F][C]
      const T* q = feat + ((b * p.F + f) * hw + pix) * C + lane * PER;
      T* o = reinterpret_cast<T*>(p.gated) + (((b * hw + pix) * p.F) + f) * C + lane * PER;
#pragma unroll
      for (int i = 0; i < PER; ++i) o[i] = from_f32<T>(to_f32<T>(q[i]) * g);
    }
  }
  if (p.gated) {  // frame 0 is passed through unscaled (:103)
    const T* q = feat + ((b * p.F) * hw + pix) * C + lane * PER;
    T* o = reinterpret_cast<T*>(p.gated) + ((b * hw + pix) * p.F) * C + lane * PER;
#pragma unroll
    for (int i = 0; i < PER; ++i) o[i] = q[i];
  }
}

// K2a (bf16 fast path, C = 64): 8 threads per pixel, each owns 8 channels (one 16-byte vector per tap).
// gate_f = sigmoid(|w.x_f - w.x_0|) with w.x the 3x3x64 dot product (linear, so the base-frame term is
// computed once); the centre tap of every frame is re-used for the gated-feature write
// (pixel-major [B][H][W][F][C]).
__global__ void __launch_bounds__(128) faf_gate_bf16_c64_kernel(const fbanet_faf_gate_params p) {
  constexpr int C = 64;
  const int sub = threadIdx.x & 7;
  const int64_t hw = (int64_t)p.H * p.W;
  const int64_t gp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool live = gp < (int64_t)p.B * hw;  // keep dead lanes for the shuffles
  const int64_t gpc = live ? gp : 0;
  const int64_t b = gpc / hw, pix = gpc % hw;
  const int y = (int)(pix / p.W), x = (int)(pix % p.W);
  const bf16* feat = reinterpret_cast<const bf16*>(p.feat) + sub * 8;
  float w[9][8];
  int64_t off[9];
  bool ok[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    ok[t] = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
    off[t] = ((int64_t)yy * p.W + xx) * C;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.wsum + t * C + sub * 8));
    const float4 c = __ldg(reinterpret_cast<const float4*>(p.wsum + t * C + sub * 8 + 4));
    w[t][0] = a.x; w[t][1] = a.y; w[t][2] = a.z; w[t][3] = a.w; w[t][4] = c.x; w[t][5] = c.y; w[t][6] = c.z; w[t][7] = c.w;
  }
  bf16* gated = p.gated ? reinterpret_cast<bf16*>(p.gated) + ((b * hw + pix) * p.F) * C + sub * 8 : nullptr;
  float s0 = 0.f;
  for (int f = 0; f < p.F; ++f) {
    const bf16* fr = feat + (b * p.F + f) * hw * C;
    uint4 v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = ok[t] ? *reinterpret_cast<const uint4*>(fr + off[t]) : make_uint4(0, 0, 0, 0);
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float xv[8];
      unpack8(v[t], xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(w[t][j], xv[j], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (f == 0) {
      s0 = acc;
      if (gated && live) *reinterpret_cast<uint4*>(gated) = v[4];  // frame 0 passes through (:103)
      continue;
    }
    const float g = 1.0f / (1.0f + __expf(-fabsf(acc - s0)));
    if (!live) continue;
    if (sub == 0 && p.gate) p.gate[(b * (p.F - 1) + (f - 1)) * hw + pix] = g;
    if (gated) {
      float xv[8], o[8];
      unpack8(v[4], xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = xv[j] * g;
      store_vec<bf16, 8>(gated + (int64_t)f * C, o);
    }
  }
}

// K2a with precomputed scores (bf16, C = 64): the 3x3x64 dot products come from the tensor cores (implicit GEMM with the
// summed kernel as a 2-row weight matrix, fp32 NHWC_F32 store), so this kernel is pure streaming: 8 threads per pixel,
// 16 bytes each, gate_f = sigmoid(|s_f - s_0|), gated features written pixel-major [B][H][W][F][C].
__global__ void __launch_bounds__(256) faf_gate_apply_bf16_c64_kernel(const fbanet_faf_gate_params p) {
  constexpr int C = 64;
  const int sub = threadIdx.x & 7;
  const int64_t hw = (int64_t)p.H * p.W;
  const int64_t gp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (gp >= (int64_t)p.B * hw) return;
  const int64_t b = gp / hw, pix = gp % hw;
  const bf16* feat = reinterpret_cast<const bf16*>(p.feat) + (b * p.F * hw + pix) * C + sub * 8;
  const float2* sc = reinterpret_cast<const float2*>(p.score) + b * p.F * hw + pix;
  bf16* gated = p.gated ? reinterpret_cast<bf16*>(p.gated) + ((b * hw + pix) * p.F) * C + sub * 8 : nullptr;
  const float2 s00 = __ldg(sc);
  const float s0 = s00.x + s00.y;
  if (gated) *reinterpret_cast<uint4*>(gated) = __ldg(reinterpret_cast<const uint4*>(feat));   // frame 0 passes through (:103)
#pragma unroll 1
  for (int f0 = 1; f0 < p.F; f0 += 4) {   // 4 frames in flight per thread
    uint4 v[4];
    float2 s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (f0 + k < p.F) {
        s[k] = __ldg(sc + (int64_t)(f0 + k) * hw);
        if (gated) v[k] = __ldg(reinterpret_cast<const uint4*>(feat + (int64_t)(f0 + k) * hw * C));
      }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (f0 + k < p.F) {
        const int f = f0 + k;
        const float g = 1.0f / (1.0f + __expf(-fabsf((s[k].x + s[k].y) - s0)));
        if (sub == 0 && p.gate) p.gate[(b * (p.F - 1) + (f - 1)) * hw + pix] = g;
        if (gated) {
          const f32x2 g2 = pack_f2(g, g);
          uint4 o;
          o.x = f2_to_bf16x2(mul_f2(bf16x2_to_f2(v[k].x), g2));
          o.y = f2_to_bf16x2(mul_f2(bf16x2_to_f2(v[k].y), g2));
          o.z = f2_to_bf16x2(mul_f2(bf16x2_to_f2(v[k].z), g2));
          o.w = f2_to_bf16x2(mul_f2(bf16x2_to_f2(v[k].w), g2));
          *reinterpret_cast<uint4*>(gated + (int64_t)f * C) = o;
        }
      }
  }
}

// ------------------------------------------------------------------------------------------------
// full-size tiling: reflect index helper (torch 'reflect': no edge repeat)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// divide: dst[tile][t][c][ty][tx] = padded(src)[t][c][i*ps + ty - ov][j*ps + tx - ov], where `padded`
// is reflect-pad bottom/right to a multiple of psize, then reflect-pad `ov` all round (two nested reflects).
__global__ void __launch_bounds__(256) tile_divide_kernel(const fbanet_tile_params p) {
  const int ts = p.psize + 2 * p.overlap;
  const int Hp = (p.H + p.psize - 1) / p.psize * p.psize, Wp = (p.W + p.psize - 1) / p.psize * p.psize;
  const int nwt = Wp / p.psize;
  const int64_t per_tile = (int64_t)p.T * p.C * ts * ts;
  const int64_t total = (int64_t)(p.tile_end - p.tile_begin) * per_tile;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int tx = (int)(idx % ts);
    const int ty = (int)((idx / ts) % ts);
    const int64_t tc = (idx / ((int64_t)ts * ts)) % ((int64_t)p.T * p.C);
    const int tile = p.tile_begin + (int)(idx / per_tile);
    const int i = tile / nwt, j = tile % nwt;
    int yy = reflect_idx(i * p.psize + ty - p.overlap, Hp);  // outer pad (on the Hp x Wp image)
    int xx = reflect_idx(j * p.psize + tx - p.overlap, Wp);
    yy = reflect_idx(yy, p.H);                                // inner pad (bottom/right only => i >= H)
    xx = reflect_idx(xx, p.W);
    p.dst[idx] = __ldg(p.src + (tc * p.H + yy) * p.W + xx);
  }
}

// merge: dst[c][Y][X] = tiles[tile(Y,X)][c][ov4 + Y%ps4][ov4 + X%ps4]   (psize/overlap given low-res; x scale)
__global__ void __launch_bounds__(256) tile_merge_kernel(const fbanet_tile_params p) {
  const int sc = p.scale, ps = p.psize * sc, ov = p.overlap * sc, ts = ps + 2 * ov;
  const int Wp = (p.W + p.psize - 1) / p.psize * p.psize;
  const int nwt = Wp / p.psize;
  const int HH = p.H * sc, WW = p.W * sc;
  const int64_t total = (int64_t)p.C * HH * WW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(idx % WW);
    const int Y = (int)((idx / WW) % HH);
    const int c = (int)(idx / ((int64_t)WW * HH));
    const int tile = (Y / ps) * nwt + X / ps;
    if (tile < p.tile_begin || tile >= p.tile_end) continue;
    p.dst[idx] = __ldg(p.src + (((int64_t)(tile - p.tile_begin) * p.C + c) * ts + ov + Y % ps) * ts + ov + X % ps);
  }
}

// ---- row-band variants (cfg4 on several GPUs): the source / destination image is a set of row bands, band k possibly in a peer
// GPU's memory.  A tile row that crosses a band boundary is read from (written to) the neighbour over NVLink by plain loads
// (stores): the halo exchange is part of the gather, no staging copy and no collective.
__device__ __forceinline__ int band_of(const fbanet_tile_band_params& p, int y) {
  int k = 0;
#pragma unroll
  for (int b = 1; b < FBANET_MAX_BANDS; ++b) k += (b < p.nbands && y >= p.row0[b]) ? 1 : 0;
  return k;
}

__global__ void __launch_bounds__(256) tile_divide_banded_kernel(const fbanet_tile_band_params p) {
  const int ts = p.psize + 2 * p.overlap;
  const int Hp = (p.H + p.psize - 1) / p.psize * p.psize, Wp = (p.W + p.psize - 1) / p.psize * p.psize;
  const int nwt = Wp / p.psize;
  const int64_t per_tile = (int64_t)p.T * p.C * ts * ts;
  const int64_t total = (int64_t)(p.tile_end - p.tile_begin) * per_tile;
  float* dst = reinterpret_cast<float*>(p.tiles);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int tx = (int)(idx % ts);
    const int ty = (int)((idx / ts) % ts);
    const int64_t tc = (idx / ((int64_t)ts * ts)) % ((int64_t)p.T * p.C);
    const int tile = p.tile_begin + (int)(idx / per_tile);
    const int i = tile / nwt, j = tile % nwt;
    int yy = reflect_idx(i * p.psize + ty - p.overlap, Hp);
    int xx = reflect_idx(j * p.psize + tx - p.overlap, Wp);
    yy = reflect_idx(yy, p.H);
    xx = reflect_idx(xx, p.W);
    const int k = band_of(p, yy);
    const int r0 = p.row0[k], rows = p.row0[k + 1] - r0;
    dst[idx] = *(reinterpret_cast<const float*>(p.band[k]) + (tc * rows + (yy - r0)) * p.W + xx);   // possibly a peer load: no __ldg
  }
}

// one thread per pixel of a tile's x`scale` centre square; the owner band of the output row takes the store
__global__ void __launch_bounds__(256) tile_merge_banded_kernel(const fbanet_tile_band_params p) {
  const int sc = p.scale, ps = p.psize * sc, ov = p.overlap * sc, ts = ps + 2 * ov;
  const int Wp = (p.W + p.psize - 1) / p.psize * p.psize;
  const int nwt = Wp / p.psize;
  const int HH = p.H * sc, WW = p.W * sc;
  const int64_t per_tile = (int64_t)p.C * ps * ps;
  const int64_t total = (int64_t)(p.tile_end - p.tile_begin) * per_tile;
  const float* src = reinterpret_cast<const float*>(p.tiles);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % ps);
    const int y = (int)((idx / ps) % ps);
    const int c = (int)((idx / ((int64_t)ps * ps)) % p.C);
    const int lt = (int)(idx / per_tile), tile = p.tile_begin + lt;
    const int Y = (tile / nwt) * ps + y, X = (tile % nwt) * ps + x;
    if (Y >= HH || X >= WW) continue;   // the reflect padding to a multiple of psize is cropped away
    const int k = band_of(p, Y / sc);
    const int r0 = p.row0[k] * sc, rows = (p.row0[k + 1] - p.row0[k]) * sc;
    reinterpret_cast<float*>(p.band[k])[((int64_t)c * rows + (Y - r0)) * WW + X] = __ldg(src + (((int64_t)lt * p.C + c) * ts + ov + y) * ts + ov + x);
  }
}

// ------------------------------------------------------------------------------------------------
// 8f-4  optical-flow registration: bilinear sample at (y - fy, x - fx), indices clamped (map_coordinates order=1, mode="nearest")
// ------------------------------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(256) flow_warp_kernel(const fbanet_flow_warp_params p) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;   // pixel inside the frame; frame = blockIdx.y (no 64-bit divisions)
  if (r >= p.H * p.W) return;
  const int C = CT ? CT : p.C;
  const int y = r / p.W;
  const int x = r - y * p.W;
  const int f = blockIdx.y;
  const float* s = p.src + (int64_t)f * p.s_frame;
  float* d = p.dst + (int64_t)f * p.d_frame + (int64_t)y * p.d_y + (int64_t)x * p.d_x;
  int64_t ff = f;   // index of this frame's flow field
  if (p.frames_per_burst > 0) {
    const int b = f / p.frames_per_burst, t = f - b * p.frames_per_burst;
    if (t == 0) {   // base frame: not registered
      const float* s0 = s + (int64_t)y * p.s_y + (int64_t)x * p.s_x;
      for (int c = 0; c < C; ++c) d[(int64_t)c * p.d_c] = __ldg(s0 + (int64_t)c * p.s_c);
      return;
    }
    ff = (int64_t)b * (p.frames_per_burst - 1) + (t - 1);
  }
  const float2 fl = __ldg(reinterpret_cast<const float2*>(p.flow) + (ff * p.H + y) * p.W + x);   // (dy, dx)
  const float cy = __fsub_rn((float)y, fl.x), cx = __fsub_rn((float)x, fl.y);                      // grid - flow, one fp32 rounding
  const float fy = floorf(cy), fx = floorf(cx);
  const float uy = cy - fy, ux = cx - fx, ly = 1.f - uy, lx = 1.f - ux;                              // upper / lower weights
  // clamp in float first (a wild flow must not overflow the int conversion), then to the image
  const int y0i = (int)fminf(fmaxf(fy, -1.f), (float)p.H), x0i = (int)fminf(fmaxf(fx, -1.f), (float)p.W);
  const int y0 = min(max(y0i, 0), p.H - 1), y1 = min(max(y0i + 1, 0), p.H - 1);
  const int x0 = min(max(x0i, 0), p.W - 1), x1 = min(max(x0i + 1, 0), p.W - 1);
  const float* r0 = s + (int64_t)y0 * p.s_y;
  const float* r1 = s + (int64_t)y1 * p.s_y;
  const int64_t o0 = (int64_t)x0 * p.s_x, o1 = (int64_t)x1 * p.s_x;
  float t[CT ? CT : 1][4];
  if (CT) {
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const int64_t oc = (int64_t)c * p.s_c;
      t[c][0] = __ldg(r0 + o0 + oc); t[c][1] = __ldg(r0 + o1 + oc); t[c][2] = __ldg(r1 + o0 + oc); t[c][3] = __ldg(r1 + o1 + oc);
    }
#pragma unroll
    for (int c = 0; c < CT; ++c)   // jax sums the four (index, weight) products in the order (lo,lo), (lo,hi), (hi,lo), (hi,hi)
      d[(int64_t)c * p.d_c] = ((t[c][0] * (ly * lx) + t[c][1] * (ly * ux)) + t[c][2] * (uy * lx)) + t[c][3] * (uy * ux);
  } else {
    for (int c = 0; c < C; ++c) {
      const int64_t oc = (int64_t)c * p.s_c;
      d[(int64_t)c * p.d_c] = ((__ldg(r0 + o0 + oc) * (ly * lx) + __ldg(r0 + o1 + oc) * (ly * ux)) + __ldg(r1 + o0 + oc) * (uy * lx)) + __ldg(r1 + o1 + oc) * (uy * ux);
    }
  }
}

static int grid_for(int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  const int64_t cap = 148 * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fbanet

namespace fbanet {
int head_conv_tc_supported(const fbanet_head_conv_params* p);
int head_conv_tc_launch(const fbanet_head_conv_params* p, cudaStream_t stream);
}  // namespace fbanet

using namespace fbanet;

extern "C" int fbanet_warp_sm100(const fbanet_warp_params* p, void* stream) {
  if (!p || !p->src || !p->dst || !p->M || p->frames <= 0 || p->frames_per_burst <= 0 || p->H <= 0 || p->W <= 0 || p->C <= 0)
    return FBANET_E_BADSHAPE;
  const int64_t total = (int64_t)p->frames * p->H * p->W;
  const bool planar4 = !p->coords && p->s_x == 1 && p->d_x == 1 && (p->W % 4) == 0 && (p->C == 3 || p->C == 4) && ((uintptr_t)p->src % 16) == 0 &&
                       ((uintptr_t)p->dst % 16) == 0 && (p->s_y % 4) == 0 && (p->d_y % 4) == 0 && (p->s_c % 4) == 0 && (p->d_c % 4) == 0 &&
                       (p->s_frame % 4) == 0 && (p->d_frame % 4) == 0 && (int64_t)p->C * p->s_c < ((int64_t)1 << 31) &&
                       (int64_t)p->H * p->s_y < ((int64_t)1 << 31) && (int64_t)p->C * p->d_c < ((int64_t)1 << 31) && (int64_t)p->H * p->d_y < ((int64_t)1 << 31) &&
                       total / 4 < ((int64_t)1 << 31) - 128;
  static const char* p4env = getenv("FBANET_WARP_PLANAR4");   // experiment switch: 0 = one pixel per thread everywhere
  // (one warp per destination row with the lanes on consecutive pixels was tried against the four-pixels-per-thread form: 0.334 vs
  // 0.203 ms -- the kernel is bound by instruction issue, not by L1 wavefronts; profiles/r2_ncu_warp.txt)
  // shared-memory staged form (default for planar layouts; FBANET_WARP_TILE=0: four adjacent pixels per thread gathering from global memory)
  static const char* tlenv = getenv("FBANET_WARP_TILE");
  if (planar4 && !(tlenv && tlenv[0] == '0') && !(p4env && p4env[0] == '0')) {
    // tile 32 x 16, or 16 x 32 where that covers the frame with fewer idle lanes (W = 80: 5 x 16 instead of 3 x 32)
    const int64_t cover32 = (int64_t)ceil_div(p->W, 32) * 32 * ceil_div(p->H, 16) * 16, cover16 = (int64_t)ceil_div(p->W, 16) * 16 * ceil_div(p->H, 32) * 32;
    const int tw = cover16 < cover32 ? 16 : 32;
    const int tiles_x = ceil_div(p->W, tw), tiles_y = ceil_div(p->H, 512 / tw);
    const int64_t blocks = (int64_t)tiles_x * tiles_y * p->frames;
    if (blocks < ((int64_t)1 << 31)) {
      const cudaStream_t st = (cudaStream_t)stream;
      if (p->C == 3 && tw == 32) warp_tile_kernel<3, 32><<<(unsigned)blocks, 128, 0, st>>>(*p, tiles_x, tiles_y);
      else if (p->C == 3) warp_tile_kernel<3, 16><<<(unsigned)blocks, 128, 0, st>>>(*p, tiles_x, tiles_y);
      else if (tw == 32) warp_tile_kernel<4, 32><<<(unsigned)blocks, 128, 0, st>>>(*p, tiles_x, tiles_y);
      else warp_tile_kernel<4, 16><<<(unsigned)blocks, 128, 0, st>>>(*p, tiles_x, tiles_y);
      return check_launch();
    }
  }
  if (planar4 && !(p4env && p4env[0] == '0')) {
    const int b4 = ceil_div(total / 4, 128);
    if (p->C == 3) warp_planar4_kernel<3><<<b4, 128, 0, (cudaStream_t)stream>>>(*p);
    else warp_planar4_kernel<4><<<b4, 128, 0, (cudaStream_t)stream>>>(*p);
    return check_launch();
  }
  if (p->frames > 65535 || (int64_t)p->H * p->W > (int64_t)1 << 30) return FBANET_E_BADSHAPE;
  const dim3 blocks((unsigned)ceil_div((int64_t)p->H * p->W, 256), (unsigned)p->frames);
  if (p->C == 3) warp_kernel<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  else if (p->C == 4) warp_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  else warp_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}

extern "C" int fbanet_to_nhwc_sm100(const fbanet_to_nhwc_params* p, void* stream) {
  if (!p || !p->src || !p->dst || p->Cp < p->C || p->frames <= 0) return FBANET_E_BADSHAPE;
  const int64_t total = (int64_t)p->frames * p->H * p->W;
  if (p->dtype == FBANET_F32) to_nhwc_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  else if (p->dtype == FBANET_BF16) {
    if (p->im2col3x3 && p->Cp % 8 == 0 && ((uintptr_t)p->dst % 16) == 0)
      im2col3x3_bf16_kernel<<<grid_for(total * (p->Cp / 8), 256), 256, 0, (cudaStream_t)stream>>>(*p);
    else
      to_nhwc_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  } else return FBANET_E_DTYPE;
  return check_launch();
}

template <typename T>
static int launch_head(const fbanet_head_conv_params* p, cudaStream_t s) {
  const int64_t total = (int64_t)p->frames * p->H * ((p->W + 1) / 2);
  if (p->C == 3) head_conv_kernel<T, 3><<<ceil_div(total, 128), 128, 0, s>>>(*p);
  else if (p->C == 4) head_conv_kernel<T, 4><<<ceil_div(total, 128), 128, 0, s>>>(*p);
  else return FBANET_E_BADSHAPE;
  return check_launch();
}

extern "C" int fbanet_head_conv_sm100(const fbanet_head_conv_params* p, void* stream) {
  if (!p || !p->src || !p->dst || !p->weight || !p->bias || p->frames <= 0 || p->Cout != 64) return FBANET_E_BADSHAPE;
  if ((uintptr_t)p->dst % 16) return FBANET_E_ALIGN;
  if (p->M && (p->frames_per_burst <= 0 || ((uintptr_t)p->M % 8))) return FBANET_E_BADSHAPE;
  if (p->dtype == FBANET_F32) return p->M ? FBANET_E_UNSUPPORTED : launch_head<float>(p, (cudaStream_t)stream);
  if (p->dtype == FBANET_BF16) {
    static const char* no_tc = getenv("FBANET_HEAD_TC");   // experiment switch: 0 = CUDA-core head conv
    if (!(no_tc && no_tc[0] == '0') && head_conv_tc_supported(p)) return head_conv_tc_launch(p, (cudaStream_t)stream);
    return p->M ? FBANET_E_UNSUPPORTED : launch_head<bf16>(p, (cudaStream_t)stream);   // the fused warp lives in the tensor-core kernel only
  }
  return FBANET_E_DTYPE;
}

extern "C" int fbanet_assemble_sm100(const fbanet_assemble_params* p, void* stream) {
  if (!p || !p->sr || !p->base || !p->out || p->N <= 0 || p->C <= 0 || p->Cp < p->C || (p->H % 4) || (p->W % 4) || p->lo_offset < 0 || (p->lo_offset > 0 && p->lo_offset + p->C > p->Cp)) return FBANET_E_BADSHAPE;
  const int64_t total = (int64_t)p->N * p->H * p->W;
  if (p->dtype == FBANET_F32 && ((p->Cp == 8 && p->lo_offset == 4) || (p->Cp == 4 && p->lo_offset == 0)) && p->C <= 4 && ((uintptr_t)p->sr % 16) == 0 &&
      p->H <= 65535) {
    static const char* qenv = getenv("FBANET_ASSEMBLE_QUAD");   // experiment switch: 0 = one output pixel per thread
    if ((p->W % 4) == 0 && p->W >= 16 && ((uintptr_t)p->out % 16) == 0 && !(qenv && qenv[0] == '0')) {
      const int quads = p->W / 4;
      const int bt = quads >= 128 ? (quads % 128 == 0 ? 128 : (quads % 160 == 0 ? 160 : 128)) : (quads + 31) / 32 * 32;   // 640 px rows: 160 threads, no idle warps
      const dim3 gq((unsigned)ceil_div(quads, bt), (unsigned)p->H, (unsigned)(p->N < 65535 ? p->N : 65535));
      if (p->Cp == 8) assemble_f32x8_quad_kernel<8><<<gq, bt, 0, (cudaStream_t)stream>>>(*p);
      else assemble_f32x8_quad_kernel<4><<<gq, bt, 0, (cudaStream_t)stream>>>(*p);
      return check_launch();
    }
    const dim3 grid((unsigned)ceil_div(p->W, 256), (unsigned)p->H, (unsigned)(p->N < 65535 ? p->N : 65535));
    if (p->Cp == 8) assemble_f32x8_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    else assemble_f32x8_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    return check_launch();
  }
  if (p->dtype == FBANET_F32) assemble_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  else if (p->dtype == FBANET_BF16) assemble_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  else return FBANET_E_DTYPE;
  return check_launch();
}

extern "C" int fbanet_space_to_depth_sm100(const fbanet_s2d_params* p, void* stream) {
  if (!p || !p->src || !p->dst || p->N <= 0 || (p->H % 2) || (p->W % 2)) return FBANET_E_BADSHAPE;
  const int v = p->dtype == FBANET_F32 ? 4 : 8;
  if (p->C % v) return FBANET_E_BADSHAPE;
  if ((p->ld % v) || (p->img_stride % v) || ((uintptr_t)p->src % 16) || ((uintptr_t)p->dst % 16)) return FBANET_E_ALIGN;
  if (p->H / 2 > 65535) return FBANET_E_BADSHAPE;
  const dim3 grid((unsigned)ceil_div((int64_t)(p->W / 2) * 4 * (p->C / v), 256), (unsigned)(p->H / 2), (unsigned)(p->N < 65535 ? p->N : 65535));
  if (p->dtype == FBANET_F32) s2d_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  else if (p->dtype == FBANET_BF16) s2d_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  else return FBANET_E_DTYPE;
  return check_launch();
}

template <typename T>
static int launch_ln(const fbanet_layernorm_params* p, cudaStream_t s) {
  const int blocks = ceil_div(p->rows, 8);
  switch (p->C) {
    case 32: layernorm_kernel<T, 32><<<blocks, 256, 0, s>>>(*p); break;
    case 64: layernorm_kernel<T, 64><<<blocks, 256, 0, s>>>(*p); break;
    case 128: layernorm_kernel<T, 128><<<blocks, 256, 0, s>>>(*p); break;
    case 256: layernorm_kernel<T, 256><<<blocks, 256, 0, s>>>(*p); break;
    case 512: layernorm_kernel<T, 512><<<blocks, 256, 0, s>>>(*p); break;
    default: return FBANET_E_BADSHAPE;
  }
  return check_launch();
}

extern "C" int fbanet_layernorm_sm100(const fbanet_layernorm_params* p, void* stream) {
  if (!p || !p->x || (!p->y && !p->stats) || p->rows <= 0) return FBANET_E_BADSHAPE;
  if (p->y && (!p->gamma || !p->beta)) return FBANET_E_BADSHAPE;
  const int v = p->dtype == FBANET_F32 ? 4 : 8;
  if (p->x_ld % v || ((uintptr_t)p->x % 16)) return FBANET_E_ALIGN;
  if (p->stats && (p->dtype != FBANET_BF16 || (p->C != 64 && p->C != 128 && p->C != 256) || ((uintptr_t)p->stats % 8))) return FBANET_E_UNSUPPORTED;
  if (p->dtype == FBANET_F32) return launch_ln<float>(p, (cudaStream_t)stream);
  if (p->dtype == FBANET_BF16) {
    const bool fast = !p->y || ((p->y_ld % 8) == 0 && ((uintptr_t)p->y % 16) == 0 && ((uintptr_t)p->gamma % 16) == 0 && ((uintptr_t)p->beta % 16) == 0);
    if (p->stats && !fast) return FBANET_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    if (!p->y) {   // statistics only
      constexpr int RPT = 4;
      const int64_t groups = (p->rows + RPT - 1) / RPT;
      if (p->C == 64) rowstats_bf16_kernel<64, RPT><<<ceil_div(groups * 8, 256), 256, 0, st>>>(*p);
      else if (p->C == 128) rowstats_bf16_kernel<128, RPT><<<ceil_div(groups * 16, 256), 256, 0, st>>>(*p);
      else rowstats_bf16_kernel<256, RPT><<<ceil_div(groups * 32, 256), 256, 0, st>>>(*p);
      return check_launch();
    }
    static const int ln_rpt = [] { const char* e = getenv("FBANET_LN_RPT"); return e ? atoi(e) : 4; }();   // rows in flight per thread group: 4 (2 for C = 256) measured 1-7 % faster than 2 (1); same arithmetic per row
    if (ln_rpt == 4) {
      constexpr int RPT = 4;
      const int64_t groups = (p->rows + RPT - 1) / RPT;
      if (fast && p->C == 64) { layernorm_bf16_kernel<64, RPT><<<ceil_div(groups * 8, 256), 256, 0, st>>>(*p); return check_launch(); }
      if (fast && p->C == 128) { layernorm_bf16_kernel<128, RPT><<<ceil_div(groups * 16, 256), 256, 0, st>>>(*p); return check_launch(); }
      if (fast && p->C == 256) { layernorm_bf16_kernel<256, 2><<<ceil_div(((p->rows + 1) / 2) * 32, 256), 256, 0, st>>>(*p); return check_launch(); }
    }
    {
      constexpr int RPT = 2;
      const int64_t groups = (p->rows + RPT - 1) / RPT;
      if (fast && p->C == 64) { layernorm_bf16_kernel<64, RPT><<<ceil_div(groups * 8, 256), 256, 0, st>>>(*p); return check_launch(); }
      if (fast && p->C == 128) { layernorm_bf16_kernel<128, RPT><<<ceil_div(groups * 16, 256), 256, 0, st>>>(*p); return check_launch(); }
      if (fast && p->C == 256) { layernorm_bf16_kernel<256, 1><<<ceil_div(p->rows * 32, 256), 256, 0, st>>>(*p); return check_launch(); }   // a full warp per row already
    }
    return launch_ln<bf16>(p, st);
  }
  return FBANET_E_DTYPE;
}

extern "C" int fbanet_dwconv3x3_sm100(const fbanet_dwconv_params* p, void* stream) {
  if (!p || !p->x || !p->y || !p->weight || !p->bias || p->N <= 0) return FBANET_E_BADSHAPE;
  const int v = p->dtype == FBANET_F32 ? 4 : 8;
  if (p->C % v) return FBANET_E_BADSHAPE;
  if (((uintptr_t)p->x % 16) || ((uintptr_t)p->y % 16)) return FBANET_E_ALIGN;
  const int64_t total = (int64_t)p->N * p->H * p->W * (p->C / v);
  if (p->dtype == FBANET_F32) dwconv3x3_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  else if (p->dtype == FBANET_BF16) {
    if (((uintptr_t)p->weight % 16) || ((uintptr_t)p->bias % 16)) return FBANET_E_ALIGN;
    constexpr int SEG = 20, V = 4;
    const int64_t threads = (int64_t)p->N * p->H * ((p->W + SEG - 1) / SEG) * (p->C / V);
    dwconv3x3_bf16_run_kernel<SEG, V><<<ceil_div(threads, 128), 128, 0, (cudaStream_t)stream>>>(*p);
  } else return FBANET_E_DTYPE;
  return check_launch();
}

template <typename T>
static int launch_gate(const fbanet_faf_gate_params* p, cudaStream_t s) {
  const int blocks = ceil_div((int64_t)p->B * p->H * p->W, 8);
  switch (p->C) {
    case 32: faf_gate_kernel<T, 32><<<blocks, 256, 0, s>>>(*p); break;
    case 64: faf_gate_kernel<T, 64><<<blocks, 256, 0, s>>>(*p); break;
    case 128: faf_gate_kernel<T, 128><<<blocks, 256, 0, s>>>(*p); break;
    default: return FBANET_E_BADSHAPE;
  }
  return check_launch();
}

extern "C" int fbanet_faf_gate_sm100(const fbanet_faf_gate_params* p, void* stream) {
  if (!p || !p->feat || (!p->gate && !p->gated) || !p->wsum || p->B <= 0 || p->F < 2) return FBANET_E_BADSHAPE;
  if (p->dtype == FBANET_F32) return launch_gate<float>(p, (cudaStream_t)stream);
  if (p->score) {   // scores precomputed on the tensor cores
    if (p->dtype != FBANET_BF16 || p->C != 64 || ((uintptr_t)p->feat % 16) || ((uintptr_t)p->gated % 16) || ((uintptr_t)p->score % 8)) return FBANET_E_UNSUPPORTED;
    const int64_t threads = (int64_t)p->B * p->H * p->W * 8;
    faf_gate_apply_bf16_c64_kernel<<<ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(*p);
    return check_launch();
  }
  if (p->dtype == FBANET_BF16) {
    if (p->C == 64 && ((uintptr_t)p->feat % 16) == 0 && ((uintptr_t)p->wsum % 16) == 0 && ((uintptr_t)p->gated % 16) == 0) {
      const int64_t threads = (int64_t)p->B * p->H * p->W * 8;
      faf_gate_bf16_c64_kernel<<<ceil_div(threads, 128), 128, 0, (cudaStream_t)stream>>>(*p);
      return check_launch();
    }
    return launch_gate<bf16>(p, (cudaStream_t)stream);
  }
  return FBANET_E_DTYPE;
}

extern "C" int fbanet_tile_divide_sm100(const fbanet_tile_params* p, void* stream) {
  if (!p || !p->src || !p->dst || p->tile_end <= p->tile_begin || p->overlap >= p->H || p->overlap >= p->W) return FBANET_E_BADSHAPE;
  const int ts = p->psize + 2 * p->overlap;
  const int64_t total = (int64_t)(p->tile_end - p->tile_begin) * p->T * p->C * ts * ts;
  tile_divide_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}

static bool bands_ok(const fbanet_tile_band_params* p) {
  if (!p || !p->tiles || p->nbands < 1 || p->nbands > FBANET_MAX_BANDS || p->tile_end <= p->tile_begin) return false;
  if (p->row0[0] != 0 || p->row0[p->nbands] != p->H) return false;
  for (int k = 0; k < p->nbands; ++k)
    if (!p->band[k] || p->row0[k + 1] <= p->row0[k]) return false;
  return true;
}

extern "C" int fbanet_tile_divide_banded_sm100(const fbanet_tile_band_params* p, void* stream) {
  if (!bands_ok(p) || p->overlap >= p->H || p->overlap >= p->W) return FBANET_E_BADSHAPE;
  const int ts = p->psize + 2 * p->overlap;
  const int64_t total = (int64_t)(p->tile_end - p->tile_begin) * p->T * p->C * ts * ts;
  tile_divide_banded_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}

extern "C" int fbanet_tile_merge_banded_sm100(const fbanet_tile_band_params* p, void* stream) {
  if (!bands_ok(p) || p->scale < 1) return FBANET_E_BADSHAPE;
  const int64_t total = (int64_t)(p->tile_end - p->tile_begin) * p->C * p->psize * p->scale * p->psize * p->scale;
  tile_merge_banded_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}

namespace fbanet {
// Planar layouts (x stride 1 on both sides, W % 4 == 0, 16-byte aligned rows): one thread = 4 consecutive destination pixels -- two
// 16-byte flow loads, 16 * CT unconditional taps in flight, one float4 store per channel row; 32-bit element offsets from frame
// bases pinned in registers (the one-pixel kernel is latency bound at 3.2 TB/s).  Same arithmetic per pixel as flow_warp_kernel.
template <int CT>
__global__ void __launch_bounds__(128) flow_warp_planar4_kernel(const fbanet_flow_warp_params p) {
  const int W4 = p.W >> 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)p.frames * p.H * W4) return;
  const int x = (int)(idx % W4) * 4;
  const int y = (int)((idx / W4) % p.H);
  const int f = (int)(idx / ((int64_t)W4 * p.H));
  const float* s = p.src + (int64_t)f * p.s_frame;
  float* d = p.dst + (int64_t)f * p.d_frame;
  asm volatile("" : "+l"(s), "+l"(d));
  const unsigned sy_ = (unsigned)p.s_y, sc_ = (unsigned)p.s_c, dc_ = (unsigned)p.d_c;
  const unsigned dofs = (unsigned)y * (unsigned)p.d_y + (unsigned)x;
  int64_t ff = f;
  if (p.frames_per_burst > 0) {
    const int b = f / p.frames_per_burst, t = f - b * p.frames_per_burst;
    if (t == 0) {   // base frame: not registered
      const unsigned sofs = (unsigned)y * sy_ + (unsigned)x;
#pragma unroll
      for (int c = 0; c < CT; ++c) *reinterpret_cast<float4*>(d + (dofs + c * dc_)) = __ldg(reinterpret_cast<const float4*>(s + (sofs + c * sc_)));
      return;
    }
    ff = (int64_t)b * (p.frames_per_burst - 1) + (t - 1);
  }
  const float4* fp = reinterpret_cast<const float4*>(p.flow + ((ff * p.H + y) * p.W + x) * 2);
  const float4 fa = __ldg(fp), fb = __ldg(fp + 1);
  const float fdy[4] = {fa.x, fa.z, fb.x, fb.z}, fdx[4] = {fa.y, fa.w, fb.y, fb.w};   // (dy, dx) per pixel
  float w4[4][4];
  unsigned o00[4], o01[4], o10[4], o11[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float cy = __fsub_rn((float)y, fdy[k]), cx = __fsub_rn((float)(x + k), fdx[k]);   // grid - flow, one fp32 rounding
    const float fy = floorf(cy), fx = floorf(cx);
    const float uy = cy - fy, ux = cx - fx, ly = 1.f - uy, lx = 1.f - ux;
    const int y0i = (int)fminf(fmaxf(fy, -1.f), (float)p.H), x0i = (int)fminf(fmaxf(fx, -1.f), (float)p.W);
    const unsigned y0 = (unsigned)min(max(y0i, 0), p.H - 1) * sy_, y1 = (unsigned)min(max(y0i + 1, 0), p.H - 1) * sy_;
    const unsigned x0 = (unsigned)min(max(x0i, 0), p.W - 1), x1 = (unsigned)min(max(x0i + 1, 0), p.W - 1);
    o00[k] = y0 + x0; o01[k] = y0 + x1; o10[k] = y1 + x0; o11[k] = y1 + x1;
    w4[k][0] = ly * lx; w4[k][1] = ly * ux; w4[k][2] = uy * lx; w4[k][3] = uy * ux;
  }
  float t[CT][4][4];
#pragma unroll
  for (int c = 0; c < CT; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned oc = c * sc_;
      t[c][k][0] = __ldg(s + (o00[k] + oc));
      t[c][k][1] = __ldg(s + (o01[k] + oc));
      t[c][k][2] = __ldg(s + (o10[k] + oc));
      t[c][k][3] = __ldg(s + (o11[k] + oc));
    }
#pragma unroll
  for (int c = 0; c < CT; ++c) {
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)   // jax sums the four (index, weight) products in the order (lo,lo), (lo,hi), (hi,lo), (hi,hi)
      o[k] = ((t[c][k][0] * w4[k][0] + t[c][k][1] * w4[k][1]) + t[c][k][2] * w4[k][2]) + t[c][k][3] * w4[k][3];
    *reinterpret_cast<float4*>(d + (dofs + c * dc_)) = make_float4(o[0], o[1], o[2], o[3]);
  }
}
}  // namespace fbanet

extern "C" int fbanet_flow_warp_sm100(const fbanet_flow_warp_params* p, void* stream) {
  if (!p || !p->src || !p->dst || !p->flow || p->frames <= 0 || p->frames_per_burst < 0 || p->frames_per_burst == 1 || p->H <= 0 ||
      p->W <= 0 || p->C <= 0 || ((uintptr_t)p->flow % 8))
    return FBANET_E_BADSHAPE;
  if (p->frames_per_burst > 0 && p->frames % p->frames_per_burst) return FBANET_E_BADSHAPE;
  // experiment switch, off by default: four pixels per thread measured SLOWER than one (0.246 vs 0.225 ms, 64 x 14 x 3 x 160^2:
  // profiles/r2_w_warp_flow_ab.log) -- the flow field adds 8 bytes per pixel of streaming reads and the kernel wants more threads
  // in flight, not more work per thread
  const bool planar4 = p->s_x == 1 && p->d_x == 1 && (p->W % 4) == 0 && (p->C == 3 || p->C == 4) && ((uintptr_t)p->src % 16) == 0 &&
                       ((uintptr_t)p->dst % 16) == 0 && ((uintptr_t)p->flow % 16) == 0 && (p->s_y % 4) == 0 && (p->d_y % 4) == 0 && (p->s_c % 4) == 0 &&
                       (p->d_c % 4) == 0 && (p->s_frame % 4) == 0 && (p->d_frame % 4) == 0 && (int64_t)p->C * p->s_c < ((int64_t)1 << 31) &&
                       (int64_t)p->H * p->s_y < ((int64_t)1 << 31) && (int64_t)p->C * p->d_c < ((int64_t)1 << 31) && (int64_t)p->H * p->d_y < ((int64_t)1 << 31);
  static const char* f4env = getenv("FBANET_FLOW_PLANAR4");
  if (planar4 && f4env && f4env[0] == '1') {
    const int b4 = ceil_div((int64_t)p->frames * p->H * (p->W / 4), 128);
    if (p->C == 3) flow_warp_planar4_kernel<3><<<b4, 128, 0, (cudaStream_t)stream>>>(*p);
    else flow_warp_planar4_kernel<4><<<b4, 128, 0, (cudaStream_t)stream>>>(*p);
    return check_launch();
  }
  if (p->frames > 65535 || (int64_t)p->H * p->W > (int64_t)1 << 30) return FBANET_E_BADSHAPE;
  const dim3 blocks((unsigned)ceil_div((int64_t)p->H * p->W, 256), (unsigned)p->frames);
  if (p->C == 3) flow_warp_kernel<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  else if (p->C == 4) flow_warp_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  else flow_warp_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}

extern "C" int fbanet_tile_merge_sm100(const fbanet_tile_params* p, void* stream) {
  if (!p || !p->src || !p->dst || p->tile_end <= p->tile_begin || p->scale < 1) return FBANET_E_BADSHAPE;
  const int64_t total = (int64_t)p->C * p->H * p->scale * p->W * p->scale;
  tile_merge_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}
