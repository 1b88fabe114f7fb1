// ECC homography estimation on the GPU -- the alignment front end that feeds K1 (fbanet_warp_sm100).
// Replaces `cv2.findTransformECC(gray(img1), gray(img2), eye(3), MOTION_HOMOGRAPHY, (COUNT|EPS, 100, 1e-10))` of
// `register_frame`, homography_alignment.py:19-45 (OpenCV modules/video/src/ecc.cpp; Evangelidis & Psarakis, PAMI 2008,
// forward-additive ECC): every non-base frame of a burst is aligned to frame 0.
//
//   prepare:  gray = sum_c w_c x_c  ->  5x5 Gaussian [1,4,6,4,1]/16 (separable, BORDER_REFLECT_101)  ->  central differences
//   iterate:  ONE CTA per (burst, frame) pair runs all iterations: per pixel warp-back of (image, d/dx, d/dy) with bilinear taps
//             (border 0, exact coordinates -- OpenCV quantises them to 1/32 px), the 8-column homography Jacobian, and 74 running
//             sums (J^T J, J^T i, J^T t, first/second moments inside the warped mask); block reduction in fp64; thread 0 solves the
//             8x8 system (Cholesky, fp64) and updates the warp.  The planes of a pair (4 x 100 KB at 160^2) stay in L2 across the
//             100 iterations; nothing returns to the host until the warp matrices are final.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace fbanet {

__device__ __forceinline__ int refl101(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}

// planes[f][0] = blur5(gray(frame f))
__global__ void __launch_bounds__(256) ecc_blur_kernel(const fbanet_ecc_prepare_params p) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.H * p.W) return;
  const int y = r / p.W, x = r - y * p.W, f = blockIdx.y;
  const float* s = p.src + (int64_t)f * p.s_frame;
  const float k[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  float acc = 0.f;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    const float* row = s + (int64_t)refl101(y + dy, p.H) * p.s_y;
    float racc = 0.f;
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const float* px = row + (int64_t)refl101(x + dx, p.W) * p.s_x;
      float g = 0.f;
      for (int c = 0; c < p.C; ++c) g = fmaf(p.gray_weight[c], __ldg(px + (int64_t)c * p.s_c), g);
      racc = fmaf(k[dx + 2], g, racc);
    }
    acc = fmaf(k[dy + 2], racc, acc);
  }
  p.planes[((int64_t)f * 3 + 0) * p.H * p.W + r] = acc;
}

// planes[f][1] = 0.5 (b(x+1) - b(x-1)), planes[f][2] = 0.5 (b(y+1) - b(y-1)), reflect-101 border (zero on the border)
__global__ void __launch_bounds__(256) ecc_grad_kernel(const fbanet_ecc_prepare_params p) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.H * p.W) return;
  const int y = r / p.W, x = r - y * p.W, f = blockIdx.y;
  const int64_t hw = (int64_t)p.H * p.W;
  const float* b = p.planes + (int64_t)f * 3 * hw;
  const float gx = 0.5f * (b[y * p.W + refl101(x + 1, p.W)] - b[y * p.W + refl101(x - 1, p.W)]);
  const float gy = 0.5f * (b[refl101(y + 1, p.H) * p.W + x] - b[refl101(y - 1, p.H) * p.W + x]);
  p.planes[((int64_t)f * 3 + 1) * hw + r] = gx;
  p.planes[((int64_t)f * 3 + 2) * hw + r] = gy;
}

constexpr int ECC_THREADS = 256;
constexpr int ECC_NSUM = 74;   // 6 moments + 3 x 8 masked projections + 8 unmasked J.i + 36 (upper triangle of J^T J)

__device__ __forceinline__ float tap(const float* pl, int W, int H, int xx, int yy) {
  return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? __ldg(pl + yy * W + xx) : 0.f;
}

// one bilinear tap of (image, d/dx, d/dy) from the blurred image plane alone: the gradients are re-derived on the fly with the same
// formula ecc_grad_kernel uses (bit-identical), so only ONE plane has to live in shared memory
__device__ __forceinline__ void tap3(const float* b, int W, int H, int xx, int yy, float w, float& iw, float& gx, float& gy) {
  // branch-free: out-of-image taps get weight 0 and a clamped (valid) address
  const bool ok = xx >= 0 && xx < W && yy >= 0 && yy < H;
  w = ok ? w : 0.f;
  xx = min(max(xx, 0), W - 1);
  yy = min(max(yy, 0), H - 1);
  const int xp = xx + 1 < W ? xx + 1 : W - 2, xm = xx > 0 ? xx - 1 : 1;      // BORDER_REFLECT_101 neighbours
  const int yp = yy + 1 < H ? yy + 1 : H - 2, ym = yy > 0 ? yy - 1 : 1;
  const float* row = b + yy * W;
  iw = fmaf(w, row[xx], iw);
  gx = fmaf(w, 0.5f * (row[xp] - row[xm]), gx);
  gy = fmaf(w, 0.5f * (b[yp * W + xx] - b[ym * W + xx]), gy);
}

__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // 1 ulp: ~1e-5 px on a 160-px coordinate, far inside the ECC tolerance
  return y;
}

// SMEM: the pair's blurred image and template planes are copied into shared memory once (2 x 100 KB at 160 x 160) and all taps of
// all iterations are served from there (832 pairs: 52.7 -> 29.7 ms together with the branch-free taps); otherwise they are read through L2.
template <bool SMEM>
__global__ void __launch_bounds__(ECC_THREADS) ecc_iterate_kernel(const fbanet_ecc_params p) {
  const int f = blockIdx.x;
  const int fpb = p.frames_per_burst;
  if (f % fpb == 0) return;                                   // base frame: its warp stays as given (identity)
  const int H = p.H, W = p.W, npx = H * W;
  const int64_t hw = (int64_t)npx;
  const float* tmpl = p.planes + (int64_t)(f - f % fpb) * 3 * hw;   // blurred base frame of this burst
  const float* img = p.planes + (int64_t)f * 3 * hw;
  const float* gxp_ = img + hw;
  const float* gyp_ = img + 2 * hw;
  extern __shared__ __align__(16) float ecc_sm[];
  if (SMEM) {
    for (int i = threadIdx.x; i < npx; i += ECC_THREADS) { ecc_sm[i] = __ldg(img + i); ecc_sm[npx + i] = __ldg(tmpl + i); }
  }
  const float* img_s = ecc_sm;
  const float* tmpl_s = ecc_sm + npx;

  __shared__ double red[ECC_THREADS / 32][ECC_NSUM];
  __shared__ double Msh[9];
  __shared__ int stop;
  __shared__ double rho_sh, last_rho_sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 9) Msh[threadIdx.x] = p.warp[(int64_t)f * 9 + threadIdx.x];
  if (threadIdx.x == 0) { stop = 0; rho_sh = -1.0; last_rho_sh = -p.eps; }
  __syncthreads();

  int it = 0;
  for (; it < p.max_iters; ++it) {
    if (fabs(rho_sh - last_rho_sh) < p.eps) break;            // uniform: shared values, read after a barrier
    const float m00 = (float)Msh[0], m01 = (float)Msh[1], m02 = (float)Msh[2], m10 = (float)Msh[3], m11 = (float)Msh[4], m12 = (float)Msh[5],
                m20 = (float)Msh[6], m21 = (float)Msh[7], m22 = (float)Msh[8];
    float a[ECC_NSUM];
#pragma unroll
    for (int k = 0; k < ECC_NSUM; ++k) a[k] = 0.f;
    const int step_y = ECC_THREADS / W, step_x = ECC_THREADS - step_y * W;   // r += ECC_THREADS without a division per pixel
    int py = threadIdx.x / W, px = threadIdx.x - py * W;
    for (int r = threadIdx.x; r < npx; r += ECC_THREADS, px += step_x, py += step_y) {
      if (px >= W) { px -= W; ++py; }
      const float x = (float)px, y = (float)py;
      const float rden = SMEM ? rcp_fast(fmaf(m20, x, fmaf(m21, y, m22))) : 1.0f / fmaf(m20, x, fmaf(m21, y, m22));
      const float sx = fmaf(m00, x, fmaf(m01, y, m02)) * rden, sy = fmaf(m10, x, fmaf(m11, y, m12)) * rden;
      const float fx = floorf(sx), fy = floorf(sy);
      const float ax = sx - fx, ay = sy - fy;
      const int x0 = (int)fminf(fmaxf(fx, -2.f), (float)W + 1.f), y0 = (int)fminf(fmaxf(fy, -2.f), (float)H + 1.f);
      const float w00 = (1.f - ay) * (1.f - ax), w01 = (1.f - ay) * ax, w10 = ay * (1.f - ax), w11 = ay * ax;
      float iw, gxw, gyw;
      if (SMEM) {
        iw = gxw = gyw = 0.f;
        tap3(img_s, W, H, x0, y0, w00, iw, gxw, gyw);
        tap3(img_s, W, H, x0 + 1, y0, w01, iw, gxw, gyw);
        tap3(img_s, W, H, x0, y0 + 1, w10, iw, gxw, gyw);
        tap3(img_s, W, H, x0 + 1, y0 + 1, w11, iw, gxw, gyw);
      } else {
        iw = w00 * tap(img, W, H, x0, y0) + w01 * tap(img, W, H, x0 + 1, y0) + w10 * tap(img, W, H, x0, y0 + 1) + w11 * tap(img, W, H, x0 + 1, y0 + 1);
        gxw = w00 * tap(gxp_, W, H, x0, y0) + w01 * tap(gxp_, W, H, x0 + 1, y0) + w10 * tap(gxp_, W, H, x0, y0 + 1) + w11 * tap(gxp_, W, H, x0 + 1, y0 + 1);
        gyw = w00 * tap(gyp_, W, H, x0, y0) + w01 * tap(gyp_, W, H, x0 + 1, y0) + w10 * tap(gyp_, W, H, x0, y0 + 1) + w11 * tap(gyp_, W, H, x0 + 1, y0 + 1);
      }
      const float rx = rintf(sx), ry = rintf(sy);             // warped all-ones mask, nearest, border 0
      const bool mask = rx >= 0.f && rx <= (float)(W - 1) && ry >= 0.f && ry <= (float)(H - 1);
      const float t = SMEM ? tmpl_s[r] : __ldg(tmpl + r);
      // image_jacobian_homo_ECC
      const float den_ = SMEM ? rcp_fast(fmaf(x, m20, fmaf(y, m21, 1.0f))) : 1.0f / fmaf(x, m20, fmaf(y, m21, 1.0f));
      const float hx = -fmaf(x, m00, fmaf(y, m01, m02)) * den_, hy = -fmaf(x, m10, fmaf(y, m11, m12)) * den_;
      const float gxp = gxw * den_, gyp = gyw * den_;
      const float tmp = fmaf(hx, gxp, hy * gyp);
      const float J[8] = {gxp * x, gyp * x, tmp * x, gxp * y, gyp * y, tmp * y, gxp, gyp};
      if (mask) {
        a[0] += 1.f; a[1] += iw; a[2] = fmaf(iw, iw, a[2]); a[3] += t; a[4] = fmaf(t, t, a[4]); a[5] = fmaf(iw, t, a[5]);
#pragma unroll
        for (int k = 0; k < 8; ++k) { a[6 + k] += J[k]; a[14 + k] = fmaf(J[k], iw, a[14 + k]); a[22 + k] = fmaf(J[k], t, a[22 + k]); }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[30 + k] = fmaf(J[k], iw, a[30 + k]);   // outside the mask the warped image is not zero-meaned
      }
      int q = 38;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = i; j < 8; ++j) { a[q] = fmaf(J[i], J[j], a[q]); ++q; }
    }
    // block reduction: warp shuffles in fp32 partials promoted to fp64, then across warps through shared memory
#pragma unroll
    for (int k = 0; k < ECC_NSUM; ++k) {
      double v = (double)a[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double S[ECC_NSUM];
      for (int k = 0; k < ECC_NSUM; ++k) {
        double v = 0.0;
        for (int w = 0; w < ECC_THREADS / 32; ++w) v += red[w][k];
        S[k] = v;
      }
      const double n = S[0];
      int fail = 0;
      if (!(n > 0.0)) fail = 1;
      const double im = S[1] / n, tm = S[3] / n;
      const double inorm2 = S[2] - n * im * im, tnorm2 = S[4] - n * tm * tm;
      const double corr = S[5] - n * im * tm;
      const double rho = corr / sqrt(inorm2 * tnorm2);
      double ip[8], tp[8], A[8][8];
      for (int k = 0; k < 8; ++k) { ip[k] = S[14 + k] - im * S[6 + k] + S[30 + k]; tp[k] = S[22 + k] - tm * S[6 + k]; }
      int q = 38;
      for (int i = 0; i < 8; ++i)
        for (int j = i; j < 8; ++j) { A[i][j] = S[q]; A[j][i] = S[q]; ++q; }
      // Cholesky A = L L^T (A = J^T J is symmetric positive definite unless the image is degenerate)
      for (int i = 0; i < 8 && !fail; ++i)
        for (int j = 0; j <= i; ++j) {
          double s = A[i][j];
          for (int k = 0; k < j; ++k) s -= A[i][k] * A[j][k];
          if (i == j) { if (!(s > 0.0)) { fail = 1; break; } A[i][i] = sqrt(s); }
          else A[i][j] = s / A[j][j];
        }
      auto solve = [&](const double* b, double* xo) {
        double yv[8];
        for (int i = 0; i < 8; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[i][k] * yv[k]; yv[i] = s / A[i][i]; }
        for (int i = 7; i >= 0; --i) { double s = yv[i]; for (int k = i + 1; k < 8; ++k) s -= A[k][i] * xo[k]; xo[i] = s / A[i][i]; }
      };
      if (!fail && !(rho == rho)) fail = 1;                    // NaN
      if (!fail) {
        double iph[8], rhs[8], dp[8];
        solve(ip, iph);
        double lam_n = inorm2, lam_d = corr;
        for (int k = 0; k < 8; ++k) { lam_n -= ip[k] * iph[k]; lam_d -= tp[k] * iph[k]; }
        if (!(lam_d > 0.0)) fail = 1;                          // "the algorithm stopped before its convergence"
        else {
          const double lam = lam_n / lam_d;
          for (int k = 0; k < 8; ++k) rhs[k] = lam * tp[k] - ip[k];   // J^T (lam * t_zm - i_zm)
          solve(rhs, dp);
          Msh[0] += dp[0]; Msh[3] += dp[1]; Msh[6] += dp[2];
          Msh[1] += dp[3]; Msh[4] += dp[4]; Msh[7] += dp[5];
          Msh[2] += dp[6]; Msh[5] += dp[7];
        }
      }
      last_rho_sh = rho_sh;
      rho_sh = fail ? -1.0 : rho;
      if (fail) stop = 1;
    }
    __syncthreads();
    if (stop) { ++it; break; }
  }
  if (threadIdx.x < 9) p.warp[(int64_t)f * 9 + threadIdx.x] = Msh[threadIdx.x];
  if (threadIdx.x == 0) {
    if (p.rho) p.rho[f] = rho_sh;
    if (p.iters_done) p.iters_done[f] = stop ? -it : it;
  }
}

}  // namespace fbanet

using namespace fbanet;

extern "C" int fbanet_ecc_prepare_sm100(const fbanet_ecc_prepare_params* p, void* stream) {
  if (!p || !p->src || !p->planes || p->frames <= 0 || p->frames > 65535 || p->H < 3 || p->W < 3 || p->C < 1 || p->C > 4 ||
      (int64_t)p->H * p->W > (int64_t)1 << 28)
    return FBANET_E_BADSHAPE;
  const dim3 grid((unsigned)ceil_div((int64_t)p->H * p->W, 256), (unsigned)p->frames);
  ecc_blur_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  ecc_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}

extern "C" int fbanet_ecc_homography_sm100(const fbanet_ecc_params* p, void* stream) {
  if (!p || !p->planes || !p->warp || p->frames <= 0 || p->frames_per_burst < 2 || p->frames % p->frames_per_burst || p->H < 3 || p->W < 3 ||
      p->max_iters < 1 || (int64_t)p->H * p->W > (int64_t)1 << 28)
    return FBANET_E_BADSHAPE;
  // both planes of a pair in shared memory when they fit (160 x 160: 200 KB); FBANET_ECC_SMEM=0 forces the L2 path (experiments)
  const size_t need = (size_t)2 * p->H * p->W * sizeof(float);
  static const char* env = getenv("FBANET_ECC_SMEM");
  if (need <= 200 * 1024 && !(env && env[0] == '0')) {
    static size_t opted = 0;
    if (need > opted) {
      cudaError_t e = cudaFuncSetAttribute(ecc_iterate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
      if (e != cudaSuccess) { cudaGetLastError(); set_last_error(e); return FBANET_E_LAUNCH; }
      opted = need;
    }
    ecc_iterate_kernel<true><<<p->frames, ECC_THREADS, need, (cudaStream_t)stream>>>(*p);
  } else {
    ecc_iterate_kernel<false><<<p->frames, ECC_THREADS, 0, (cudaStream_t)stream>>>(*p);
  }
  return check_launch();
}
