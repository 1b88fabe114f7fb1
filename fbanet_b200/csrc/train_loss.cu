// SURVEY 8f-3, first brick of the training step: the training loss of train.py.bak:118-119,168
//   loss = CharbonnierLoss()(restored, target) + gw_weight * GWLoss()(restored, target)        (losses.py:39-80)
// forward value AND its gradient with respect to `restored`, in one pass over the images, deterministic (two-stage fp64 reduction).
//   Charbonnier: mean(sqrt(d^2 + eps^2)),  d = x - y                                   -> dL/dx = d / sqrt(d^2 + eps^2) / N
//   clamp_restored = 1 (the trainer, train.py.bak:167: `restored = torch.clamp(restored, 0, 1)` BEFORE both criteria): d = clamp(x,0,1) - y
//                and the Charbonnier gradient is gated by [0 <= x <= 1] too (torch.clamp's gradient); the target is not clamped there.
//   GW:          mean((1 + 4|Sx*e|)(1 + 4|Sy*e|)|e|),  e = clamp(x,0,1) - clamp(y,0,1), Sx / Sy = Sobel cross-correlations with zero
//                padding per (n, c) plane (the Sobel filters are linear, so Ix1 - Ix2 = Sx * e)
//                -> dL/de_q = [A_q B_q sgn(e_q) + sum_p Sx[q-p] 4 sgn(gx_p) B_p C_p + sum_p Sy[q-p] 4 sgn(gy_p) A_p C_p] / N,
//                   dL/dx_q = dL/de_q * [0 <= x_q <= 1]            (torch.clamp passes the gradient on the closed interval)
#include "common.cuh"

namespace fbanet {

constexpr int TL_THREADS = 256;

__device__ __forceinline__ float tl_clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float tl_sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// e at (yy, xx) of this plane, 0 outside (conv2d padding = 1)
__device__ __forceinline__ float tl_e(const float* x, const float* y, int H, int W, int yy, int xx) {
  if (yy < 0 || yy >= H || xx < 0 || xx >= W) return 0.f;
  return tl_clamp01(__ldg(x + yy * W + xx)) - tl_clamp01(__ldg(y + yy * W + xx));
}

// Sobel responses of e at pixel p = (py, px): gx = Sx * e, gy = Sy * e (cross-correlation, losses.py:62-72)
__device__ __forceinline__ void tl_sobel(const float* x, const float* y, int H, int W, int py, int px, float& gx, float& gy, float& e0) {
  float e[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) e[i][j] = tl_e(x, y, H, W, py + i - 1, px + j - 1);
  gx = (e[0][2] - e[0][0]) + 2.f * (e[1][2] - e[1][0]) + (e[2][2] - e[2][0]);
  gy = (e[2][0] - e[0][0]) + 2.f * (e[2][1] - e[0][1]) + (e[2][2] - e[0][2]);
  e0 = e[1][1];
}

__global__ void __launch_bounds__(TL_THREADS) train_loss_kernel(const fbanet_train_loss_params p) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;      // pixel inside the plane; plane = blockIdx.y
  const int plane = blockIdx.y;
  const int H = p.H, W = p.W;
  const float* x = p.x + (int64_t)plane * H * W;
  const float* y = p.y + (int64_t)plane * H * W;
  double part_c = 0.0, part_g = 0.0;
  if (r < H * W) {
    const int qy = r / W, qx = r - qy * W;
    const float xv = __ldg(x + r), yv = __ldg(y + r);
    const bool inside = xv >= 0.f && xv <= 1.f;
    const float d = (p.clamp_restored ? tl_clamp01(xv) : xv) - yv;
    const float root = sqrtf(fmaf(d, d, p.eps * p.eps));
    part_c = (double)root;
    float grad = (p.clamp_restored && !inside) ? 0.f : d / root * p.inv_n;     // Charbonnier
    if (p.gw_weight != 0.f) {
      float gxq, gyq, eq;
      tl_sobel(x, y, H, W, qy, qx, gxq, gyq, eq);
      const float Aq = 1.f + 4.f * fabsf(gxq), Bq = 1.f + 4.f * fabsf(gyq), Cq = fabsf(eq);
      part_g = (double)(Aq * Bq * Cq);
      float ge = Aq * Bq * tl_sgn(eq);
      // transposed Sobel: pixel p = q + (i-1, j-1) saw e_q through filter tap (2-i, 2-j)
      const float SX[3][3] = {{-1.f, 0.f, 1.f}, {-2.f, 0.f, 2.f}, {-1.f, 0.f, 1.f}};
      const float SY[3][3] = {{-1.f, -2.f, -1.f}, {0.f, 0.f, 0.f}, {1.f, 2.f, 1.f}};
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (i == 1 && j == 1) continue;                                      // both centre taps are zero
          const int py = qy + i - 1, px = qx + j - 1;
          if (py < 0 || py >= H || px < 0 || px >= W) continue;
          float gxp, gyp, ep;
          tl_sobel(x, y, H, W, py, px, gxp, gyp, ep);
          const float Ap = 1.f + 4.f * fabsf(gxp), Bp = 1.f + 4.f * fabsf(gyp), Cp = fabsf(ep);
          ge += SX[2 - i][2 - j] * 4.f * tl_sgn(gxp) * Bp * Cp + SY[2 - i][2 - j] * 4.f * tl_sgn(gyp) * Ap * Cp;
        }
      if (inside) grad = fmaf(p.gw_weight * p.inv_n, ge, grad);
    }
    if (p.grad) p.grad[(int64_t)plane * H * W + r] = grad;
  }
  // block reduction of the two partial sums (fp64), one slot per block: deterministic second stage
  __shared__ double sc[TL_THREADS / 32], sg[TL_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { part_c += __shfl_xor_sync(0xffffffffu, part_c, o); part_g += __shfl_xor_sync(0xffffffffu, part_g, o); }
  if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = part_c; sg[threadIdx.x >> 5] = part_g; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < TL_THREADS / 32; ++w) { a += sc[w]; b += sg[w]; }
    const int64_t slot = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    p.partial[2 * slot] = a;
    p.partial[2 * slot + 1] = b;
  }
}

// second stage: one block sums the per-block partials in a fixed order; loss[0] = total, loss[1] = Charbonnier, loss[2] = GW
__global__ void __launch_bounds__(TL_THREADS) train_loss_finish_kernel(const fbanet_train_loss_params p, int64_t nslots) {
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < nslots; i += TL_THREADS) { a += p.partial[2 * i]; b += p.partial[2 * i + 1]; }
  __shared__ double sa[TL_THREADS], sb[TL_THREADS];
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = TL_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double c = sa[0] * (double)p.inv_n, g = sb[0] * (double)p.inv_n;
    p.loss[1] = c; p.loss[2] = g; p.loss[0] = c + (double)p.gw_weight * g;
  }
}

}  // namespace fbanet

using namespace fbanet;

extern "C" int64_t fbanet_train_loss_workspace_doubles(int32_t planes, int32_t H, int32_t W) {
  if (planes <= 0 || H <= 0 || W <= 0) return -1;
  return 2 * (int64_t)planes * ceil_div((int64_t)H * W, TL_THREADS);
}

extern "C" int fbanet_train_loss_sm100(const fbanet_train_loss_params* p, void* stream) {
  if (!p || !p->x || !p->y || !p->loss || !p->partial || p->planes <= 0 || p->planes > 65535 || p->H <= 0 || p->W <= 0 || !(p->eps > 0.f) ||
      (int64_t)p->H * p->W > (int64_t)1 << 30)
    return FBANET_E_BADSHAPE;
  const dim3 grid((unsigned)ceil_div((int64_t)p->H * p->W, TL_THREADS), (unsigned)p->planes);
  train_loss_kernel<<<grid, TL_THREADS, 0, (cudaStream_t)stream>>>(*p);
  train_loss_finish_kernel<<<1, TL_THREADS, 0, (cudaStream_t)stream>>>(*p, (int64_t)grid.x * grid.y);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// SURVEY 8f-3, second brick: the optimizer step of train.py.bak:72-78 -- torch.optim.Adam / AdamW(lr, betas = (0.9, 0.999),
// eps = 1e-8, weight_decay) -- over ONE flat fp32 buffer holding all 19.2 M parameters (and flat gradient / moment buffers), so the
// whole update is one coalesced bandwidth pass (16 B read + 12 B written per parameter) instead of ~460 per-tensor launches.
//   AdamW (decoupled):  p *= 1 - lr*wd;            Adam (L2):  g += wd * p
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// grad_scale multiplies the incoming gradient first (1 / world after a sum all-reduce, or a loss-scale inverse).
// ------------------------------------------------------------------------------------------------
namespace fbanet {
__global__ void __launch_bounds__(256) adam_step_kernel(const fbanet_adam_params p) {
  const float b2 = p.beta2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    float w = p.param[i], g = p.grad[i] * p.grad_scale, m = p.exp_avg[i], v = p.exp_avg_sq[i];
    if (p.decoupled) w *= 1.f - p.lr * p.weight_decay;
    else g = fmaf(p.weight_decay, w, g);
    m = fmaf(p.one_minus_beta1, g - m, m);        // torch: exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(p.one_minus_beta2 * g, g, b2 * v);   // torch: exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = sqrtf(v) / p.bias2_sqrt + p.eps;
    w -= p.step_size * (m / denom);
    p.param[i] = w; p.exp_avg[i] = m; p.exp_avg_sq[i] = v;
  }
}
}  // namespace fbanet

extern "C" int fbanet_adam_step_sm100(const fbanet_adam_params* p, void* stream) {
  if (!p || !p->param || !p->grad || !p->exp_avg || !p->exp_avg_sq || p->n <= 0 || !(p->bias2_sqrt > 0.f)) return FBANET_E_BADSHAPE;
  int64_t blocks = (p->n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch();
}
