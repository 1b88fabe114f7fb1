// Implicit-GEMM convolution / linear layer on the CUDA cores (fp32 FMA, fp32 accumulate).
//
// This is the PARITY path (dtype fp32: max-abs <= 1e-3 against the CPU oracle) and the general
// fallback for shapes the tcgen05 kernel does not take.  One kernel serves conv3x3, conv1x1, conv4x4 s2,
// ConvTranspose 2x2 s2 (as a per-pixel GEMM with a scatter store), Linear, with up to 16 concat-free
// sources, optional per-pixel source scaling (FAF gate), bias, ReLU/PReLU/GELU, residual add, and the
// four store modes of include/fbanet_b200.h.
#include "common.cuh"

namespace fbanet {

constexpr int BM = 64, BN = 64, BK = 16, SPAD = 4;

struct RowInfo { int n, iy0, ix0, valid; };

template <typename T>
__global__ void __launch_bounds__(256) conv_gemm_simt_kernel(const fbanet_conv_params p, const int vec_ok) {
  __shared__ __align__(16) float As[BK][BM + SPAD];
  __shared__ __align__(16) float Bs[BK][BN + SPAD];
  __shared__ RowInfo rows[BM];
  __shared__ int cstart[FBANET_MAX_SRC + 1];
  __shared__ fbanet_src ssrc[FBANET_MAX_SRC];  // dynamic indexing of kernel params would go through local memory

  const int tid = threadIdx.x;
  const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;

  if (tid < p.nsrc) ssrc[tid] = p.src[tid];
  if (tid == 0) {
    int acc = 0;
    for (int s = 0; s < p.nsrc; ++s) { cstart[s] = acc; acc += p.src[s].C; }
    for (int s = p.nsrc; s <= FBANET_MAX_SRC; ++s) cstart[s] = acc;
  }
  if (tid < BM) {
    const int64_t m = m0 + tid;
    RowInfo ri;
    ri.valid = m < M;
    const int64_t mm = ri.valid ? m : 0;
    const int ox = (int)(mm % p.Wo), oy = (int)((mm / p.Wo) % p.Ho);
    ri.n = (int)(mm / ((int64_t)p.Wo * p.Ho));
    ri.iy0 = oy * p.stride - p.pad;
    ri.ix0 = ox * p.stride - p.pad;
    rows[tid] = ri;
  }
  __syncthreads();
  const int Ctot = cstart[FBANET_MAX_SRC];
  const int K = p.KH * p.KW * Ctot;
  const T* Wt = reinterpret_cast<const T*>(p.weight);

  const int lr = tid >> 2;        // tile row (A) / tile col (B) this thread loads
  const int lk = (tid & 3) * 4;   // first of 4 consecutive k
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const RowInfo ri = rows[lr];

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- gather A (im2col on the fly) ----
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const int k = k0 + lk;
      if (ri.valid && k < K) {
        if (vec_ok) {
          const int tap = k / Ctot, cg = k - tap * Ctot;
          const int ky = tap / p.KW, kx = tap - ky * p.KW;
          const int iy = ri.iy0 + ky, ix = ri.ix0 + kx;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
            int s = 0;
            while (cg >= cstart[s + 1]) ++s;
            const fbanet_src& S = ssrc[s];
            const int64_t pix = (int64_t)iy * p.W + ix;
            const T* q = reinterpret_cast<const T*>(S.ptr) + ri.n * S.img_stride + pix * S.ld + (cg - cstart[s]);
            load_vec<T, 4>(q, a);
            if (S.row_scale) {
              const float g = __ldg(S.row_scale + ri.n * S.scale_img_stride + pix);
#pragma unroll
              for (int e = 0; e < 4; ++e) a[e] *= g;
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int kk = k + e;
            if (kk >= K) break;
            const int tap = kk / Ctot, cg = kk - tap * Ctot;
            const int ky = tap / p.KW, kx = tap - ky * p.KW;
            const int iy = ri.iy0 + ky, ix = ri.ix0 + kx;
            if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
            int s = 0;
            while (cg >= cstart[s + 1]) ++s;
            const fbanet_src& S = ssrc[s];
            const int64_t pix = (int64_t)iy * p.W + ix;
            float v = to_f32<T>(reinterpret_cast<const T*>(S.ptr)[ri.n * S.img_stride + pix * S.ld + (cg - cstart[s])]);
            if (S.row_scale) v *= __ldg(S.row_scale + ri.n * S.scale_img_stride + pix);
            a[e] = v;
          }
        }
      }
    }
    // ---- load B (weights, K-major rows) ----
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const int k = k0 + lk;
      const int col = col0 + lr;
      if (col < p.Cout && k < K) {
        const T* q = Wt + (int64_t)col * K + k;
        if (vec_ok) load_vec<T, 4>(q, b);
        else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (k + e < K) b[e] = to_f32<T>(q[e]);
        }
      }
    }
    __syncthreads();  // previous tile fully consumed
#pragma unroll
    for (int e = 0; e < 4; ++e) { As[lk + e][lr] = a[e]; Bs[lk + e][lr] = b[e]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }

  // ---- epilogue ----
  const float alpha = (p.act == FBANET_ACT_PRELU && p.alpha) ? __ldg(p.alpha) : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    const int64_t m = m0 + r;
    if (m >= M) continue;
    const int ox = (int)(m % p.Wo), oy = (int)((m / p.Wo) % p.Ho), n = (int)(m / ((int64_t)p.Wo * p.Ho));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= p.Cout_store) continue;
      float v = acc[i][j] + (p.bias ? __ldg(p.bias + col) : 0.f);
      v = apply_act(v, p.act, alpha);
      if (p.store_mode == FBANET_STORE_NHWC) {
        const int64_t pix = (int64_t)oy * p.Wo + ox;
        if (p.residual) v += to_f32<T>(reinterpret_cast<const T*>(p.residual)[n * p.res_img_stride + pix * p.res_ld + col]);
        reinterpret_cast<T*>(p.out)[n * p.out_img_stride + pix * p.out_ld + col] = from_f32<T>(v);
      } else if (p.store_mode == FBANET_STORE_PS2) {
        const int c = col >> 2, di = (col >> 1) & 1, dj = col & 1;
        const int64_t pix = (int64_t)(2 * oy + di) * (2 * p.Wo) + (2 * ox + dj);
        reinterpret_cast<T*>(p.out)[n * p.out_img_stride + pix * p.out_ld + c] = from_f32<T>(v);
      } else if (p.store_mode == FBANET_STORE_CONVT2) {
        const int Co = p.Cout >> 2;
        const int q = col / Co, co = col - q * Co, di = q >> 1, dj = q & 1;
        const int64_t pix = (int64_t)(2 * oy + di) * (2 * p.Wo) + (2 * ox + dj);
        reinterpret_cast<T*>(p.out)[n * p.out_img_stride + pix * p.out_ld + co] = from_f32<T>(v);
      } else {  // FBANET_STORE_NCHW_BASE: + bilinear x4 (align_corners=False) of the low-res base frame
        const int Hb = p.Ho >> 2, Wb = p.Wo >> 2;
        float sy = 0.25f * (oy + 0.5f) - 0.5f, sx = 0.25f * (ox + 0.5f) - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        sx = sx < 0.f ? 0.f : sx;
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < Hb - 1 ? 1 : 0), x1 = x0 + (x0 < Wb - 1 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
        const float* bp = p.base + n * p.base_img_stride + (int64_t)col * Hb * Wb;
        const float bl = hy * (hx * __ldg(bp + y0 * Wb + x0) + lx * __ldg(bp + y0 * Wb + x1)) +
                         ly * (hx * __ldg(bp + y1 * Wb + x0) + lx * __ldg(bp + y1 * Wb + x1));
        reinterpret_cast<float*>(p.out)[n * p.out_img_stride + ((int64_t)col * p.Ho + oy) * p.Wo + ox] = v + bl;
      }
    }
  }
}

int conv_gemm_validate(const fbanet_conv_params* p) {
  if (!p || !p->weight || !p->out || p->nsrc < 1 || p->nsrc > FBANET_MAX_SRC) return FBANET_E_BADSHAPE;
  if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->Ho <= 0 || p->Wo <= 0 || p->Cout <= 0 || p->KH <= 0 || p->KW <= 0 || p->stride <= 0)
    return FBANET_E_BADSHAPE;
  if (p->Cout_store <= 0 || p->Cout_store > p->Cout) return FBANET_E_BADSHAPE;
  if (p->src_s2d && (p->KH != 4 || p->KW != 4 || p->stride != 2 || p->pad != 1 || (p->H % 2) || (p->W % 2))) return FBANET_E_BADSHAPE;
  if ((p->H + 2 * p->pad - p->KH) / p->stride + 1 != p->Ho || (p->W + 2 * p->pad - p->KW) / p->stride + 1 != p->Wo) return FBANET_E_BADSHAPE;
  for (int s = 0; s < p->nsrc; ++s)
    if (!p->src[s].ptr || p->src[s].C <= 0 || p->src[s].ld < p->src[s].C) return FBANET_E_BADSHAPE;
  if (p->residual && p->store_mode != FBANET_STORE_NHWC) return FBANET_E_BADSHAPE;
  if ((p->store_mode == FBANET_STORE_PS2 || p->store_mode == FBANET_STORE_CONVT2) && (p->Cout % 4)) return FBANET_E_BADSHAPE;
  if (p->store_mode == FBANET_STORE_NCHW_BASE && (!p->base || (p->Ho % 4) || (p->Wo % 4))) return FBANET_E_BADSHAPE;
  if (p->store_mode < 0 || p->store_mode > 4) return FBANET_E_BADSHAPE;
  if (p->store_mode == FBANET_STORE_NHWC_F32 && p->dtype != FBANET_BF16) return FBANET_E_BADSHAPE;
  if (p->act == FBANET_ACT_PRELU && !p->alpha) return FBANET_E_BADSHAPE;
  if (p->dtype != FBANET_F32 && p->dtype != FBANET_BF16) return FBANET_E_DTYPE;
  return FBANET_OK;
}

int conv_gemm_simt_launch(const fbanet_conv_params* p, cudaStream_t stream) {
  const int esz = p->dtype == FBANET_F32 ? 4 : 2;
  const int valign = 4 * esz;  // bytes of a 4-element vector
  int vec_ok = ((uintptr_t)p->weight % valign) == 0;
  for (int s = 0; s < p->nsrc; ++s) {
    const fbanet_src& S = p->src[s];
    if ((S.C % 4) || (S.ld % 4) || (S.img_stride % 4) || ((uintptr_t)S.ptr % valign)) vec_ok = 0;
  }
  const int64_t M = (int64_t)p->N * p->Ho * p->Wo;
  dim3 grid(ceil_div(M, BM), ceil_div(p->Cout, BN));
  if (p->dtype == FBANET_F32) conv_gemm_simt_kernel<float><<<grid, 256, 0, stream>>>(*p, vec_ok);
  else conv_gemm_simt_kernel<bf16><<<grid, 256, 0, stream>>>(*p, vec_ok);
  return check_launch();
}

}  // namespace fbanet
