/*
 * fbanet_b200 -- C ABI of the B200 (sm_100a) kernels behind the FBANet BaseModel burst-SR forward.
 *
 * Drop-in boundary: the reference has no first-party native code; every op below
 * replaces a library call the reference reaches through JAX/XLA/cuDNN or OpenCV.  The Python host
 * (fbanet_b200/model.py) mirrors the reference's model API (`get_arch(opt)`, `model(burst) -> SR`,
 * state_dict key layout) and calls these entry points through ctypes.
 *
 * Conventions
 *   - every entry point: `int fbanet_<op>_sm100(const <op>_params*, void* stream)`; `stream` is a
 *     cudaStream_t.  Returns 0 or a negative FBANET_E_* code; never throws, never synchronises, never
 *     allocates or frees caller memory.  Launch errors are reported through the return code
 *     (cudaGetLastError) -- the Python wrapper turns them into RuntimeError.
 *   - all pointers are DEVICE pointers owned by the caller.  Tensors are channels-last views:
 *     element (n, y, x, c) of a view lives at  ptr + n*img_stride + (y*W + x)*ld + c   (in elements).
 *     Tokens [T, C] of the transformer are exactly such views with T = H*W (t = y*W + x).
 *   - dtype: FBANET_F32 (parity path, fp32 storage + fp32 FMA) or FBANET_BF16 (bf16 storage, fp32
 *     accumulate; dense contractions on tcgen05 tensor cores when `impl` allows).
 *   - stateless and re-entrant; stream ordered; safe for one-process-per-GPU sharding.
 */
#ifndef FBANET_B200_H
#define FBANET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FBANET_ABI_VERSION 29

enum { FBANET_F32 = 0, FBANET_BF16 = 1 };

enum {
  FBANET_OK = 0,
  FBANET_E_BADSHAPE = -1,   /* unsupported / inconsistent shape            */
  FBANET_E_ALIGN = -2,      /* pointer or stride not aligned as required   */
  FBANET_E_DTYPE = -3,      /* unsupported dtype for this op               */
  FBANET_E_LAUNCH = -4,     /* CUDA launch failure (see fbanet_last_cuda_error) */
  FBANET_E_UNSUPPORTED = -5 /* requested impl cannot run this problem      */
};

enum { FBANET_ACT_NONE = 0, FBANET_ACT_RELU = 1, FBANET_ACT_PRELU = 2, FBANET_ACT_GELU_TANH = 3, FBANET_ACT_GELU_ERF = 4 };

/* how the GEMM result tile [rows = N*Ho*Wo pixels, cols = Cout] is stored */
enum {
  FBANET_STORE_NHWC = 0,     /* out(n,y,x,col)                                                     */
  FBANET_STORE_PS2 = 1,      /* PixelShuffle(2): col = 4c+2i+j -> out(n, 2y+i, 2x+j, c)            */
  FBANET_STORE_CONVT2 = 2,   /* ConvTranspose 2x2 s2: col = (2i+j)*Co + co -> out(n,2y+i,2x+j,co)  */
  FBANET_STORE_NCHW_BASE = 3, /* fp32 planar out[n][col][y][x] + bilinear x4 of `base` (final conv) */
  FBANET_STORE_NHWC_F32 = 4   /* fp32 out(n,y,x,col) for col < Cout_store, whatever the compute dtype (tensor-core path only):
                                 narrow score maps that must not be rounded to bf16, e.g. the FAF affinity scores */
};

enum { FBANET_IMPL_AUTO = 0, FBANET_IMPL_SIMT = 1, FBANET_IMPL_TCGEN05 = 2 };

#define FBANET_MAX_SRC 16

/* One channels-last source view of an implicit-GEMM operand (concat-free inputs: the K axis of the
 * contraction runs over taps x (src0 channels, src1 channels, ...)). */
typedef struct fbanet_src {
  const void* ptr;        /* element (0,0,0,0)                                                   */
  const float* row_scale; /* optional per-pixel fp32 multiplier (FAF gate), index n*scale_img_stride + y*W + x */
  int64_t img_stride;     /* elements between images                                             */
  int64_t scale_img_stride;
  int32_t C;              /* channels taken from this source                                     */
  int32_t ld;             /* elements between pixels                                             */
} fbanet_src;

/* K3/K4/K5/K9: implicit-GEMM convolution / linear layer with fused epilogue.
 * Replaces eqx.nn.Conv2d / ConvTranspose2d / Linear call sites:
 *   layers/conv2d.py:33-43, layers/conv2d_transpose.py:20-26, layers/linear_projection.py:27-28,
 *   layers/window_attention.py:157, layers/locally_enhanced_feed_forward.py:27,56,
 *   blocks/residual.py:21-29, blocks/upsampler.py:22-32, models/fba_net.py:255-320.
 * out = act(conv(concat(src...)) + bias) + residual, stored per `store_mode`. */
typedef struct fbanet_conv_params {
  fbanet_src src[FBANET_MAX_SRC];
  const void* weight;     /* packed [Cout][KH*KW*Ctot] (k = (ky*KW+kx)*Ctot + c), dtype = `dtype`    */
  const float* bias;      /* [Cout] fp32 or NULL                                                    */
  const float* alpha;     /* PReLU slope (device scalar) when act == PRELU                          */
  const void* residual;   /* optional, NHWC view shaped like the output rows (store NHWC only)      */
  void* out;
  const float* base;      /* STORE_NCHW_BASE: low-res frame, planar fp32 [n][c][Ho/4][Wo/4]          */
  int64_t res_img_stride;
  int64_t out_img_stride; /* elements between output images                                         */
  int64_t base_img_stride;
  int32_t dtype;
  int32_t impl;
  int32_t nsrc;
  int32_t N, H, W;        /* input images and their size                                            */
  int32_t KH, KW, stride, pad;
  int32_t Ho, Wo;         /* GEMM rows per image = Ho*Wo                                            */
  int32_t Cout;           /* GEMM columns (weight rows)                                             */
  int32_t Cout_store;     /* columns actually stored (<= Cout; Cout may be zero-padded)             */
  int32_t act;
  int32_t store_mode;
  int32_t res_ld;
  int32_t out_ld;         /* elements between output pixels (NHWC family)                           */
  int32_t src_s2d;        /* 1: sources are space-to-depth(2) views [N,H/2,W/2,(ys,xs,c)] of the logical
                             [N,H,W,C] inputs (fbanet_space_to_depth_sm100); only KH=KW=4, stride 2, pad 1,
                             tensor-core path.  src[i].C is then 4*C_i.                                */
  int32_t fold_hi_lo;     /* 1 (3x3 stride-1 pad-1 conv, FBANET_STORE_NHWC_F32, tensor-core path): the Cout_store (even, <= 8)
                             weight rows are hi / lo bf16 halves of Cout_store/2 filters -- rows [0, h) hi, [h, 2h) lo -- and
                             out(n,y,x,j) = conv_j^hi + conv_j^lo for j < h is stored (out_ld >= 4 when h = 3: a zero 4th column
                             is written); the final conv of models/fba_net.py:315 in the bf16 path.  FBANET_E_UNSUPPORTED if
                             the tap-stacked kernel mode cannot take the problem.                             */
  /* LayerNorm folded into a 1x1 GEMM (layers/fba_net.py:196,246 feeding linear_projection.py:27-28 /
   * locally_enhanced_feed_forward.py:27), tensor-core path.  With W' = W diag(gamma) and W'' = W' - rowmean(W') 1^T
   * (every row centred, so W'' x = W' (x - mean(x) 1): the mean subtraction lives in the weights), the GEMM runs on the RAW
   * rows x with weight = W'' and
   *   y[r][n] = rstd[r] * acc[r][n] + bias[n],   bias = W beta + b,
   * which equals W LN(x) + b.  ln_stats: fp32 [rows][2] = (mean, rstd) per row from fbanet_layernorm_sm100 (stats mode;
   * only rstd is read).  NULL: plain GEMM. */
  const float* ln_stats;
  /* LayerNorm applied to the GEMM's input INSIDE the kernel (same layers, tensor-core path, one source of 64 / 128 / 256 channels):
   * the raw rows x are loaded as for a plain GEMM and normalised in shared memory -- (x - mean) * rstd * ln_gamma + ln_beta rounded
   * to bf16, operation for operation what fbanet_layernorm_sm100 computes -- before the tensor core reads them; weight / bias are
   * the layer's own.  out = W LN(x) + b bit-identical to LayerNorm kernel + GEMM, without the normalised tensor in HBM.
   * ln_gamma / ln_beta: fp32 [C]; NULL: plain GEMM.  Excludes ln_stats.  FBANET_E_UNSUPPORTED when the shape does not fit. */
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  int32_t store_f16;      /* 1 (FBANET_STORE_NHWC, tensor-core path): `out` receives IEEE fp16 instead of bf16
                             (the hidden map of the dim-256 LeFF, consumed by fbanet_leff_fc2_sm100 with f16 = 1); no residual */
} fbanet_conv_params;

/* K1: homography warp with bilinear sampling.  Replaces cv2.warpPerspective / cv2.warpAffine with
 * INTER_LINEAR + WARP_INVERSE_MAP (homography_alignment.py:46-55,120-129): M maps dst -> src,
 * taps outside the image contribute 0, frame 0 of every burst is copied.  Source coordinates are
 * evaluated in fp64 (exact to <1e-9 px; no 1/32-px quantisation).  fp32 samples, arbitrary strides so
 * both the [F,H,W,C] (cv2) and the [B,T,C,H,W] (model input) layouts are served. */
typedef struct fbanet_warp_params {
  const float* src;
  float* dst;
  const double* M;        /* [frames][9] row-major 3x3                                              */
  double* coords;         /* optional debug output [frames][H][W][2] = (sx, sy), else NULL          */
  int64_t s_frame, s_y, s_x, s_c; /* source strides in elements                                     */
  int64_t d_frame, d_y, d_x, d_c; /* destination strides                                            */
  int32_t frames;         /* total frames = bursts * frames_per_burst                               */
  int32_t frames_per_burst;
  int32_t H, W, C;
  int32_t _pad;
} fbanet_warp_params;

/* planar fp32 burst [frames][C][H][W] -> channels-last [frames][H][W][Cp] (zero padded channels).
 * im2col3x3 = 1: channel (ky*3+kx)*C + c holds src(c, y+ky-1, x+kx-1) (zero outside), so the head conv
 * 3x3 (C_in = 3 or 4) becomes a K = Cp 1x1 GEMM for the tensor cores (models/fba_net.py:255). */
typedef struct fbanet_to_nhwc_params {
  const float* src;
  void* dst;
  int32_t dtype;
  int32_t frames, C, H, W, Cp;
  int32_t im2col3x3;
  int32_t _pad;
} fbanet_to_nhwc_params;

/* Head conv (models/fba_net.py:88,255): 3x3 pad 1 straight from the planar fp32 burst [frames][C][H][W]
 * (C = 3 RGB / 4 RAW) to channels-last [frames][H][W][Cout], Cout = 64.  K = 9*C is too shallow for the
 * tensor cores and the op is store-bound, so it is a CUDA-core kernel (fp32 FMA, fp32 accumulate).
 * weight: fp32 [9*C][Cout] with k = (ky*3+kx)*C + c. */
typedef struct fbanet_head_conv_params {
  const float* src;
  void* dst;
  const float* weight;
  const float* bias;
  int32_t dtype;          /* of dst */
  int32_t frames, C, H, W, Cout;
  /* K1 fused into K0 (homography_alignment.py:46-55 -> models/fba_net.py:255), bf16 / 64-channel tensor-core path only: with `M`
   * ([frames][3][3] float64 dst->src, 8-byte aligned) every input sample is the bilinearly warped burst pixel -- the arithmetic of
   * fbanet_warp_sm100, bit for bit -- so the warped burst is never written; frame f with f % frames_per_burst == 0 is the base frame
   * (copied).  NULL: plain head conv.  FBANET_E_UNSUPPORTED when given on a path that cannot fuse (caller warps first). */
  int32_t frames_per_burst;
  const double* M;
} fbanet_head_conv_params;

/* Final assembly (models/fba_net.py:317-320): out[n][c][Y][X] = sr(n,Y,X,c) + bilinear_x4(base)[n][c][Y][X], with
 * sr channels-last [N,4h,4w,Cp] (first C channels used), base planar fp32 [n][c][h][w] (frame 0 of the burst, half-pixel
 * centres, edge clamp), out planar fp32.  Splitting this from the last conv keeps that conv's epilogue a plain
 * channels-last store and makes the planar fp32 writes fully coalesced. */
typedef struct fbanet_assemble_params {
  const void* sr;
  const float* base;
  float* out;
  int64_t base_img_stride;
  int32_t dtype;          /* of sr */
  int32_t N, C, Cp, H, W; /* H, W = output (x4) size */
  int32_t lo_offset;      /* > 0: SR value of channel c = sr[c] + sr[c + lo_offset] (final conv with hi/lo split weights) */
  int32_t _pad;
} fbanet_assemble_params;

/* Narrow host I/O of the end-to-end path (the reference's data path is 8-bit at both ends): mode U8_TO_F32 = the input
 * normalisation `x.astype(float32) / 255.0` (train.py:82-83); F32_TO_U8 = `clamp(x, 0, 1)` + torchvision `ToPILImage` (`mul(255).byte()`:
 * truncation), test_in_any_resolution.py:93-101; F32_TO_F16 = round-to-nearest cast.  n elements, n % 16 == 0, both pointers 16-byte
 * aligned. */
enum { FBANET_CONVERT_U8_TO_F32 = 0, FBANET_CONVERT_F32_TO_U8 = 1, FBANET_CONVERT_F32_TO_F16 = 2 };
typedef struct fbanet_convert_io_params {
  const void* src;
  void* dst;
  int64_t n;
  int32_t mode, _pad;
} fbanet_convert_io_params;

/* channels-last view [N,H,W,C] -> contiguous [N,H/2,W/2,4C], channel (ys*2+xs)*C + c = src(2y+ys, 2x+xs, c).
 * Feeds the 4x4 stride-2 downsampling convs (layers/downsample_flatten.py:6-13) to the TMA/tcgen05 path. */
typedef struct fbanet_s2d_params {
  const void* src;
  void* dst;
  int64_t img_stride;
  int32_t dtype;
  int32_t N, H, W, C, ld;
} fbanet_s2d_params;

/* K8: LayerNorm over channels, eps, affine (layers/fba_net.py:77-79,196,246). */
typedef struct fbanet_layernorm_params {
  const void* x;
  void* y;                /* NULL with `stats` set: statistics only (the normalisation is folded into the consumer GEMM) */
  const float* gamma;
  const float* beta;
  int64_t rows;
  int32_t C, x_ld, y_ld;
  int32_t dtype;
  float eps;
  int32_t _pad;
  float* stats;           /* optional fp32 [rows][2] = (mean, 1/sqrt(var + eps)) per row */
} fbanet_layernorm_params;

/* K6: windowed multi-head self-attention over a token map (layers/fba_net.py:139-250 +
 * layers/window_attention.py:159-248): cyclic shift, window partition, q*scale, q k^T + relative
 * position bias (+ shift mask, -100), softmax, P v, window reverse, inverse shift.
 * qkv is [B*H*W, 3C] = (q | k | v), heads-major inside each; out is [B*H*W, C]. */
typedef struct fbanet_attn_params {
  const void* qkv;
  void* out;
  const float* bias_table; /* [(2*win-1)^2][heads] fp32                                             */
  int32_t dtype;
  int32_t B, H, W, C, heads, win, shift;
  int32_t qkv_ld, out_ld;
  float scale;
  int32_t impl;
  const float* bias_expanded; /* optional, tensor-core path: dense fp32 [heads][win^2][NP] = log2(e) * bias_table[index(i,j)][h]
                                 for j < win^2 and -1e30 for the padded keys (NP = win^2 rounded up to 16); NULL: the kernel
                                 expands bias_table itself */
  int32_t q_prescaled;        /* nonzero: the q columns already carry scale*log2(e) (folded into the q projection weights by
                                 the caller); `scale` is then ignored.  Tensor-core (bf16) kernels only. */
  int32_t _pad;
  const float* bias_wrap;     /* optional, tcgen05 kernel (d_h = 64) with shift > 0: dense fp32 [heads][3][win^2][NP], the table of
                                 bias_expanded for the windows that wrap around the bottom edge (0), the right edge (1) or both (2):
                                 rows and columns in the order the kernel fetches the tokens (box order: a split along x puts
                                 columns 0..4 of every window row first) and with the shift mask (-100 * log2(e) between tokens of
                                 different boxes = different shift regions) added */
} fbanet_attn_params;

/* K7: LeFF depthwise 3x3 (pad 1) + bias + GELU on a channels-last map
 * (layers/locally_enhanced_feed_forward.py:39-52). weight packed [9][C] fp32. */
typedef struct fbanet_dwconv_params {
  const void* x;
  void* y;
  const float* weight;
  const float* bias;
  int32_t dtype;
  int32_t N, H, W, C;
  int32_t act;
} fbanet_dwconv_params;

/* K7+K5 fused (bf16, tensor cores): LeFF tail  out = Linear2(GELU(depthwise3x3(h1) + b_dw)) + b2 + residual
 * (layers/locally_enhanced_feed_forward.py:39-57 + the residual of layers/fba_net.py:248).  h1 = GELU(Linear1(x))
 * is a contiguous channels-last [N,H,W,Hd] tensor; the depthwise output never goes to HBM.
 * dw_weight: fp32 [9][Hd]; w2: bf16 [C][Hd] (K-major); C in {64,128,256}; Hd % 64 == 0. */
typedef struct fbanet_leff_fc2_params {
  const void* h1;
  const float* dw_weight;
  const float* dw_bias;
  const void* w2;
  const float* bias2;
  const void* residual;   /* optional view [N,H,W,C] */
  void* out;              /* view [N,H,W,C] */
  int64_t res_img_stride, out_img_stride;
  int32_t res_ld, out_ld;
  int32_t N, H, W, C, Hd;
  int32_t act;            /* FBANET_ACT_GELU_TANH / _ERF (applied after the depthwise conv) */
  int32_t f16;            /* 1: `h1` and `w2` hold IEEE fp16 (the fc1 GEMM stored its GELU output with fbanet_conv_params.store_f16): the
                             depthwise conv and its GELU run on packed half2 and the A operand of Linear2 is fp16 (tanh GELU only) */
  int32_t _pad;
} fbanet_leff_fc2_params;

/* The whole LeFF MLP in one kernel (bf16, tensor cores), hidden tensor never in HBM:
 *   out = Linear2(GELU(depthwise3x3(GELU(Linear1(x)) reshaped to the H x W map) + b_dw)) + b2 + residual
 * (layers/locally_enhanced_feed_forward.py:25-57 with the residual of layers/fba_net.py:248; x = LayerNorm2 output).  Linear1 is
 * recomputed on each 8 x 16 tile's one-pixel halo instead of writing / re-reading the 4C-channel hidden map.
 * CONTRACT: w1, bias1, dw_weight, dw_bias hold HALF the layer's values (exact scaling by a power of two; the kernel evaluates
 * GELU on z = x / 2).  w1: bf16 [Hd][C]; dw_weight: fp32 [9][Hd]; w2: bf16 [C][Hd]; C in {64,128}; Hd % 64 == 0, Hd <= 512. */
typedef struct fbanet_leff_mlp_params {
  const void* x;          /* view [N,H,W,C]                                                          */
  const void* w1;
  const float* bias1;
  const float* dw_weight;
  const float* dw_bias;
  const void* w2;
  const float* bias2;
  const void* residual;   /* optional view [N,H,W,C]                                                 */
  void* out;              /* view [N,H,W,C]                                                          */
  int64_t x_img_stride, res_img_stride, out_img_stride;
  int32_t x_ld, res_ld, out_ld;
  int32_t N, H, W, C, Hd;
  int32_t act;            /* FBANET_ACT_GELU_TANH / _ERF (both GELUs)                                */
  int32_t w2_f16;         /* 1: `w2` holds IEEE fp16 (not bf16) and the kernel keeps its on-chip hidden tile and the A operand of
                             Linear2 in fp16 (11 significant bits instead of bf16's 8; |hidden| < 65504), running the depthwise conv
                             and its GELU on packed half2 -- tanh GELU only.  x, w1, residual, out stay bf16.               */
} fbanet_leff_mlp_params;

/* K2 (gate): Federated-Affinity gate (blocks/federated_affinity_fusion.py:79-99).
 * gate[b][f-1][p] = sigmoid(| sum_{tap,c} wsum[tap][c] * (feat[b][f] - feat[b][0])(p+tap, c) |), f >= 1,
 * where wsum = sum over output channels of temporal_attn1.weight -- algebraically identical to the
 * reference's  |sum_c(E_f - R) - sum_c(E_0 - R)|  (R and the bias cancel). */
typedef struct fbanet_faf_gate_params {
  const void* feat;       /* [B][F][H][W][C] channels-last                                          */
  float* gate;            /* [B][F-1][H][W] fp32, or NULL                                           */
  const float* wsum;      /* [9][C] fp32                                                            */
  void* gated;            /* optional [B][H][W][F][C] (pixel-major = the 1x1 fusion conv's K axis f*C + c):
                             feat[b][0] for f = 0, feat[b][f] * gate[b][f-1] for f >= 1 (:102-105,121-128) */
  int32_t dtype;
  int32_t B, F, H, W, C;
  int32_t _pad;
  const float* score;     /* optional [B][F][H][W][2] fp32: the 3x3xC dot products  wsum (*) feat  already computed (on the tensor
                             cores, as hi + lo bf16 halves of wsum: score = s[0] + s[1]); the kernel then only forms the gates
                             and streams the gated features.  NULL: the kernel computes them itself from `wsum`. */
} fbanet_faf_gate_params;

/* K2 in one pass (bf16, tensor cores, C = 64): gate + K = F*64 1x1 fusion conv + bias + PReLU of FAFBlock
 * (blocks/federated_affinity_fusion.py:79-105 and :121-128) straight from the features, each read from HBM once:
 *   s_f = wbar (*) feat_f (3x3),  g_f = sigmoid(|s_f - s_0|) (g_0 = 1),  out = PReLU(sum_f g_f * (feat_f W_f^T) + bias).
 * feat: contiguous [B][F][H][W][64] bf16.  score_weight: bf16 [32][64], rows 2t / 2t+1 = hi / lo bf16 halves of wbar[tap t] (the
 * gate weights summed over output channels, DESIGN.md "FAF gate identity"), rows 18..31 zero.  fuse_weight: bf16 [64][F*64]
 * (feature_fusion.0.weight, input channel f*64 + c).  gate: optional fp32 [B][F-1][H][W].  2 <= F <= 14. */
typedef struct fbanet_faf_fuse_params {
  const void* feat;
  const void* score_weight;
  const void* fuse_weight;
  const float* bias;      /* [64] or NULL */
  const float* alpha;     /* PReLU slope (device scalar) or NULL = 0 */
  float* gate;            /* optional */
  void* out;              /* view [B,H,W,64] bf16 */
  int64_t out_img_stride;
  int32_t out_ld;
  int32_t B, F, H, W, C;
  int32_t _pad[2];
} fbanet_faf_fuse_params;

/* Full-size tiling (utils/dataset_utils.py:5-58,140-180): reflect-pad + overlapping tile gather,
 * centre-crop stitch.  Planar fp32. */
typedef struct fbanet_tile_params {
  const float* src;       /* divide: [T][C][H][W] burst ; merge: [tiles][C][4*(psize+2*ov)]^2 tiles  */
  float* dst;             /* divide: [tiles][T][C][psize+2ov][psize+2ov] ; merge: [C][4H][4W]        */
  int32_t T, C, H, W;     /* low-res burst size                                                     */
  int32_t psize, overlap; /* low-res tile core and halo (80, 40)                                    */
  int32_t tile_begin, tile_end; /* tile range handled by this call (rank sharding)                  */
  int32_t scale;          /* 1 for divide, 4 for merge                                              */
  int32_t _pad;
} fbanet_tile_params;

/* Full-size tiling over ROW BANDS held on several GPUs of one box (BASELINE config 4): the [T][C][H][W] burst is
 * sharded by image rows, band k = rows [row0[k], row0[k+1]) of every (t, c) plane, stored [T][C][rows_k][W] on GPU k.  band[k]
 * is that buffer's address AS MAPPED INTO THE CALLING PROCESS (CUDA peer / symmetric memory: the kernel's loads and stores go
 * over NVLink), so one launch does the reflect-padded tile gather AND the halo exchange: a tile whose 40-pixel halo crosses a band
 * boundary simply reads those rows from the neighbour's memory.  The merge writes each x4 tile centre into the band that owns
 * its output rows (band k of the output = rows [scale*row0[k], scale*row0[k+1]) of [C][scale*H][scale*W], stored
 * [C][scale*rows_k][scale*W]).  Same index arithmetic as fbanet_tile_params (utils/dataset_utils.py:5-58,140-180;
 * test_in_any_resolution.py:62-101); with nbands = 1 and row0 = {0, H} it is the plain divide / merge. */
#define FBANET_MAX_BANDS 8
typedef struct fbanet_tile_band_params {
  void* band[FBANET_MAX_BANDS];       /* divide: source bands (read) ; merge: output bands (written)     */
  void* tiles;                        /* divide: dst [tiles][T][C][ts][ts] ; merge: src [tiles][C][scale*ts]^2, local */
  int32_t row0[FBANET_MAX_BANDS + 1]; /* low-res row boundaries, row0[0] = 0, row0[nbands] = H           */
  int32_t nbands;
  int32_t T, C, H, W;
  int32_t psize, overlap;
  int32_t tile_begin, tile_end;
  int32_t scale;                      /* 1 for divide, 4 for merge                                       */
  int32_t _pad;
} fbanet_tile_band_params;

/* Optical-flow registration of a burst.  Replaces registration/optical_flow/register.py:11-47
 * (`jsp.ndimage.map_coordinates(frame, grid - flow, order=1, mode="nearest")` per channel): destination pixel (y, x) samples
 * the source at (y - flow[y][x][0], x - flow[y][x][1]) -- flow's last axis is (dy, dx) -- bilinearly, indices clamped to the
 * image (edge replicate).  The coordinate is formed in fp32 exactly as the reference does (int grid - fp32 flow, one
 * rounding); the interpolation weights are (c - floor(c)) and 1 - that, as jax's map_coordinates computes them.
 * flow holds one field per NON-base frame: [bursts][frames_per_burst - 1][H][W][2]; frame 0 of every burst is copied
 * (pipeline/real_bsr_iterator.py:121-166: "the reference frame should not be included").  frames_per_burst = 0: every
 * frame has a flow field (the single-frame register_frame call).  Strides as in fbanet_warp_params. */
typedef struct fbanet_flow_warp_params {
  const float* src;
  float* dst;
  const float* flow;
  int64_t s_frame, s_y, s_x, s_c;
  int64_t d_frame, d_y, d_x, d_c;
  int32_t frames, frames_per_burst;
  int32_t H, W, C;
  int32_t _pad;
} fbanet_flow_warp_params;

/* ECC homography estimation.  Replaces `cv2.findTransformECC(gray(img1), gray(img2), eye(3), MOTION_HOMOGRAPHY,
 * (COUNT | EPS, 100, 1e-10))` of `register_frame`, homography_alignment.py:19-45: every non-base frame of a burst is aligned to
 * frame 0; the resulting 3x3 matrices map base-frame coordinates to frame coordinates, i.e. they are exactly the `M` that
 * fbanet_warp_sm100 (cv2.warpPerspective with WARP_INVERSE_MAP, :46-55) takes.
 * Step 1, fbanet_ecc_prepare_sm100: gray = sum_c gray_weight[c] * x_c (cv2.cvtColor(..., COLOR_BGR2GRAY): 0.114, 0.587, 0.299),
 * GaussianBlur 5x5 (OpenCV's fixed [1,4,6,4,1]/16 kernel, BORDER_REFLECT_101), central-difference gradients ->
 * planes [frames][3][H][W] = (blurred, d/dx, d/dy).  Strides as in fbanet_warp_params.
 * Step 2, fbanet_ecc_homography_sm100: the forward-additive ECC iterations, one CTA per (burst, frame) pair, all on the device. */
typedef struct fbanet_ecc_prepare_params {
  const float* src;
  float* planes;                  /* out [frames][3][H][W] fp32                                          */
  int64_t s_frame, s_y, s_x, s_c;
  float gray_weight[4];
  int32_t frames, H, W, C;
} fbanet_ecc_prepare_params;

typedef struct fbanet_ecc_params {
  const float* planes;            /* [frames][3][H][W] from fbanet_ecc_prepare_sm100                     */
  double* warp;                   /* in/out [frames][9] row-major 3x3: initial guess (identity) -> estimate; base frames untouched */
  double* rho;                    /* optional out [frames]: final enhanced correlation coefficient (-1: failed) */
  int32_t* iters_done;            /* optional out [frames]: iterations run; negative = stopped on failure (degenerate image,
                                     lambda_d <= 0: cv2 throws "the algorithm stopped before its convergence")   */
  double eps;                     /* stop when |rho - last_rho| < eps                                     */
  int32_t frames, frames_per_burst, H, W, max_iters, _pad;
} fbanet_ecc_params;

/* Training loss (first brick of the training step): train.py.bak:118-119,168
 *   loss = CharbonnierLoss()(restored, target) + gw_weight * GWLoss()(restored, target)      (losses.py:39-51, 53-80)
 * value and gradient with respect to `restored` in one pass.  x = restored, y = target: planar fp32 [planes = B*C][H][W];
 * inv_n = 1 / (planes*H*W) (both losses are means); eps = 1e-3, gw_weight = 3 in the reference; gw_weight = 0 skips GWLoss.
 * partial: workspace of fbanet_train_loss_workspace_doubles(planes, H, W) doubles; loss: 3 doubles (total, Charbonnier, GW).
 * clamp_restored = 1 restates train.py.bak:167 (clamp of the network output before the criteria): Charbonnier then sees
 * clamp(x,0,1) - y and passes no gradient where x left [0,1]; 0 = the bare losses.py criteria on x as given.
 * The reduction is two-stage in fp64 with a fixed order: results are bit-reproducible. */
typedef struct fbanet_train_loss_params {
  const float* x;
  const float* y;
  float* grad;            /* optional dL/dx, same shape as x                                        */
  double* partial;
  double* loss;
  float eps, gw_weight, inv_n;
  int32_t planes, H, W;
  int32_t clamp_restored; /* 1: the trainer's `restored = clamp(restored, 0, 1)` (train.py.bak:167) applied before BOTH criteria */
  int32_t _pad;
} fbanet_train_loss_params;

/* Optimizer step: torch.optim.Adam / AdamW(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay), train.py.bak:72-78, over
 * flat fp32 buffers of n elements.  The host passes the step-dependent scalars: step_size = lr / (1 - beta1^t),
 * bias2_sqrt = sqrt(1 - beta2^t).  decoupled = 1: AdamW (p *= 1 - lr*wd), 0: Adam (g += wd*p).  grad_scale: applied to the gradient
 * first (1/world after a sum all-reduce). */
typedef struct fbanet_adam_params {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay, step_size, bias2_sqrt, grad_scale;
  float one_minus_beta1, one_minus_beta2;   /* formed in double on the host, as torch does (1 - 0.999f is off by 1.3e-5 relative) */
  int32_t decoupled, _pad;
} fbanet_adam_params;

/* Backward bricks of the training step (what torch autograd did for train.py.bak:163-169, what
 * eqx.filter_value_and_grad does at train.py:63).  The DATA gradient of every stride-1 convolution / linear layer is the forward
 * implicit GEMM itself (fbanet_conv_gemm_sm100) run on dY with the spatially flipped, in/out-transposed weights; these three ops are
 * the rest that the convolutional / token-wise layers need.
 *
 * Weight + bias gradient of a convolution or linear layer (layers/conv2d.py:33-43, eqx.nn.Linear call sites):
 *   dw[co][ci][ky][kx] (+)= sum_{n,yo,xo} dy(n,yo,xo,co) * x(n, yo*stride + ky - pad, xo*stride + kx - pad, ci)   (zero outside)
 *   db[co]             (+)= sum_{n,yo,xo} dy(n,yo,xo,co)
 * x, dy: channels-last views (see the conventions above) of `dtype`; dw, db: fp32 in the torch parameter layouts.  The pixel axis is
 * cut into `splits` contiguous chunks whose partial sums go to `partial` (splits * (Cout*KH*KW*Cin + Cout) floats) and are added in
 * a fixed order: bit-reproducible.  accumulate = 1 adds to dw / db instead of overwriting them. */
typedef struct fbanet_wgrad_params {
  const void* x;
  const void* dy;
  float* dw;
  float* db;              /* optional */
  float* partial;
  int64_t x_img_stride, dy_img_stride;
  int32_t x_ld, dy_ld;
  int32_t dtype, N, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad, splits, accumulate;
} fbanet_wgrad_params;

/* LayerNorm backward (eqx.nn.LayerNorm at layers/fba_net.py:77-78, applied per token at :196,246): rows x C, contiguous rows.
 *   xhat = (x - mean) * rstd,  g = dy * gamma,  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),
 *   dgamma[c] (+)= sum_rows dy * xhat,  dbeta[c] (+)= sum_rows dy.       C <= 256.
 * partial: fbanet_layernorm_bwd_blocks(rows) * 2 * C floats; fixed-order reduction. */
typedef struct fbanet_layernorm_bwd_params {
  const void* x;
  const void* dy;
  const float* gamma;
  void* dx;
  float* dgamma;
  float* dbeta;
  float* partial;
  int64_t rows;
  float eps;
  int32_t dtype, C, accumulate;
} fbanet_layernorm_bwd_params;

/* Activation backward: dx = dy * act'(x) on n contiguous elements, x = the PRE-activation.  act: FBANET_ACT_RELU / PRELU /
 * GELU_TANH / GELU_ERF.  PReLU (eqx.nn.PReLU, scalar slope): alpha = device scalar, dalpha (+)= sum dy * x over x < 0 (optional;
 * partial: fbanet_act_bwd_blocks(n) floats). */
typedef struct fbanet_act_bwd_params {
  const void* x;
  const void* dy;
  void* dx;
  const float* alpha;
  float* dalpha;
  float* partial;
  int64_t n;
  int32_t dtype, act, accumulate, _pad;
} fbanet_act_bwd_params;

/* Training-mode forward of a stand-alone activation: y = act(x) on n contiguous elements, the caller keeps the pre-activation x for
 * fbanet_act_bwd_sm100 (inference applies activations inside the producing GEMM's epilogue).  alpha: device scalar, PReLU only. */
typedef struct fbanet_act_fwd_params {
  const void* x;
  void* y;
  const float* alpha;
  int64_t n;
  int32_t dtype, act;
} fbanet_act_fwd_params;

/* ---- SURVEY 8f-3, second set of backward bricks: the layers whose gradient is not a dense GEMM ---- */

/* Backward of the LeFF depthwise 3x3 (layers/locally_enhanced_feed_forward.py:39-52; forward: fbanet_dwconv3x3_sm100 with act NONE,
 * the GELU after it goes through fbanet_act_bwd_sm100).  x, dy, dx: contiguous channels-last [N,H,W,C] of `dtype`; weight: the
 * forward's packed fp32 [9][C].
 *   dx(q,c) = sum_tap w[tap][c] dy(q - (tap - 1), c),  dw[c][tap] (+)= sum_q dy(q,c) x(q + tap - 1, c),  db[c] (+)= sum_q dy(q,c)
 * dw, db: fp32 in the torch parameter layouts [C,1,3,3] / [C]; any of dx, dw, db may be NULL (not all three).
 * partial: fbanet_dwconv_bwd_blocks(N*H*W) * 10 * C floats; fixed-order reduction. */
typedef struct fbanet_dwconv_bwd_params {
  const void* x;
  const void* dy;
  const float* weight;
  void* dx;
  float* dw;
  float* db;
  float* partial;
  int32_t dtype, N, H, W, C, accumulate;
} fbanet_dwconv_bwd_params;

/* Backward of the windowed attention core (layers/window_attention.py:159-248 with the cyclic shift / window partition / shift mask
 * of layers/fba_net.py:149-238; forward: fbanet_window_attention_sm100, same token layout and the same `scale`).
 *   P = softmax(scale q k^T + bias + mask),  dv = P^T dO,  dP = dO v^T,  dS = P o (dP - rowsum(P o dP)),
 *   dq = scale dS k,  dk = dS^T (scale q),  dbias_table[index(i,j)][h] (+)= sum over windows of dS_ij
 * qkv, dqkv: [B*H*W, 3C] (q | k | v); dout: [B*H*W, C]; every element of dqkv is written.  dbias: optional fp32
 * [(2*win-1)^2][heads]; partial: fbanet_attn_bwd_partial_floats(...) floats (needed with dbias); fixed-order reduction. */
typedef struct fbanet_attn_bwd_params {
  const void* qkv;
  const void* dout;
  void* dqkv;
  const float* bias_table;
  float* dbias;
  float* partial;
  int32_t dtype, B, H, W, C, heads, win, shift;
  int32_t qkv_ld, dout_ld, dqkv_ld;
  float scale;
  int32_t accumulate, _pad;
} fbanet_attn_bwd_params;

/* Backward of the Federated-Affinity gate (blocks/federated_affinity_fusion.py:79-105; forward: fbanet_faf_gate_sm100).
 * With s_f(p) = sum_{tap,c} wsum[tap][c] feat_f(p + tap - 1, c) and g_f = sigmoid(|s_f - s_0|):
 *   ds_f = sign(s_f - s_0) g_f (1 - g_f) sum_c dgated_f feat_f  (f >= 1),  ds_0 = -sum_f ds_f,
 *   dfeat_f(q,c) = dgated_f(q,c) g_f(q) [g_0 = 1] + sum_tap wsum[tap][c] ds_f(q - (tap - 1)),
 *   dwsum[tap][c] (+)= sum_{b,f,q} feat_f(q,c) ds_f(q - (tap - 1))
 * (every output channel of temporal_attn1.weight receives dwsum; temporal_attn0 and both biases cancel out of the gate as written
 * and receive zero).  feat, dfeat: [B][F][H][W][C]; dgated: [B][H][W][F][C] (the forward's `gated` layout) of `dtype`;
 * gate: fp32 [B][F-1][H][W] as the forward stored it; score: fp32 [B][F][H][W] (s_f); dscore: fp32 [B][F][H][W] workspace / output;
 * partial: fbanet_faf_gate_bwd_blocks(B*F*H*W) * 9 * C floats. */
typedef struct fbanet_faf_gate_bwd_params {
  const void* feat;
  const void* dgated;
  const float* gate;
  const float* score;
  const float* wsum;
  void* dfeat;
  float* dscore;
  float* dwsum;
  float* partial;
  int32_t dtype, B, F, H, W, C, accumulate, _pad;
} fbanet_faf_gate_bwd_params;

/* DropPath residual (layers/drop_path.py:39-63 in its "global" mode under jax.vmap: ONE Bernoulli draw per burst and call;
 * layers/fba_net.py:245,248):  out[b][i] = skip[b][i] + scale[b] * x[b][i],  scale[b] = 0 or 1 / keep_prob, drawn by the caller
 * (fp32 device array).  skip = NULL gives the branch's backward  dx = scale[b] * dy.  A dropped burst (scale 0) contributes
 * exactly 0.  x, skip, out: B * per_burst contiguous elements of `dtype`; out may alias skip or x. */
typedef struct fbanet_drop_path_params {
  const void* x;
  const void* skip;       /* optional */
  void* out;
  const float* scale;     /* [B] fp32 */
  int64_t per_burst;
  int32_t dtype, B;
} fbanet_drop_path_params;

int fbanet_abi_version(void);
/* sizeof() of the named parameter struct as compiled, for binding self-checks; -1 if unknown */
int fbanet_abi_sizeof(const char* struct_name);
/* text of the last CUDA error seen by this library on the calling thread */
const char* fbanet_last_cuda_error(void);
/* 1 if the fused LeFF kernel takes this problem, else 0 */
int fbanet_leff_fc2_supported(const fbanet_leff_fc2_params* p);
/* 1 if the one-kernel LeFF MLP takes this problem, else 0 */
int fbanet_leff_mlp_supported(const fbanet_leff_mlp_params* p);
/* 1 if the one-pass FAF gate + fusion kernel takes this problem, else 0 */
int fbanet_faf_fuse_supported(const fbanet_faf_fuse_params* p);
/* 1 if fbanet_window_attention_sm100 runs this problem on the tcgen05 / TMEM kernel (d_h = 64, window 10, dense bias, prescaled q) */
int fbanet_window_attention_tcgen05_supported(const fbanet_attn_params* p);
/* 1 if the tcgen05 implicit-GEMM can run this problem, else 0 */
int fbanet_conv_gemm_tcgen05_supported(const fbanet_conv_params* p);

int fbanet_warp_sm100(const fbanet_warp_params* p, void* stream);
int fbanet_to_nhwc_sm100(const fbanet_to_nhwc_params* p, void* stream);
int fbanet_space_to_depth_sm100(const fbanet_s2d_params* p, void* stream);
int fbanet_head_conv_sm100(const fbanet_head_conv_params* p, void* stream);
int fbanet_assemble_sm100(const fbanet_assemble_params* p, void* stream);
int fbanet_convert_io_sm100(const fbanet_convert_io_params* p, void* stream);
int fbanet_conv_gemm_sm100(const fbanet_conv_params* p, void* stream);
int fbanet_layernorm_sm100(const fbanet_layernorm_params* p, void* stream);
int fbanet_window_attention_sm100(const fbanet_attn_params* p, void* stream);
int fbanet_dwconv3x3_sm100(const fbanet_dwconv_params* p, void* stream);
int fbanet_faf_gate_sm100(const fbanet_faf_gate_params* p, void* stream);
int fbanet_faf_fuse_sm100(const fbanet_faf_fuse_params* p, void* stream);
int fbanet_leff_fc2_sm100(const fbanet_leff_fc2_params* p, void* stream);
int fbanet_leff_mlp_sm100(const fbanet_leff_mlp_params* p, void* stream);
int fbanet_tile_divide_sm100(const fbanet_tile_params* p, void* stream);
int fbanet_tile_merge_sm100(const fbanet_tile_params* p, void* stream);
int fbanet_tile_divide_banded_sm100(const fbanet_tile_band_params* p, void* stream);
int fbanet_tile_merge_banded_sm100(const fbanet_tile_band_params* p, void* stream);
int fbanet_flow_warp_sm100(const fbanet_flow_warp_params* p, void* stream);
int fbanet_ecc_prepare_sm100(const fbanet_ecc_prepare_params* p, void* stream);
int fbanet_ecc_homography_sm100(const fbanet_ecc_params* p, void* stream);
int fbanet_train_loss_sm100(const fbanet_train_loss_params* p, void* stream);
int fbanet_adam_step_sm100(const fbanet_adam_params* p, void* stream);
int fbanet_wgrad_sm100(const fbanet_wgrad_params* p, void* stream);
int fbanet_layernorm_bwd_sm100(const fbanet_layernorm_bwd_params* p, void* stream);
int fbanet_act_bwd_sm100(const fbanet_act_bwd_params* p, void* stream);
int fbanet_act_fwd_sm100(const fbanet_act_fwd_params* p, void* stream);
int fbanet_dwconv3x3_bwd_sm100(const fbanet_dwconv_bwd_params* p, void* stream);
int fbanet_window_attention_bwd_sm100(const fbanet_attn_bwd_params* p, void* stream);
int fbanet_faf_gate_bwd_sm100(const fbanet_faf_gate_bwd_params* p, void* stream);
int fbanet_drop_path_add_sm100(const fbanet_drop_path_params* p, void* stream);
/* thread blocks (= partial rows) the two reductions above use for a problem of this size */
int fbanet_layernorm_bwd_blocks(int64_t rows);
int fbanet_act_bwd_blocks(int64_t n);
int fbanet_dwconv_bwd_blocks(int64_t pixels);            /* pixels = N*H*W */
int fbanet_faf_gate_bwd_blocks(int64_t frame_pixels);    /* frame_pixels = B*F*H*W */
/* floats of `partial` fbanet_window_attention_bwd_sm100 needs for the table gradient; -1 for an inconsistent shape */
int64_t fbanet_attn_bwd_partial_floats(int32_t B, int32_t H, int32_t W, int32_t heads, int32_t win);
/* doubles of workspace fbanet_train_loss_sm100 needs (2 per thread block) */
int64_t fbanet_train_loss_workspace_doubles(int32_t planes, int32_t H, int32_t W);

#ifdef __cplusplus
}
#endif
#endif /* FBANET_B200_H */
