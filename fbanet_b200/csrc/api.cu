// C-ABI glue: error reporting, ABI self-description, dispatch between implementations.
#include <string.h>

#include "common.cuh"

namespace fbanet {

static thread_local char g_err[256] = "";

void set_last_error(cudaError_t e) {
  const char* s = cudaGetErrorString(e);
  strncpy(g_err, s ? s : "unknown", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return FBANET_OK;
  set_last_error(e);
  return FBANET_E_LAUNCH;
}

int conv_gemm_validate(const fbanet_conv_params* p);
int conv_gemm_simt_launch(const fbanet_conv_params* p, cudaStream_t stream);
int conv_gemm_tc_supported(const fbanet_conv_params* p);
int conv_gemm_tc_launch(const fbanet_conv_params* p, cudaStream_t stream);
int window_attention_validate(const fbanet_attn_params* p);
int window_attention_simt_launch(const fbanet_attn_params* p, cudaStream_t s);
int window_attention_tc_supported(const fbanet_attn_params* p);
int window_attention_tc_launch(const fbanet_attn_params* p, cudaStream_t s);
int window_attention_dh16_supported(const fbanet_attn_params* p);
int window_attention_dh16_launch(const fbanet_attn_params* p, cudaStream_t s);
int window_attention_tcgen05_supported(const fbanet_attn_params* p);
int window_attention_tcgen05_launch(const fbanet_attn_params* p, cudaStream_t s);

}  // namespace fbanet

using namespace fbanet;

extern "C" int fbanet_abi_version(void) { return FBANET_ABI_VERSION; }

extern "C" int fbanet_abi_sizeof(const char* n) {
  if (!n) return -1;
#define SZ(T) if (!strcmp(n, #T)) return (int)sizeof(T)
  SZ(fbanet_src);
  SZ(fbanet_conv_params);
  SZ(fbanet_warp_params);
  SZ(fbanet_to_nhwc_params);
  SZ(fbanet_s2d_params);
  SZ(fbanet_head_conv_params);
  SZ(fbanet_assemble_params);
  SZ(fbanet_convert_io_params);
  SZ(fbanet_layernorm_params);
  SZ(fbanet_attn_params);
  SZ(fbanet_dwconv_params);
  SZ(fbanet_faf_gate_params);
  SZ(fbanet_faf_fuse_params);
  SZ(fbanet_leff_fc2_params);
  SZ(fbanet_leff_mlp_params);
  SZ(fbanet_tile_params);
  SZ(fbanet_tile_band_params);
  SZ(fbanet_flow_warp_params);
  SZ(fbanet_ecc_prepare_params);
  SZ(fbanet_ecc_params);
  SZ(fbanet_train_loss_params);
  SZ(fbanet_adam_params);
  SZ(fbanet_wgrad_params);
  SZ(fbanet_layernorm_bwd_params);
  SZ(fbanet_act_bwd_params);
  SZ(fbanet_act_fwd_params);
  SZ(fbanet_dwconv_bwd_params);
  SZ(fbanet_attn_bwd_params);
  SZ(fbanet_faf_gate_bwd_params);
  SZ(fbanet_drop_path_params);
#undef SZ
  return -1;
}

extern "C" const char* fbanet_last_cuda_error(void) { return g_err; }

extern "C" int fbanet_conv_gemm_tcgen05_supported(const fbanet_conv_params* p) {
  if (conv_gemm_validate(p) != FBANET_OK) return 0;
  return conv_gemm_tc_supported(p);
}

extern "C" int fbanet_conv_gemm_sm100(const fbanet_conv_params* p, void* stream) {
  int rc = conv_gemm_validate(p);
  if (rc != FBANET_OK) return rc;
  const bool tc_ok = conv_gemm_tc_supported(p) != 0;
  if (p->impl == FBANET_IMPL_TCGEN05 && !tc_ok) return FBANET_E_UNSUPPORTED;
  if (tc_ok && p->impl != FBANET_IMPL_SIMT) return conv_gemm_tc_launch(p, (cudaStream_t)stream);
  if (p->src_s2d || p->store_mode == FBANET_STORE_NHWC_F32 || p->ln_stats || p->ln_gamma || p->store_f16) return FBANET_E_UNSUPPORTED;  // TMA / tensor-core path only
  return conv_gemm_simt_launch(p, (cudaStream_t)stream);
}

extern "C" int fbanet_window_attention_tcgen05_supported(const fbanet_attn_params* p) {
  if (window_attention_validate(p) != FBANET_OK) return 0;
  return window_attention_tcgen05_supported(p);
}

extern "C" int fbanet_window_attention_sm100(const fbanet_attn_params* p, void* stream) {
  int rc = window_attention_validate(p);
  if (rc != FBANET_OK) return rc;
  const bool tc_ok = window_attention_tc_supported(p) != 0;
  if (p->impl == FBANET_IMPL_TCGEN05 && !tc_ok) return FBANET_E_UNSUPPORTED;
  if (tc_ok && p->impl != FBANET_IMPL_SIMT) {
    if (window_attention_tcgen05_supported(p)) return window_attention_tcgen05_launch(p, (cudaStream_t)stream);   // d_h = 64: tcgen05 + TMEM
    if (window_attention_dh16_supported(p)) return window_attention_dh16_launch(p, (cudaStream_t)stream);
    return window_attention_tc_launch(p, (cudaStream_t)stream);
  }
  if (p->q_prescaled) return FBANET_E_UNSUPPORTED;  // the fp32 / SIMT kernel applies `scale` itself
  return window_attention_simt_launch(p, (cudaStream_t)stream);
}
