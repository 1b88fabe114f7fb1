// placeholder until the tensor-core window attention lands; reports "unsupported".
#include "common.cuh"
namespace fbanet {
int window_attention_tc_supported(const fbanet_attn_params*) { return 0; }
int window_attention_tc_launch(const fbanet_attn_params*, cudaStream_t) { return FBANET_E_UNSUPPORTED; }
}  // namespace fbanet
