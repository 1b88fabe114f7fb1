// placeholder until the tcgen05 implicit-GEMM lands; reports "unsupported".
#include "common.cuh"
namespace fbanet {
int conv_gemm_tc_supported(const fbanet_conv_params*) { return 0; }
int conv_gemm_tc_launch(const fbanet_conv_params*, cudaStream_t) { return FBANET_E_UNSUPPORTED; }
}  // namespace fbanet
