// placeholder, replaced below
#include "common.cuh"
#include "kernels.h"
namespace dx {
bool conv_gemm_tc_supported(const ConvGemmArgs&) { return false; }
int conv_gemm_tc(const ConvGemmArgs&, cudaStream_t) { set_last_error("tcgen05 path not built"); return DX_ERR_UNSUPPORTED; }
bool conv_wgrad_tc_supported(const ConvWgradArgs&) { return false; }
int conv_wgrad_tc(const ConvWgradArgs&, cudaStream_t) { set_last_error("tcgen05 path not built"); return DX_ERR_UNSUPPORTED; }
size_t conv_wgrad_tc_workspace(const ConvWgradArgs&) { return 0; }
}
