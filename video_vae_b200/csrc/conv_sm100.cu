// Tensor-core implicit-GEMM conv3d (bf16).  Placeholder dispatch: reports "unsupported" until the kernel lands, so
// every convolution currently runs on the generic functor GEMM (conv_simt.cu).
#include "common.cuh"

namespace vvae {
int conv_tc_supported(const vvae_conv_args&, int) { return 0; }
int conv_tc_launch(const vvae_conv_args&, int, cudaStream_t) {
  set_error("conv3d tensor-core path not available");
  return VVAE_ERR_UNSUPPORTED;
}
}  // namespace vvae
