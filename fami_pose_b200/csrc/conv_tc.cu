// bf16 tcgen05 implicit-GEMM convolution (tensor-core arm of fami_conv2d_bn_act_fwd).
// Placeholder until the UMMA/TMA kernel lands: reports "unsupported" so callers get a loud error.
#include "common.cuh"

namespace fami {

int conv_bf16_tc_supported(const fami_conv_desc*) { return 0; }
int conv_bf16_tc_launch(const fami_conv_desc*, const void*, const void*, const float*, const float*, const void*,
                        void*, double*, cudaStream_t) {
  set_error("bf16 tensor-core convolution not built");
  return 3;
}
int64_t pack_w_bf16_elems(int, int, int, int) { return 0; }
int pack_w_bf16_launch(const float*, void*, int, int, int, int, cudaStream_t) {
  set_error("bf16 tensor-core convolution not built");
  return 3;
}

}  // namespace fami
