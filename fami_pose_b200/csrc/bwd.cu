// Backward kernels (DCN, translation warp).  Not implemented yet: loud errors.
#include "common.cuh"

namespace fami {

int dcn_bwd_launch(const fami_dcn_desc*, const float*, const float*, const float*, const float*, const float*, float*,
                   float*, float*, float*, float*, cudaStream_t) {
  set_error("fami_dcn_bwd: not implemented in this build");
  return 3;
}
int warp_translate_bwd_launch(const float*, int, const float*, const float*, int, float*, int, float*, int, int, int,
                              int, cudaStream_t) {
  set_error("fami_warp_translate_bwd: not implemented in this build");
  return 3;
}

}  // namespace fami
