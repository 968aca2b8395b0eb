// conv3x3_tc.cu — tcgen05/TMEM implicit-GEMM 3x3 convolution (placeholder until the kernel lands).
#include "conv3x3.cuh"

namespace pu {
bool conv3x3_tc_supported(const Conv3x3Args&) { return false; }
int conv3x3_fwd_tc(const Conv3x3Args&, cudaStream_t) {
  set_error("conv3x3_fwd_tc: not built");
  return PU_ERR_UNSUPPORTED;
}
}  // namespace pu

extern "C" int pu_tc_available(void) { return 0; }
