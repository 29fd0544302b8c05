// tcgen05 implicit-GEMM convolution (placeholder until the kernel lands: reports "unsupported").
#include "common.cuh"
namespace cgat {
int conv_tc_supported(const cgat_conv_desc*, int) { return 0; }
int conv_fprop_tc_launch(const cgat_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t) {
  return fail(CGAT_EUNSUPPORTED, "tcgen05 conv not built");
}
int conv_dgrad_tc_launch(const cgat_conv_desc*, const void*, const void*, void*, cudaStream_t) {
  return fail(CGAT_EUNSUPPORTED, "tcgen05 conv not built");
}
int conv_wgrad_tc_launch(const cgat_conv_desc*, const void*, const void*, float*, float*, cudaStream_t) {
  return fail(CGAT_EUNSUPPORTED, "tcgen05 conv not built");
}
}  // namespace cgat
