// tcgen05 / TMEM / TMA implicit-GEMM convolution path (bf16 operands, fp32 accumulate in TMEM).
// Placeholder translation unit: the tensor-core kernels land here; until then every shape reports
// "not handled" and the CUDA-core path in conv_simt.cu runs.
#include "common.cuh"

extern "C" size_t rcgan_conv_wpack_bytes(const rcgan_conv_desc* d) { (void)d; return 0; }
extern "C" int rcgan_conv_wpack(const rcgan_conv_desc* d, const float* w, const float* scale_dev, void* pack, void* stream) {
  (void)d; (void)w; (void)scale_dev; (void)pack; (void)stream;
  rcgan_set_error("conv_wpack: shape has no tensor-core pack");
  return RCGAN_EUNSUPPORTED;
}
int rcgan_tc_fprop(const rcgan_conv_desc*, const void*, const void*, const float*, void*, int, int, float, cudaStream_t, int* handled) {
  *handled = 0;
  return 0;
}
int rcgan_tc_dgrad(const rcgan_conv_desc*, const void*, const void*, const float*, void*, int, int, float, int, cudaStream_t, int* handled) {
  *handled = 0;
  return 0;
}
