// tcgen05 / TMEM implicit-GEMM convolution (placeholder until the UMMA path lands: reports "unsupported" so that the dispatcher
// keeps every shape on the SIMT kernel).
#include "common.cuh"

int yad_conv2d_tc_supported(const yad_tensor*, const yad_conv_desc*, const yad_tensor*) { return 0; }
int yad_conv2d_tc(const yad_tensor*, const void*, const yad_conv_desc*, const yad_epilogue*, const yad_tensor*, void*) {
  yad_set_error("conv2d: tcgen05 path not built");
  return 1;
}
extern "C" int yad_tc_gemm_selftest(const void*, const void*, float*, int, int, int, void*) {
  yad_set_error("tc_gemm_selftest: tcgen05 path not built");
  return 1;
}
