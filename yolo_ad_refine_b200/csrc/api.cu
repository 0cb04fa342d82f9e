// libyad.so: error plumbing, version / device queries and the conv dispatcher.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void yad_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_pdl = -1;
int yad_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("YAD_PDL");
    g_pdl = (e && e[0] == '1') ? 1 : 0;  // measured neutral on the graph-replayed steps: off unless asked for
  }
  return g_pdl;
}

int yad_conv2d_simt(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, int dtype,
                    void* stream);
int yad_conv2d_tc(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream);
int yad_conv2d_tc_supported(const yad_tensor* x, const yad_conv_desc* d, const yad_tensor* y);
int yad_conv2d_small(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream);
int yad_conv2d_c3(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream);

extern "C" {

const char* yad_last_error(void) { return g_err; }
int yad_version(void) { return 100; }
int yad_set_pdl(int enabled) {
  const int old = yad_pdl_enabled();
  g_pdl = enabled ? 1 : 0;
  return old;
}

// sizeof of the structs that cross the C ABI, so that a binding (ctypes, cgo, JNI ...) can check its own declarations against this build
int yad_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(yad_tensor);
    case 1: return (int)sizeof(yad_epilogue);
    case 2: return (int)sizeof(yad_conv_desc);
    case 3: return (int)sizeof(yad_image_desc);
    case 4: return (int)sizeof(yad_permute_entry);
    default: return -1;
  }
}

int yad_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

int yad_conv2d(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, int dtype,
               void* stream) {
  YAD_CHECK(x && w && d && e && y && x->ptr && y->ptr, "conv2d: null argument");
  yad_epilogue eg;
  if (e->gate_h || e->gate_w) {  // separable gate (ELA_HSFPN flag=False + Multiply): 1x1 stride-1 NORMAL convolutions, rows aligned like every other operand
    YAD_CHECK(e->gate_h && e->gate_w, "conv2d: gate_h and gate_w come together");
    YAD_CHECK(d->mode == YAD_CONV_NORMAL && d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad_h == 0 && d->pad_w == 0,
              "conv2d: the separable gate epilogue covers 1x1 stride-1 convolutions only");
    YAD_CHECK(!e->mul && !e->gn_stats, "conv2d: the separable gate epilogue cannot be combined with mul / fused statistics");
    YAD_CHECK(e->gate_ld >= y->c && e->gate_ld % 8 == 0 && ((uintptr_t)e->gate_h & 15) == 0 && ((uintptr_t)e->gate_w & 15) == 0,
              "conv2d: gate rows must be 16-byte aligned with gate_ld >= cout, a multiple of 8");
    eg = *e;
    eg.gate_hm = y->h;
    eg.gate_wm = y->w;
    e = &eg;
  }
  int impl = d->impl;
  if ((impl == 0 || impl == 6) && dtype == YAD_BF16) {  // 3x3 stride-1 convolutions with <= 32 channels: tcgen05 straight on a no-swizzle patch (conv_v2.cu)
    const int r = yad_conv2d_c3(x, w, d, e, y, stream);
    if (r >= 0) return r;
    YAD_CHECK(impl != 6, "conv2d: impl 6 (small-channel tcgen05 kernel) does not cover this call");
  }
  if (impl == 0 && dtype == YAD_BF16) {  // HBM-bound 3x3 convolutions with <= 32 channels: single-pass mma.sync kernel (conv_small.cu)
    const int r = yad_conv2d_small(x, w, d, e, y, stream);
    if (r >= 0) return r;
  }
  if (impl == 0) impl = (dtype == YAD_BF16 && yad_conv2d_tc_supported(x, d, y)) ? 2 : 1;
  if (impl >= 2 && impl <= 5) {
    YAD_CHECK(dtype == YAD_BF16, "conv2d: the tcgen05 path is bf16 only");
    YAD_CHECK(yad_conv2d_tc_supported(x, d, y), "conv2d: shape not supported by the tcgen05 path");
    return yad_conv2d_tc(x, w, d, e, y, stream);
  }
  if (e->gn_stats) {  // the SIMT kernel has no fused statistics: run it, then the stand-alone statistics kernel on its output
    yad_epilogue e2 = *e;
    e2.gn_stats = nullptr;
    YAD_CHECK(e->mul == nullptr && e->add == nullptr, "conv2d: fused GroupNorm statistics cannot be combined with mul / add");
    int r = yad_conv2d_simt(x, w, d, &e2, y, dtype, stream);
    if (r) return r;
    if (e->gn_groups == 0) {  // batch statistics (train-mode BatchNorm): the batch as one image, one channel per group
      yad_tensor flat = *y;
      flat.n = 1;
      flat.h = y->n * y->h;
      return yad_gn_stats(&flat, y->c, e->gn_stats, dtype, stream);
    }
    return yad_gn_stats(y, e->gn_groups, e->gn_stats, dtype, stream);
  }
  return yad_conv2d_simt(x, w, d, e, y, dtype, stream);
}

}  // extern "C"
