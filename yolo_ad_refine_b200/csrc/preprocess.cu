// Image pre-processing and box post-scaling on the device (include/yad.h "f1").
//
// letterbox_kernel: LetterBox.__call__ (data/augment.py:1475-1600) + BasePredictor.preprocess' host half (engine/predictor.py:127-129) for a ragged
// batch of HWC uint8 images in ONE launch.  cv2.resize(INTER_LINEAR) on 8-bit images is fixed-point: per axis two taps with coefficients rounded
// to 1/2048 (horizontal taps outside the row are snapped to the border with the fraction zeroed; vertical taps are only clipped), horizontal
// pass in int32, vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  Those integer steps are reproduced exactly, so the
// uint8 tensor is byte-identical to the reference's.  HBM-bound: reads the source pixels once (L1/L2 serve the 2x2 tap overlap), writes 3 bytes
// per output pixel as 32-bit words (4 consecutive pixels of one colour plane per thread => 128-byte lines per warp).
//
// scale_boxes_kernel: scale_boxes + clip_boxes (utils/ops.py:88-123, 315-334) on the batched NMS output, fp32 with the reference's operation order.
#include "common.cuh"

namespace {

// tap indices and 11-bit coefficients of destination index d (cv::resize, INTER_LINEAR, CV_8U).  The source coordinate is evaluated in double
// and rounded to float exactly as OpenCV does; the explicit _rn intrinsics keep nvcc from contracting the multiply-subtract into an FMA.
__device__ __forceinline__ void linear_taps(int d, int src, int dst, bool horizontal, int& s0, int& s1, int& a0, int& a1) {
  const double scale = __ddiv_rn((double)src, (double)dst);
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (horizontal) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  s0 = min(max(s, 0), src - 1);
  s1 = min(max(s + 1, 0), src - 1);
}

__global__ void __launch_bounds__(128) letterbox_kernel(const yad_image_desc* __restrict__ desc, uint8_t* __restrict__ out, int out_h, int out_w,
                                                        int pad_value, int swap_rb) {
  const int n = blockIdx.z, y = blockIdx.y;
  const int x0 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (x0 >= out_w) return;
  const yad_image_desc d = desc[n];
  const uint32_t pv = (uint32_t)(pad_value & 0xff) * 0x01010101u;
  uint32_t o[3] = {pv, pv, pv};
  const int yy = y - d.top;
  if (yy >= 0 && yy < d.new_h && x0 + 3 >= d.left && x0 < d.left + d.new_w) {
    int sy0, sy1, b0, b1;
    linear_taps(yy, d.src_h, d.new_h, false, sy0, sy1, b0, b1);
    const uint8_t* __restrict__ r0 = d.src + (int64_t)sy0 * d.src_pitch;
    const uint8_t* __restrict__ r1 = d.src + (int64_t)sy1 * d.src_pitch;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int xx = x0 + i - d.left;
      if (xx < 0 || xx >= d.new_w) continue;
      int sx0, sx1, a0, a1;
      linear_taps(xx, d.src_w, d.new_w, true, sx0, sx1, a0, a1);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const int S0 = (int)__ldg(r0 + sx0 * 3 + c) * a0 + (int)__ldg(r0 + sx1 * 3 + c) * a1;
        const int S1 = (int)__ldg(r1 + sx0 * 3 + c) * a0 + (int)__ldg(r1 + sx1 * 3 + c) * a1;
        int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        const int cc = swap_rb ? 2 - c : c;
        o[cc] = (o[cc] & ~(0xffu << (8 * i))) | ((uint32_t)v << (8 * i));
      }
    }
  }
  const int64_t plane = (int64_t)out_h * out_w;
  uint8_t* p = out + (int64_t)n * 3 * plane + (int64_t)y * out_w + x0;
#pragma unroll
  for (int c = 0; c < 3; c++) *reinterpret_cast<uint32_t*>(p + c * plane) = o[c];
}

__global__ void scale_boxes_kernel(float* __restrict__ det, int row_ld, const int32_t* __restrict__ count, int max_det,
                                   const yad_image_desc* __restrict__ desc) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = count ? min(count[b], max_det) : max_det;
  if (i >= k) return;
  const yad_image_desc d = desc[b];
  float* r = det + ((int64_t)b * max_det + i) * row_ld;
  const float w0 = (float)d.src_w, h0 = (float)d.src_h;
  // boxes[..., 0::2] -= pad[0]; boxes[..., 1::2] -= pad[1]; boxes /= gain; clamp(0, shape)   (utils/ops.py:114-123, 327-331)
  const float x1 = __fdiv_rn(__fsub_rn(r[0], d.pad_x), d.gain), y1 = __fdiv_rn(__fsub_rn(r[1], d.pad_y), d.gain);
  const float x2 = __fdiv_rn(__fsub_rn(r[2], d.pad_x), d.gain), y2 = __fdiv_rn(__fsub_rn(r[3], d.pad_y), d.gain);
  r[0] = fminf(fmaxf(x1, 0.f), w0);
  r[1] = fminf(fmaxf(y1, 0.f), h0);
  r[2] = fminf(fmaxf(x2, 0.f), w0);
  r[3] = fminf(fmaxf(y2, 0.f), h0);
}

}  // namespace

extern "C" {

int yad_letterbox(const yad_image_desc* desc, int batch, uint8_t* out, int out_h, int out_w, int pad_value, int swap_rb, void* stream) {
  YAD_CHECK(desc && out, "letterbox: null argument");
  YAD_CHECK(batch >= 0 && batch <= 65535 && out_h > 0 && out_h <= 65535, "letterbox: batch %d / out_h %d outside the grid limits", batch, out_h);
  YAD_CHECK(out_w > 0 && out_w % 4 == 0, "letterbox: out_w %d must be a positive multiple of 4", out_w);
  YAD_CHECK(((uintptr_t)out & 3) == 0, "letterbox: out must be 4-byte aligned");
  if (batch == 0) return 0;
  const dim3 grid(cdiv(out_w / 4, 128), out_h, batch);
  letterbox_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(desc, out, out_h, out_w, pad_value, swap_rb);
  YAD_LAUNCH_CHECK("letterbox");
  return 0;
}

int yad_scale_boxes(float* det, int row_ld, const int32_t* count, int batch, int max_det, const yad_image_desc* desc, void* stream) {
  YAD_CHECK(det && desc, "scale_boxes: null argument");
  YAD_CHECK(row_ld >= 4, "scale_boxes: rows must hold at least x1, y1, x2, y2 (row_ld %d)", row_ld);
  YAD_CHECK(batch >= 0 && batch <= 65535 && max_det >= 0, "scale_boxes: batch %d / max_det %d out of range", batch, max_det);
  if (batch == 0 || max_det == 0) return 0;
  const dim3 grid(cdiv(max_det, 128), batch);
  scale_boxes_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(det, row_ld, count, max_det, desc);
  YAD_LAUNCH_CHECK("scale_boxes");
  return 0;
}

}  // extern "C"
