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

// tap indices and 11-bit coefficients of destination index d (cv::resize, INTER_LINEAR, CV_8U); scale = src / dst in double.  The source
// coordinate is evaluated in double and rounded to float exactly as OpenCV does; the explicit _rn intrinsics keep nvcc from contracting the
// multiply-subtract into an FMA.
__device__ __forceinline__ void linear_taps(int d, int src, double scale, bool horizontal, int& s0, int& s1, int& a0, int& a1) {
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

constexpr int LB_ROWS = 4;  // output rows per thread: the horizontal taps (the double-precision part) are computed once and reused

__global__ void __launch_bounds__(128) letterbox_kernel(const yad_image_desc* __restrict__ desc, uint8_t* __restrict__ out, int out_h, int out_w,
                                                        int pad_value, int swap_rb) {
  pdl_sync();
  const int n = blockIdx.z, y_begin = blockIdx.y * LB_ROWS;
  const int x0 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (x0 >= out_w) return;
  const yad_image_desc d = desc[n];
  const uint32_t pv = (uint32_t)(pad_value & 0xff) * 0x01010101u;
  const int64_t plane = (int64_t)out_h * out_w;
  uint8_t* p = out + (int64_t)n * 3 * plane + x0;
  const bool in_x = x0 + 3 >= d.left && x0 < d.left + d.new_w;
  int o0[4], o1[4], a0[4], a1[4];  // byte offsets of the two taps inside a source row, their coefficients (0 / 0 outside the resized region)
  if (in_x) {
    const double sx = __ddiv_rn((double)d.src_w, (double)d.new_w);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int xx = x0 + i - d.left;
      o0[i] = -1;
      if (xx < 0 || xx >= d.new_w) continue;
      int s0, s1;
      linear_taps(xx, d.src_w, sx, true, s0, s1, a0[i], a1[i]);
      o0[i] = s0 * 3;
      o1[i] = s1 * 3;
    }
  }
  const double sy = __ddiv_rn((double)d.src_h, (double)d.new_h);
#pragma unroll
  for (int r = 0; r < LB_ROWS; r++) {
    const int y = y_begin + r;
    if (y >= out_h) break;
    uint32_t o[3] = {pv, pv, pv};
    const int yy = y - d.top;
    if (in_x && yy >= 0 && yy < d.new_h) {
      int sy0, sy1, b0, b1;
      linear_taps(yy, d.src_h, sy, false, sy0, sy1, b0, b1);
      const uint8_t* __restrict__ r0 = d.src + (int64_t)sy0 * d.src_pitch;
      const uint8_t* __restrict__ r1 = d.src + (int64_t)sy1 * d.src_pitch;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (o0[i] < 0) continue;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const int S0 = (int)__ldg(r0 + o0[i] + c) * a0[i] + (int)__ldg(r0 + o1[i] + c) * a1[i];
          const int S1 = (int)__ldg(r1 + o0[i] + c) * a0[i] + (int)__ldg(r1 + o1[i] + c) * a1[i];
          int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
          v = min(max(v, 0), 255);
          const int cc = swap_rb ? 2 - c : c;
          o[cc] = (o[cc] & ~(0xffu << (8 * i))) | ((uint32_t)v << (8 * i));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) *reinterpret_cast<uint32_t*>(p + c * plane + (int64_t)y * out_w) = o[c];
  }
}

__global__ void scale_boxes_kernel(float* __restrict__ det, int row_ld, const int32_t* __restrict__ count, int max_det,
                                   const yad_image_desc* __restrict__ desc) {
  pdl_sync();
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
  const dim3 grid(cdiv(out_w / 4, 128), cdiv(out_h, LB_ROWS), batch);
  YAD_LAUNCH(letterbox_kernel, grid, 128, 0, (cudaStream_t)stream, desc, out, out_h, out_w, pad_value, swap_rb);
  YAD_LAUNCH_CHECK("letterbox");
  return 0;
}

int yad_scale_boxes(float* det, int row_ld, const int32_t* count, int batch, int max_det, const yad_image_desc* desc, void* stream) {
  YAD_CHECK(det && desc, "scale_boxes: null argument");
  YAD_CHECK(row_ld >= 4, "scale_boxes: rows must hold at least x1, y1, x2, y2 (row_ld %d)", row_ld);
  YAD_CHECK(batch >= 0 && batch <= 65535 && max_det >= 0, "scale_boxes: batch %d / max_det %d out of range", batch, max_det);
  if (batch == 0 || max_det == 0) return 0;
  const dim3 grid(cdiv(max_det, 128), batch);
  YAD_LAUNCH(scale_boxes_kernel, grid, 128, 0, (cudaStream_t)stream, det, row_ld, count, max_det, desc);
  YAD_LAUNCH_CHECK("scale_boxes");
  return 0;
}

}  // extern "C"
