// Row f4 (training input pipeline): the per-image augmentations the reference runs on the host with cv2 / numpy
// (data/augment.py), as device kernels on HWC uint8 BGR images -- what its dataset hands to the transforms.
//   hsv_lut_kernel   RandomHSV.__call__ (augment.py:1345-1378): cv2.cvtColor(BGR2HSV) -> three 256-entry LUTs -> cv2.cvtColor(HSV2BGR), bit for bit
//                    OpenCV 4.x's 8-bit arithmetic: BGR2HSV is integer (hsv_shift = 12 division tables), HSV2BGR is its vectorised float path --
//                    s, v scaled by 1/255, tab2 = v * fma(-s, f, 1), tab3 = v * fma(-s, 1 - f, 1), result TRUNCATED after * 255 (the scalar tail
//                    OpenCV uses for the last w mod 16 pixels of a row ROUNDS and does not fuse: <= 1 LSB away, oracle/augment.py has both).
//   flip_kernel      RandomFlip.__call__ (:1429-1472): np.flipud / np.fliplr per image.
//   mosaic4_kernel   Mosaic._mosaic4 (:657-713): four images pasted around a random centre of a 2s x 2s canvas filled with 114.
#include "common.cuh"

namespace {

__device__ __forceinline__ void bgr2hsv_u8(int b, int g, int r, int& h, int& s, int& v) {
  // OpenCV RGB2HSV_b: sdiv_table[i] = round((255 << 12) / i), hdiv_table180[i] = round((180 << 12) / (6 i))
  const int vmax = max(max(b, g), r), vmin = min(min(b, g), r), diff = vmax - vmin;
  const int sdiv = vmax ? (int)rint((double)(255 << 12) / (double)vmax) : 0;
  const int hdiv = diff ? (int)rint((double)(180 << 12) / (6.0 * (double)diff)) : 0;
  const int vr = vmax == r ? -1 : 0, vg = vmax == g ? -1 : 0;
  s = (diff * sdiv + (1 << 11)) >> 12;
  int hh = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
  hh = (hh * hdiv + (1 << 11)) >> 12;
  h = hh + (hh < 0 ? 180 : 0);
  v = vmax;
}

__device__ __forceinline__ void hsv2bgr_u8(int h, int s, int v, int& b, int& g, int& r) {
  const float sf = __fmul_rn((float)s, 1.0f / 255.0f), vf = __fmul_rn((float)v, 1.0f / 255.0f);
  float bb, gg, rr;
  if (s == 0) {
    bb = gg = rr = vf;
  } else {
    float hf = __fmul_rn((float)h, 6.0f / 180.0f);
    while (hf >= 6.0f) hf = __fsub_rn(hf, 6.0f);
    int sector = (int)floorf(hf);
    float f = __fsub_rn(hf, (float)sector);
    if ((unsigned)sector >= 6u) { sector = 0; f = 0.f; }
    const float t0 = vf, t1 = __fmul_rn(vf, __fsub_rn(1.0f, sf)), t2 = __fmul_rn(vf, __fmaf_rn(-sf, f, 1.0f)),
                t3 = __fmul_rn(vf, __fmaf_rn(-sf, __fsub_rn(1.0f, f), 1.0f));
    // sector_data = {{1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}} -> (b, g, r)
    switch (sector) {
      case 0: bb = t1; gg = t3; rr = t0; break;
      case 1: bb = t1; gg = t0; rr = t2; break;
      case 2: bb = t3; gg = t0; rr = t1; break;
      case 3: bb = t0; gg = t2; rr = t1; break;
      case 4: bb = t0; gg = t1; rr = t3; break;
      default: bb = t2; gg = t1; rr = t0; break;
    }
  }
  b = min(max((int)__fmul_rn(bb, 255.0f), 0), 255);
  g = min(max((int)__fmul_rn(gg, 255.0f), 0), 255);
  r = min(max((int)__fmul_rn(rr, 255.0f), 0), 255);
}

// one thread = 4 consecutive pixels (12 bytes, three aligned 32-bit words); luts: [n][3][256]
__global__ void hsv_lut_kernel(uint8_t* __restrict__ img, int64_t px_per_img, int n, const uint8_t* __restrict__ luts) {
  pdl_sync();
  __shared__ uint8_t lut[3 * 256];
  const int im = blockIdx.y;
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = luts[(int64_t)im * 768 + i];
  __syncthreads();
  uint8_t* base = img + (int64_t)im * px_per_img * 3;
  const int64_t groups = (px_per_img + 3) / 4;
  for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p0 = gi * 4;
    uint8_t px[12];
    const bool full = p0 + 4 <= px_per_img && (((uintptr_t)(base + p0 * 3)) & 3) == 0;
    if (full) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(base + p0 * 3);
      uint32_t a = w[0], bq = w[1], c = w[2];
      memcpy(px, &a, 4); memcpy(px + 4, &bq, 4); memcpy(px + 8, &c, 4);
    } else {
      for (int k = 0; k < 12; k++) px[k] = (p0 * 3 + k < px_per_img * 3) ? base[p0 * 3 + k] : 0;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      int h, s, v, b, g, r;
      bgr2hsv_u8(px[3 * q], px[3 * q + 1], px[3 * q + 2], h, s, v);
      hsv2bgr_u8(lut[h], lut[256 + s], lut[512 + v], b, g, r);
      px[3 * q] = (uint8_t)b; px[3 * q + 1] = (uint8_t)g; px[3 * q + 2] = (uint8_t)r;
    }
    if (full) {
      uint32_t a, bq, c;
      memcpy(&a, px, 4); memcpy(&bq, px + 4, 4); memcpy(&c, px + 8, 4);
      uint32_t* w = reinterpret_cast<uint32_t*>(base + p0 * 3);
      w[0] = a; w[1] = bq; w[2] = c;
    } else {
      for (int k = 0; k < 12; k++)
        if (p0 * 3 + k < px_per_img * 3) base[p0 * 3 + k] = px[k];
    }
  }
}

// dst[n][y][x] = src[n][ud ? h-1-y : y][lr ? w-1-x : x]; flags[n] bit 0 = up-down, bit 1 = left-right
__global__ void flip_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int h, int w, const uint8_t* __restrict__ flags) {
  pdl_sync();
  const int64_t total = (int64_t)n * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h), im = (int)(i / ((int64_t)w * h));
    const int f = flags[im];
    const int sy = (f & 1) ? h - 1 - y : y, sx = (f & 2) ? w - 1 - x : x;
    const uint8_t* s = src + (((int64_t)im * h + sy) * w + sx) * 3;
    uint8_t* d = dst + i * 3;
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
  }
}

__global__ void mosaic4_kernel(uint8_t* __restrict__ canvas, int s2, int n_out, const yad_mosaic_desc* __restrict__ desc) {
  pdl_sync();
  const int64_t total = (int64_t)n_out * s2 * s2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % s2), y = (int)((i / s2) % s2), im = (int)(i / ((int64_t)s2 * s2));
    uint8_t b = 114, g = 114, r = 114;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const yad_mosaic_desc d = desc[im * 4 + q];
      if (x >= d.x1a && x < d.x2a && y >= d.y1a && y < d.y2a) {
        const uint8_t* sp = reinterpret_cast<const uint8_t*>(d.src) + ((int64_t)(d.y1b + y - d.y1a) * d.src_w + (d.x1b + x - d.x1a)) * 3;
        b = sp[0]; g = sp[1]; r = sp[2];
      }
    }
    uint8_t* o = canvas + i * 3;
    o[0] = b; o[1] = g; o[2] = r;
  }
}

int aug_blocks(int64_t n) {
  int64_t g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

extern "C" {

int yad_hsv_lut(void* img_u8_hwc_bgr, int n, int h, int w, const void* luts_dev, void* stream) {
  YAD_CHECK(img_u8_hwc_bgr && luts_dev && n > 0 && h > 0 && w > 0, "hsv_lut: bad arguments");
  const int64_t px = (int64_t)h * w;
  int gx = aug_blocks((px + 3) / 4);
  if (gx > 148 * 4) gx = 148 * 4;
  YAD_LAUNCH(hsv_lut_kernel, dim3(gx, n), 256, 0, (cudaStream_t)stream, (uint8_t*)img_u8_hwc_bgr, px, n, (const uint8_t*)luts_dev);
  YAD_LAUNCH_CHECK("hsv_lut");
  return 0;
}

int yad_flip(const void* src, void* dst, int n, int h, int w, const void* flags_dev, void* stream) {
  YAD_CHECK(src && dst && src != dst && flags_dev && n > 0 && h > 0 && w > 0, "flip: bad arguments (out of place only)");
  YAD_LAUNCH(flip_kernel, aug_blocks((int64_t)n * h * w), 256, 0, (cudaStream_t)stream, (const uint8_t*)src, (uint8_t*)dst, n, h, w, (const uint8_t*)flags_dev);
  YAD_LAUNCH_CHECK("flip");
  return 0;
}

int yad_mosaic4(void* canvas, int s2, int n_out, const yad_mosaic_desc* desc_dev, void* stream) {
  YAD_CHECK(canvas && desc_dev && s2 > 0 && n_out > 0, "mosaic4: bad arguments");
  YAD_LAUNCH(mosaic4_kernel, aug_blocks((int64_t)n_out * s2 * s2), 256, 0, (cudaStream_t)stream, (uint8_t*)canvas, s2, n_out, desc_dev);
  YAD_LAUNCH_CHECK("mosaic4");
  return 0;
}

}  // extern "C"
