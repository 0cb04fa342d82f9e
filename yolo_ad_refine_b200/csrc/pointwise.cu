// Memory-bound refinement-block kernels (NHWC, 128-bit vector access on 8 consecutive channels, smem staging and
// warp-shuffle reductions where a reduction is involved).  Each entry point cites the reference operator it replaces.
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ int64_t pix_off(const yad_tensor& t, int n, int y, int x) { return ((int64_t)(n * t.h + y) * t.w + x) * t.ld; }

// ------------------------------------------------------------------------------------------------------------------
// NCHW fp32 image -> NHWC (zero-filled tail channels)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int c_src, yad_tensor y) {
  pdl_sync();
  int64_t hw = (int64_t)y.h * y.w, total = (int64_t)y.n * hw;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = p / hw, r = p - n * hw;
    T* o = reinterpret_cast<T*>(y.ptr) + p * y.ld;
    for (int c0 = 0; c0 < y.c; c0 += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = (c0 + i < c_src) ? src[(n * c_src + c0 + i) * hw + r] : 0.f;
      store8(o + c0, v);
    }
  }
}

// uint8 NCHW image -> NHWC * scale (the reference does H2D of the uint8 batch and then /255 on the device, engine/predictor.py:129-133)
template <typename T>
__global__ void u8_to_nhwc_kernel(const uint8_t* __restrict__ src, int c_src, float scale, yad_tensor y) {
  pdl_sync();
  int64_t hw = (int64_t)y.h * y.w, total = (int64_t)y.n * hw;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = p / hw, r = p - n * hw;
    T* o = reinterpret_cast<T*>(y.ptr) + p * y.ld;
    for (int c0 = 0; c0 < y.c; c0 += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = (c0 + i < c_src) ? (float)src[(n * c_src + c0 + i) * hw + r] * scale : 0.f;
      store8(o + c0, v);
    }
  }
}

// packed 8-channel vectors: loads stay in their storage format (4 registers for bf16) until they are consumed, so that several can be in flight
template <typename T> struct Raw8;
template <> struct Raw8<bf16> { uint4 u; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ void raw_load(const bf16* p, Raw8<bf16>& r) { r.u = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void raw_load(const float* p, Raw8<float>& r) {
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void raw_unpack(const Raw8<bf16>& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; i++) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void raw_unpack(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ float tanh_approx_pw(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16 storage: SiLU / sigmoid on one MUFU (tanh.approx, 2^-11 relative -- below the bf16 rounding of the result); fp32 storage keeps apply_act_n
template <typename T, int N = 8>
__device__ __forceinline__ void apply_act_store_n(float (&v)[N], int act) {
  if (sizeof(T) == 2 && act == YAD_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < N; i++) { const float h = 0.5f * v[i]; v[i] = fmaf(h, tanh_approx_pw(h), h); }
  } else if (sizeof(T) == 2 && act == YAD_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = fmaf(0.5f, tanh_approx_pw(0.5f * v[i]), 0.5f);
  } else {
    apply_act_n<N>(v, act);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Stem: layer 0 of the model (Conv 3->16, k3 s2 p1, BN folded, SiLU) straight from the NCHW image (uint8 or fp32) to NHWC.
// Fuses the predictor's uint8 -> float / 255 (engine/predictor.py:129-133, the scale is folded into the weights), the NCHW -> NHWC
// change and nn/modules/conv.py:52-54 into one pass: reads 3 bytes per input pixel instead of a padded 8-channel bf16 copy.
// One thread per output pixel, COUT accumulators in registers, weights in shared memory.
// ------------------------------------------------------------------------------------------------------------------
template <typename T, typename IN, int COUT>
__global__ void __launch_bounds__(128) stem_conv_kernel(const IN* __restrict__ img, int n, int h, int w, int cin, const float* __restrict__ wgt,
                                                        const float* __restrict__ bias, int act, yad_tensor y) {
  pdl_sync();
  // grid (ceil(wo / 128), ho, n): one CTA = 128 consecutive output pixels of one output row.  The 3 input rows x (2*128 + 1) columns x cin
  // it needs are staged in shared memory with coalesced loads (zero-filled outside the image); weights are broadcast reads.
  extern __shared__ float sw[];  // weights [9*cin][COUT], bias [COUT], input tile [cin][3][257]
  const int K = 9 * cin;
  float* sb = sw + K * COUT;
  float* tile = sb + COUT;
  constexpr int TW = 2 * 128 + 1;
  for (int i = threadIdx.x; i < K * COUT; i += blockDim.x) {
    int co = i % COUT, k = i / COUT;
    sw[i] = wgt[co * K + k];
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  const int ho = y.h, wo = y.w;
  const int b = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * 128;
  const int ix0 = 2 * ox0 - 1, iy0 = 2 * oy - 1;
  if (sizeof(IN) == 1 && (w & 3) == 0 && ((uintptr_t)img & 3) == 0) {
    // uint8 images: aligned 32-bit loads of 4 pixels (the tile starts at ix0 = 2 ox0 - 1; words start at 2 ox0 - 4), 66 words per row instead
    // of 257 byte loads -- the kernel was issue-bound on this staging loop
    constexpr int WPR = 66;
    const int wx0 = 2 * ox0 - 4;
    for (int i = threadIdx.x; i < cin * 3 * WPR; i += blockDim.x) {
      const int wd = i % WPR, rr = i / WPR, r = rr % 3, ci = rr / 3;
      const int iy = iy0 + r, ixw = wx0 + 4 * wd;
      uint32_t word = 0u;
      if (iy >= 0 && iy < h && ixw >= 0 && ixw < w)
        word = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(img) + ((int64_t)(b * cin + ci) * h + iy) * w + ixw);
      float* dst = tile + rr * TW;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int col = 4 * wd + k - 3;  // tile column of pixel ixw + k
        if (col >= 0 && col < TW) dst[col] = (float)((word >> (8 * k)) & 0xffu);
      }
    }
  } else {
    for (int i = threadIdx.x; i < cin * 3 * TW; i += blockDim.x) {
      const int col = i % TW, r = (i / TW) % 3, ci = i / (3 * TW);
      const int iy = iy0 + r, ix = ix0 + col;
      float v = 0.f;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = (float)img[((int64_t)(b * cin + ci) * h + iy) * w + ix];
      tile[i] = v;
    }
  }
  __syncthreads();
  const int ox = ox0 + threadIdx.x;
  if (ox >= wo) return;
  float acc[COUT];
#pragma unroll
  for (int i = 0; i < COUT; i++) acc[i] = sb[i];
  for (int ci = 0; ci < cin; ci++) {
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const float v = tile[(ci * 3 + ky) * TW + 2 * threadIdx.x + kx];
        const float4* wr = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * cin + ci) * COUT);  // K index = tap * cin + ci
#pragma unroll
        for (int i4 = 0; i4 < COUT / 4; i4++) {
          const float4 w4 = wr[i4];
          acc[4 * i4] = fmaf(v, w4.x, acc[4 * i4]); acc[4 * i4 + 1] = fmaf(v, w4.y, acc[4 * i4 + 1]);
          acc[4 * i4 + 2] = fmaf(v, w4.z, acc[4 * i4 + 2]); acc[4 * i4 + 3] = fmaf(v, w4.w, acc[4 * i4 + 3]);
        }
      }
    }
  }
  T* o = reinterpret_cast<T*>(y.ptr) + (((int64_t)b * ho + oy) * wo + ox) * y.ld;
#pragma unroll
  for (int c0 = 0; c0 < COUT; c0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = acc[c0 + i];
    apply_act_n<8>(v, act);
    store8(o + c0, v);
  }
}

// Tensor-core stem for uint8 images with 3 channels and bf16 storage: the same Conv(3 -> 16, k3 s2 p1) + bias + activation as an implicit GEMM
// M = 128 output pixels of one row per CTA (4 warps x 2 m16 tiles), N = 16 (2 n8 tiles), K = 27 zero-padded to 32 (2 k16 steps) on
// mma.sync.m16n8k16 (f16 x f16 -> f32).  uint8 pixels are exact in f16 and reach the A fragments without a single I2F: two bytes are merged with
// the f16 magic 0x6400 (= 1024.0, ulp 1) and one packed HSUB2 removes the 1024.  The weights arrive with the predictor's 1/255 folded in (fp32);
// they are scaled back by 255 before the f16 rounding (keeps them far from the f16 subnormals) and the 1/255 moves to the epilogue.  The SIMT
// kernel above spends 432 FFMA + 135 shared-memory loads per output pixel; this one 8 MMAs per 32 pixels, so the kernel is left with its
// memory traffic (3 bytes in, 32 bytes out per output pixel).  The raw uint8 tile (9 rows x 264 bytes) is staged with aligned 32-bit loads; the
// outputs leave through a per-warp staging tile as 128-bit coalesced stores.
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t u8x2_to_f16x2(uint32_t lo, uint32_t hi) {
  uint32_t m = (lo | (hi << 16)) | 0x64006400u, r;
  asm("sub.f16x2 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(0x64006400u));
  return r;
}
constexpr int STEM_ROWS = 8;  // output rows per CTA: weights / offsets set up once, the next row's tile streams in (cp.async) under the MMAs
__global__ void __launch_bounds__(128) stem_tc_kernel(const uint8_t* __restrict__ img, int h, int w, const float* __restrict__ wgt,
                                                      const float* __restrict__ bias, int act, yad_tensor y) {
  // tile row = input pixels [2 ox0 - 16, 2 ox0 + 256) of one (channel, input row): 17 aligned 16-byte chunks (w % 16 == 0, host-checked), each
  // entirely inside or outside the image, so the staging is one cp.async (zero-filling when outside) per chunk
  constexpr int CPR = 17, TWB = CPR * 16, KR = 27, TILE = 9 * TWB;
  __shared__ __align__(16) uint8_t tile[2][TILE];    // [ci * 3 + r][272 bytes] of input rows 2 oy - 1 + r
  __shared__ __align__(16) uint32_t stg[4][32 * 8];  // per warp: 32 pixels x 16 bf16
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int ho = y.h, wo = y.w;
  const int b = blockIdx.z, oy0 = blockIdx.y * STEM_ROWS, ox0 = blockIdx.x * 128;
  const int rows = min(STEM_ROWS, ho - oy0);
  // B fragments (weights, constant per thread) and this thread's 8 tile columns: K index = (ky * 3 + kx) * 3 + ci reads byte
  // (ci * 3 + ky) * 272 + 2 m + 15 + kx of the tile; m = warp * 32 + g is folded in, (m16 tile, row half) stay immediate offsets
  uint32_t bfr[2][2][2];
  const uint8_t* ap[2][2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ks++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int k0 = ks * 16 + 2 * q + 8 * j;
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        const int co = nt * 8 + g;
        const float w0 = k0 < KR ? wgt[co * KR + k0] * 255.0f : 0.f, w1 = k0 + 1 < KR ? wgt[co * KR + k0 + 1] * 255.0f : 0.f;
        const __half2 hw = __floats2half2_rn(w0, w1);
        bfr[nt][ks][j] = *reinterpret_cast<const uint32_t*>(&hw);
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int k = k0 + e;  // columns >= 27 meet zero weights: any staged byte will do
        const int tap = k < KR ? k / 3 : 0, ci = k < KR ? k - 3 * tap : 0, ky = tap / 3, kx = tap - 3 * ky;
        ap[ks][j][e] = &tile[0][0] + (ci * 3 + ky) * TWB + 15 + kx + 2 * (warp * 32 + g);
      }
    }
  float bv[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; nt++) { bv[nt][0] = bias ? bias[nt * 8 + 2 * q] : 0.f; bv[nt][1] = bias ? bias[nt * 8 + 2 * q + 1] : 0.f; }
  // staging role of this thread: chunks tid and tid + 128 of the 9 x 17
  const int c0r = tid / CPR, c0c = tid - c0r * CPR, c1r = (tid + 128) / CPR, c1c = (tid + 128) - c1r * CPR;
  const bool has1 = tid + 128 < 9 * CPR;
  const int wx0 = 2 * ox0 - 16;
  pdl_sync();
  auto stage_chunk = [&](int rr, int cc, int iy0, int buf) {
    const int ci = rr / 3, r = rr - 3 * ci, iy = iy0 + r, ix = wx0 + 16 * cc;
    const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
    const uint8_t* src = in ? img + ((int64_t)(b * 3 + ci) * h + iy) * w + ix : img;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tile[buf][rr * TWB + cc * 16]);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
  };
  auto stage = [&](int oy, int buf) {
    stage_chunk(c0r, c0c, 2 * oy - 1, buf);
    if (has1) stage_chunk(c1r, c1c, 2 * oy - 1, buf);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // m16 tiles of this warp that hold at least one real pixel (the last CTA of a row is ragged)
  const int mt_cnt = min(2, max(0, (wo - ox0 - warp * 32 + 15) / 16));
  stage(oy0, 0);
  for (int r = 0; r < rows; r++) {
    const int oy = oy0 + r;
    if (r + 1 < rows) {
      stage(oy + 1, (r + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int boff = (r & 1) * TILE;
    for (int mt = 0; mt < mt_cnt; mt++) {
      uint32_t a[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ks++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const uint8_t* p0 = ap[ks][j][0] + boff + mt * 32;
          const uint8_t* p1 = ap[ks][j][1] + boff + mt * 32;
          a[ks][2 * j] = u8x2_to_f16x2(p0[0], p1[0]);       // row g
          a[ks][2 * j + 1] = u8x2_to_f16x2(p0[16], p1[16]);  // row g + 8
        }
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma_f16_16816(c, a[0], bfr[nt][0][0], bfr[nt][0][1]);
        mma_f16_16816(c, a[1], bfr[nt][1][0], bfr[nt][1][1]);
        float v[4];
        v[0] = fmaf(c[0], 1.0f / 255.0f, bv[nt][0]); v[1] = fmaf(c[1], 1.0f / 255.0f, bv[nt][1]);
        v[2] = fmaf(c[2], 1.0f / 255.0f, bv[nt][0]); v[3] = fmaf(c[3], 1.0f / 255.0f, bv[nt][1]);
        apply_act_store_n<bf16, 4>(v, act);
        const __nv_bfloat162 r0 = __floats2bfloat162_rn(v[0], v[1]), r1 = __floats2bfloat162_rn(v[2], v[3]);
        stg[warp][(mt * 16 + g) * 8 + nt * 4 + q] = *reinterpret_cast<const uint32_t*>(&r0);
        stg[warp][(mt * 16 + g + 8) * 8 + nt * 4 + q] = *reinterpret_cast<const uint32_t*>(&r1);
      }
    }
    __syncwarp();
    bf16* orow = reinterpret_cast<bf16*>(y.ptr) + (((int64_t)b * ho + oy) * wo + ox0 + warp * 32) * y.ld;
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const int chunk = lane + 32 * rr, px = chunk >> 1;
      if (ox0 + warp * 32 + px < wo)
        *reinterpret_cast<uint4*>(orow + (int64_t)px * y.ld + (chunk & 1) * 8) = *reinterpret_cast<const uint4*>(&stg[warp][px * 8 + (chunk & 1) * 4]);
    }
    __syncthreads();  // every warp is done with tile[r & 1] (the row after next lands there) and with its staging tile
  }
}

// ------------------------------------------------------------------------------------------------------------------
// GroupNorm statistics + apply
// ------------------------------------------------------------------------------------------------------------------
// grid (chunks, n).  stats[n][g][2] += (sum, sumsq) in double.
template <typename T>
__global__ void gn_stats_kernel(yad_tensor x, int groups, double* __restrict__ stats) {
  pdl_sync();
  extern __shared__ float sm[];  // [2][c]
  const int c = x.c, n = blockIdx.y, oct = c >> 3;
  float* ssum = sm;
  float* ssq = sm + c;
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int64_t hw = (int64_t)x.h * x.w;
  const int64_t per = (hw + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * per, p1 = min(hw, p0 + per);
  const T* base = reinterpret_cast<const T*>(x.ptr) + (int64_t)n * hw * x.ld;
  // thread -> fixed octet, strided pixels (blockDim is a multiple of oct when oct | blockDim; otherwise generic loop)
  const int64_t items = (p1 - p0) * oct;
  if (blockDim.x % oct == 0) {
    const int o = threadIdx.x % oct, lane = threadIdx.x / oct, step = blockDim.x / oct;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { s[i] = 0.f; q[i] = 0.f; }
    for (int64_t p = p0 + lane; p < p1; p += step) {
      float v[8];
      load8(base + p * x.ld + o * 8, v);
#pragma unroll
      for (int i = 0; i < 8; i++) { s[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
    }
    if (oct <= 16 && (32 % oct) == 0) {  // lanes with equal lane % oct own the same octet: combine with shuffles first
      for (int d = oct; d < 32; d <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          s[i] += __shfl_xor_sync(0xffffffffu, s[i], d);
          q[i] += __shfl_xor_sync(0xffffffffu, q[i], d);
        }
      }
      if ((threadIdx.x & 31) < oct) {
#pragma unroll
        for (int i = 0; i < 8; i++) { atomicAdd(&ssum[o * 8 + i], s[i]); atomicAdd(&ssq[o * 8 + i], q[i]); }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) { atomicAdd(&ssum[o * 8 + i], s[i]); atomicAdd(&ssq[o * 8 + i], q[i]); }
    }
  } else {
    for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
      int64_t p = p0 + it / oct;
      int o = (int)(it % oct);
      float v[8];
      load8(base + p * x.ld + o * 8, v);
#pragma unroll
      for (int i = 0; i < 8; i++) { atomicAdd(&ssum[o * 8 + i], v[i]); atomicAdd(&ssq[o * 8 + i], v[i] * v[i]); }
    }
  }
  __syncthreads();
  const int cpg = c / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < cpg; i++) { a += (double)ssum[g * cpg + i]; b += (double)ssq[g * cpg + i]; }
    atomicAdd(&stats[((int64_t)n * groups + g) * 2 + 0], a);
    atomicAdd(&stats[((int64_t)n * groups + g) * 2 + 1], b);
  }
}

template <typename T>
__global__ void __launch_bounds__(TPB, 4) gn_apply_kernel(yad_tensor x, const double* __restrict__ stats, int groups, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int act, const T* __restrict__ add, int add_ld, yad_tensor y) {
  pdl_sync();
  extern __shared__ float sm[];  // scale[c], shift[c]
  const int c = x.c, n = blockIdx.y, oct = c >> 3, cpg = c / groups;
  const int64_t hw = (int64_t)x.h * x.w;
  const double cnt = (double)hw * cpg;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    int g = ch / cpg;
    double mean = stats[((int64_t)n * groups + g) * 2] / cnt;
    double var = stats[((int64_t)n * groups + g) * 2 + 1] / cnt - mean * mean;
    float rstd = (float)(1.0 / sqrt(fmax(var, 0.0) + (double)eps));
    float sc = rstd * gamma[ch];
    sm[ch] = sc;
    sm[c + ch] = beta[ch] - (float)mean * sc;
  }
  __syncthreads();
  const int64_t items = hw * oct;
  const T* xb = reinterpret_cast<const T*>(x.ptr) + (int64_t)n * hw * x.ld;
  T* yb = reinterpret_cast<T*>(y.ptr) + (int64_t)n * hw * y.ld;
  const T* ab = add ? add + (int64_t)n * hw * add_ld : nullptr;
  if (blockDim.x % oct == 0 && items < ((int64_t)1 << 31)) {
    // the grid stride is a multiple of the octets per pixel: a thread keeps its channel octet, its scale / shift live in registers and the pixel
    // index advances by a constant -- no division, no shared-memory reads, four independent pixels in flight per trip
    const int o = (threadIdx.x % oct) * 8, pstep = (int)(gridDim.x * blockDim.x) / oct, npix = (int)hw;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { sc[i] = sm[o + i]; sh[i] = sm[c + o + i]; }
    int p = (int)(blockIdx.x * blockDim.x + threadIdx.x) / oct;
    for (; p + 3 * pstep < npix; p += 4 * pstep) {  // four independent pixels in flight, held packed until consumed
      Raw8<T> r[4], ra[4];
#pragma unroll
      for (int u = 0; u < 4; u++) raw_load(xb + (int64_t)(p + u * pstep) * x.ld + o, r[u]);
      if (ab) {
#pragma unroll
        for (int u = 0; u < 4; u++) raw_load(ab + (int64_t)(p + u * pstep) * add_ld + o, ra[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        float v[8];
        raw_unpack(r[u], v);
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fmaf(v[i], sc[i], sh[i]);
        apply_act_store_n<T>(v, act);
        if (ab) {
          float a8[8];
          raw_unpack(ra[u], a8);
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] += a8[i];
        }
        store8(yb + (int64_t)(p + u * pstep) * y.ld + o, v);
      }
    }
    for (; p < npix; p += pstep) {
      float v0[8];
      load8(xb + (int64_t)p * x.ld + o, v0);
#pragma unroll
      for (int i = 0; i < 8; i++) v0[i] = fmaf(v0[i], sc[i], sh[i]);
      apply_act_store_n<T>(v0, act);
      if (ab) {
        float a0[8];
        load8(ab + (int64_t)p * add_ld + o, a0);
#pragma unroll
        for (int i = 0; i < 8; i++) v0[i] += a0[i];
      }
      store8(yb + (int64_t)p * y.ld + o, v0);
    }
    return;
  }
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = it / oct;
    int o = (int)(it - p * oct) * 8;
    float v[8];
    load8(xb + p * x.ld + o, v);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = fmaf(v[i], sm[o + i], sm[c + o + i]);
    apply_act_n<8>(v, act);
    if (ab) {
      float a[8];
      load8(ab + p * add_ld + o, a);
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] += a[i];
    }
    store8(yb + p * y.ld + o, v);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// depthwise k x k, stride 1, pad k/2 (+bias, per-channel affine, activation, GELU gate, residual)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void dwconv_kernel(yad_tensor x, const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ scale,
                              const float* __restrict__ shift, int k, int act, int gate_split, const T* __restrict__ add, int add_ld,
                              yad_tensor y) {
  pdl_sync();
  const int cw = x.c;                              // weight channel count
  const int cout = gate_split > 0 ? gate_split : x.c;
  const int oct = cout >> 3, r = k >> 1;
  const int64_t total = (int64_t)x.n * x.h * x.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    int o = (int)(it % oct) * 8;
    int64_t p = it / oct;
    int px = (int)(p % x.w);
    int py = (int)((p / x.w) % x.h);
    int n = (int)(p / ((int64_t)x.w * x.h));
    float acc[8], acc2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = bias ? bias[o + i] : 0.f; acc2[i] = (bias && gate_split > 0) ? bias[o + gate_split + i] : 0.f; }
    for (int ky = 0; ky < k; ky++) {
      int iy = py + ky - r;
      if (iy < 0 || iy >= x.h) continue;
      for (int kx = 0; kx < k; kx++) {
        int ix = px + kx - r;
        if (ix < 0 || ix >= x.w) continue;
        const T* src = reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, iy, ix);
        const float* wt = w + (int64_t)(ky * k + kx) * cw;
        float v[8], wv[8];
        load8(src + o, v);
        load8(wt + o, wv);
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fmaf(v[i], wv[i], acc[i]);
        if (gate_split > 0) {
          load8(src + o + gate_split, v);
          load8(wt + o + gate_split, wv);
#pragma unroll
          for (int i = 0; i < 8; i++) acc2[i] = fmaf(v[i], wv[i], acc2[i]);
        }
      }
    }
    if (gate_split > 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = apply_act(acc[i], YAD_ACT_GELU) * acc2[i];
    } else {
      if (scale) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fmaf(acc[i], scale[o + i], shift[o + i]);
      }
      apply_act_n<8>(acc, act);
    }
    if (add) {
      float a[8];
      load8(add + p * add_ld + o, a);
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] += a[i];
    }
    store8(reinterpret_cast<T*>(y.ptr) + p * y.ld + o, acc);
  }
}

// Row-tiled depthwise K x K (the ungated case): a thread owns 8 channels of PX adjacent output pixels of one row.  Per kernel row it keeps the K
// weight vectors in registers and streams the K + PX - 1 input columns once, so a 7x7 output costs (7 weight + 10 input) / 4 vector accesses per
// kernel row instead of 7 + 7 (the L1 wavefront count, which bounds dwconv_kernel at k = 7, drops 3.5x).  Channel octets fastest -> every warp
// load is made of full 128-byte segments.
template <typename T, int K, int PX>
__global__ void __launch_bounds__(128) dwconv_row_kernel(yad_tensor x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                                         const T* __restrict__ add, int add_ld, yad_tensor y, int groups_x, int total) {
  pdl_sync();
  constexpr int R = K / 2;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= total) return;
  const int c = x.c, oct = c >> 3;
  const int o = (item % oct) * 8;
  int g = item / oct;
  const int px0 = (g % groups_x) * PX;
  g /= groups_x;
  const int py = g % x.h, n = g / x.h;
  const T* xb = reinterpret_cast<const T*>(x.ptr) + (int64_t)n * x.h * x.w * x.ld + o;
  float acc[PX][8];
  {
    float bv[8];
#pragma unroll
    for (int i = 0; i < 8; i++) bv[i] = 0.f;
    if (bias) load8(bias + o, bv);
#pragma unroll
    for (int p = 0; p < PX; p++)
#pragma unroll
      for (int i = 0; i < 8; i++) acc[p][i] = bv[i];
  }
#pragma unroll
  for (int ky = 0; ky < K; ky++) {
    const int iy = py + ky - R;
    if (iy < 0 || iy >= x.h) continue;
    float wr[K][8];
#pragma unroll
    for (int dx = 0; dx < K; dx++) load8(w + (ky * K + dx) * c + o, wr[dx]);
    const T* xr = xb + (int64_t)iy * x.w * x.ld;
#pragma unroll
    for (int j = 0; j < K + PX - 1; j++) {
      const int ix = px0 + j - R;
      if (ix < 0 || ix >= x.w) continue;
      float xv[8];
      load8(xr + (int64_t)ix * x.ld, xv);
#pragma unroll
      for (int p = 0; p < PX; p++) {
        const int dx = j - p;  // static after unrolling
        if (dx < 0 || dx >= K) continue;
#pragma unroll
        for (int i = 0; i < 8; i++) acc[p][i] = fmaf(xv[i], wr[dx][i], acc[p][i]);
      }
    }
  }
  float sc[8], sh[8];
  if (scale) { load8(scale + o, sc); load8(shift + o, sh); }
#pragma unroll
  for (int p = 0; p < PX; p++) {
    if (px0 + p >= x.w) break;
    if (scale) {
#pragma unroll
      for (int i = 0; i < 8; i++) acc[p][i] = fmaf(acc[p][i], sc[i], sh[i]);
    }
    apply_act_n<8>(acc[p], act);
    const int64_t d = ((int64_t)n * x.h + py) * x.w + px0 + p;
    if (add) {
      float a[8];
      load8(add + d * add_ld + o, a);
#pragma unroll
      for (int i = 0; i < 8; i++) acc[p][i] += a[i];
    }
    store8(reinterpret_cast<T*>(y.ptr) + d * y.ld + o, acc[p]);
  }
}

// Shared-memory tiled depthwise K x K (ungated): a CTA owns a TH x TW output tile of one image and CC channels (64 bytes of T).  Phase 1 copies the
// haloed (TH + K - 1) x (TW + K - 1) x CC input tile (zeros outside the image) and the K*K x CC weights to shared memory with 128-bit loads that
// are all in flight together; phase 2 is dwconv_row_kernel's inner loop on shared memory (a thread = 8 channels x PX adjacent outputs, the K
// weights of a kernel row in registers).  The K-fold re-read of every input row moves from L2 (whose latency bound dwconv_row_kernel, above all
// on the 20x20 maps where a whole image is one tile) to shared memory.  Pixel pitch = 64 + 16 bytes: the two pixel groups of a quarter warp
// (PX pixels apart) fall on disjoint banks.
constexpr int DWS_PX = 4;
constexpr int DWS_THREADS = 256;
template <typename T, int K>
__global__ void __launch_bounds__(DWS_THREADS, 2) dwconv_smem_kernel(yad_tensor x, const float* __restrict__ w, const float* __restrict__ bias,
                                                                  const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                                                  const T* __restrict__ add, int add_ld, yad_tensor y, int th, int tw,
                                                                  int tiles_x, int tiles_y) {
  constexpr int R = K / 2, CC = 64 / (int)sizeof(T), OCT = CC / 8, PITCH = 80, PIX_T = PITCH / (int)sizeof(T);
  extern __shared__ __align__(16) uint8_t dws_sm[];
  const int groups_x = (tw + DWS_PX - 1) / DWS_PX;
  const int iw = groups_x * DWS_PX + K - 1, ih = th + K - 1;  // whole pixel groups: the inner loop never reads past the staged tile
  float* wsm = reinterpret_cast<float*>(dws_sm);                       // [K*K][CC]
  uint8_t* xsm = dws_sm + K * K * CC * sizeof(float);                  // [ih][iw] pixels of PITCH bytes
  int b = blockIdx.x;
  const int c0 = (b % (x.c / CC)) * CC;
  b /= x.c / CC;
  const int tx0 = (b % tiles_x) * tw;
  b /= tiles_x;
  const int ty0 = (b % tiles_y) * th, n = b / tiles_y;
  for (int i = threadIdx.x; i < K * K * (CC / 4); i += blockDim.x) {
    const int t = i / (CC / 4), q = i - t * (CC / 4);
    reinterpret_cast<float4*>(wsm)[i] = *reinterpret_cast<const float4*>(w + t * x.c + c0 + q * 4);
  }
  pdl_sync();
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(reinterpret_cast<const T*>(x.ptr) + (int64_t)n * x.h * x.w * x.ld + c0);
  const int64_t pix_bytes = (int64_t)x.ld * sizeof(T);
#pragma unroll 4
  for (int v = threadIdx.x; v < ih * iw * 4; v += blockDim.x) {
    const int pix = v >> 2, part = v & 3;
    const int r = pix / iw, cx = pix - r * iw;
    const int iy = ty0 + r - R, ix = tx0 + cx - R;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < x.h && ix >= 0 && ix < x.w) val = *reinterpret_cast<const uint4*>(xb + ((int64_t)iy * x.w + ix) * pix_bytes + part * 16);
    *reinterpret_cast<uint4*>(xsm + pix * PITCH + part * 16) = val;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < th * groups_x * OCT; item += blockDim.x) {
    const int oi = item % OCT;
    const int g = item / OCT;
    const int ry = g / groups_x, rx0 = (g - ry * groups_x) * DWS_PX;
    const int o = c0 + oi * 8;
    const int oy = ty0 + ry;
    if (oy >= x.h || tx0 + rx0 >= x.w) continue;
    float acc[DWS_PX][8];
    {
      float bv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) bv[i] = 0.f;
      if (bias) load8(bias + o, bv);
#pragma unroll
      for (int p = 0; p < DWS_PX; p++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc[p][i] = bv[i];
    }
#pragma unroll 1  // one kernel row per trip keeps the loop body (K + PX - 1 loads, K * PX * 8 FMAs) inside the instruction cache
    for (int ky = 0; ky < K; ky++) {
      float wr[K][8];
#pragma unroll
      for (int dx = 0; dx < K; dx++) load8(wsm + (ky * K + dx) * CC + oi * 8, wr[dx]);
      const T* xr = reinterpret_cast<const T*>(xsm) + ((ry + ky) * iw + rx0) * PIX_T + oi * 8;
#pragma unroll
      for (int j = 0; j < K + DWS_PX - 1; j++) {
        float xv[8];
        load8(xr + j * PIX_T, xv);
#pragma unroll
        for (int p = 0; p < DWS_PX; p++) {
          const int dx = j - p;  // static after unrolling
          if (dx < 0 || dx >= K) continue;
#pragma unroll
          for (int i = 0; i < 8; i++) acc[p][i] = fmaf(xv[i], wr[dx][i], acc[p][i]);
        }
      }
    }
    float sc[8], sh[8];
    if (scale) { load8(scale + o, sc); load8(shift + o, sh); }
#pragma unroll
    for (int p = 0; p < DWS_PX; p++) {
      const int ox = tx0 + rx0 + p;
      if (rx0 + p >= tw || ox >= x.w) break;
      if (scale) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[p][i] = fmaf(acc[p][i], sc[i], sh[i]);
      }
      apply_act_n<8>(acc[p], act);
      const int64_t d = ((int64_t)n * x.h + oy) * x.w + ox;
      if (add) {
        float a[8];
        load8(add + d * add_ld + o, a);
#pragma unroll
        for (int i = 0; i < 8; i++) acc[p][i] += a[i];
      }
      store8(reinterpret_cast<T*>(y.ptr) + d * y.ld + o, acc[p]);
    }
  }
}

// tile edge for a map edge: whole edge up to 24, else the edge in [12, 24] that wastes the fewest rows (ties -> larger)
inline int dws_tile(int n) {
  if (n <= 24) return n;
  int best = 24, waste = 1 << 30;
  for (int t = 24; t >= 12; t--) {
    const int wst = (n + t - 1) / t * t - n;
    if (wst < waste) { waste = wst; best = t; }
  }
  return best;
}

// ------------------------------------------------------------------------------------------------------------------
// SPPF: 5x5 / 9x9 / 13x13 max windows (= three chained 5x5 s1 p2 max-pools with -inf padding)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void sppf_pool_kernel(yad_tensor x, yad_tensor y1, yad_tensor y2, yad_tensor y3) {
  pdl_sync();
  const int oct = x.c >> 3;
  const int64_t total = (int64_t)x.n * x.h * x.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    int o = (int)(it % oct) * 8;
    int64_t p = it / oct;
    int px = (int)(p % x.w);
    int py = (int)((p / x.w) % x.h);
    int n = (int)(p / ((int64_t)x.w * x.h));
    float m1[8], m2[8], m3[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m1[i] = m2[i] = m3[i] = -INFINITY;
    for (int dy = -6; dy <= 6; dy++) {
      int iy = py + dy;
      if (iy < 0 || iy >= x.h) continue;
      for (int dx = -6; dx <= 6; dx++) {
        int ix = px + dx;
        if (ix < 0 || ix >= x.w) continue;
        float v[8];
        load8(reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, iy, ix) + o, v);
        int rad = max(abs(dy), abs(dx));
#pragma unroll
        for (int i = 0; i < 8; i++) {
          m3[i] = fmaxf(m3[i], v[i]);
          if (rad <= 4) m2[i] = fmaxf(m2[i], v[i]);
          if (rad <= 2) m1[i] = fmaxf(m1[i], v[i]);
        }
      }
    }
    store8(reinterpret_cast<T*>(y1.ptr) + p * y1.ld + o, m1);
    store8(reinterpret_cast<T*>(y2.ptr) + p * y2.ld + o, m2);
    store8(reinterpret_cast<T*>(y3.ptr) + p * y3.ld + o, m3);
  }
}

// Tiled SPPF for bf16 storage: the three windows as what they are in the reference -- three chained 5x5 pools -- each one a horizontal and a
// vertical 5-tap pass over a haloed tile in shared memory (packed bf16x2 max).  A CTA owns a T x T output tile (T <= 20) of one image and 16
// channels; the (T + 12)^2 haloed tile is loaded once (-inf outside the image = the pools' padding); after stage k the outer 2k rows / columns
// of the tile are stale, the centre never is (a map that is one tile needs no halo: the tile edge is the pools' own -inf padding).  30 shared-memory reads per output instead of the 275 global loads of sppf_pool_kernel.
constexpr int SP_CC = 16, SP_HALO = 6, SP_TMAX = 20;
__device__ __forceinline__ uint4 bf8_max(const uint4& a, const uint4& b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}
__global__ void __launch_bounds__(256) sppf_tile_kernel(yad_tensor x, yad_tensor y1, yad_tensor y2, yad_tensor y3, int T, int tiles_x, int tiles_y, int halo) {
  extern __shared__ __align__(16) uint4 sp_sm[];  // A[TD * TD][2], B[TD * TD][2]
  const int TD = T + 2 * halo, cells = TD * TD * 2;
  uint4* A = sp_sm;
  uint4* B = sp_sm + cells;
  int b = blockIdx.x;
  const int chunks = x.c / SP_CC, c0 = (b % chunks) * SP_CC;
  b /= chunks;
  const int tx0 = (b % tiles_x) * T;
  b /= tiles_x;
  const int ty0 = (b % tiles_y) * T, n = b / tiles_y;
  pdl_sync();
  const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
  for (int idx = threadIdx.x; idx < cells; idx += 256) {
    const int pix = idx >> 1, part = idx & 1, ty = pix / TD, tx = pix - ty * TD;
    const int iy = ty0 - halo + ty, ix = tx0 - halo + tx;
    uint4 v = ninf;
    if (iy >= 0 && iy < x.h && ix >= 0 && ix < x.w) v = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(x.ptr) + pix_off(x, n, iy, ix) + c0 + part * 8);
    A[idx] = v;
  }
  __syncthreads();
  for (int stage = 0; stage < 3; stage++) {
    for (int idx = threadIdx.x; idx < cells; idx += 256) {  // horizontal 5-tap max: A -> B
      const int pix = idx >> 1, ty = pix / TD, tx = pix - ty * TD;
      uint4 v = A[idx];
#pragma unroll
      for (int d = -2; d <= 2; d++)
        if (d != 0 && tx + d >= 0 && tx + d < TD) v = bf8_max(v, A[idx + 2 * d]);
      B[idx] = v;
    }
    __syncthreads();
    bf16* const yp = reinterpret_cast<bf16*>(stage == 0 ? y1.ptr : (stage == 1 ? y2.ptr : y3.ptr));
    const int yld = stage == 0 ? y1.ld : (stage == 1 ? y2.ld : y3.ld);
    for (int idx = threadIdx.x; idx < cells; idx += 256) {  // vertical 5-tap max: B -> A, the centre leaves for y_stage
      const int pix = idx >> 1, part = idx & 1, ty = pix / TD, tx = pix - ty * TD;
      uint4 v = B[idx];
#pragma unroll
      for (int d = -2; d <= 2; d++)
        if (d != 0 && ty + d >= 0 && ty + d < TD) v = bf8_max(v, B[idx + 2 * d * TD]);
      A[idx] = v;
      const int oy = ty0 - halo + ty, ox = tx0 - halo + tx;
      if (ty >= halo && ty < halo + T && tx >= halo && tx < halo + T && oy < x.h && ox < x.w)
        *reinterpret_cast<uint4*>(yp + ((int64_t)(n * x.h + oy) * x.w + ox) * yld + c0 + part * 8) = v;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// reductions over pixels: global average pool, row / column means, MLCA 5x5 adaptive pool
// ------------------------------------------------------------------------------------------------------------------
// Sum of 8-channel vectors over a pixel range; block = (oct lanes) x (pixel lanes).  Result for octet o, channel i lands in
// smem acc[o*8+i] (caller zeroes it).
template <typename T, typename F>
__device__ __forceinline__ void block_pixel_sum(const yad_tensor& x, int64_t count, F pixel_offset, float* acc) {
  const int oct = x.c >> 3;
  if (blockDim.x % oct == 0) {  // thread -> fixed octet, strided pixels: register accumulation, 8 smem atomics per thread in total
    const int o = (threadIdx.x % oct) * 8, lane = threadIdx.x / oct, step = blockDim.x / oct;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = 0.f;
    int64_t j = lane;
    for (; j + 3 * step < count; j += 4 * step) {  // four packed loads in flight (one per trip was latency-bound: 1.6 TB/s on a 52 MB map)
      Raw8<T> r[4];
#pragma unroll
      for (int u = 0; u < 4; u++) raw_load(reinterpret_cast<const T*>(x.ptr) + pixel_offset(j + u * step) + o, r[u]);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        float v[8];
        raw_unpack(r[u], v);
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] += v[i];
      }
    }
    for (; j < count; j += step) {
      float v[8];
      load8(reinterpret_cast<const T*>(x.ptr) + pixel_offset(j) + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] += v[i];
    }
    if (lane < count) {
#pragma unroll
      for (int i = 0; i < 8; i++) atomicAdd(&acc[o + i], s[i]);
    }
    return;
  }
  const int64_t items = count * oct;
  for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
    int64_t j = it / oct;
    int o = (int)(it - j * oct) * 8;
    float v[8];
    load8(reinterpret_cast<const T*>(x.ptr) + pixel_offset(j) + o, v);
#pragma unroll
    for (int i = 0; i < 8; i++) atomicAdd(&acc[o + i], v[i]);
  }
}

// grid (splits, n): out[n][c] += partial mean (out zeroed by the launcher)
template <typename T>
__global__ void gap_kernel(yad_tensor x, float* __restrict__ out) {
  pdl_sync();
  extern __shared__ float sm[];
  const int n = blockIdx.y, c = x.c, oct = c >> 3;
  for (int i = threadIdx.x; i < c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int64_t hw = (int64_t)x.h * x.w;
  const int64_t per = (hw + gridDim.x - 1) / gridDim.x, p0 = blockIdx.x * per, p1 = min(hw, p0 + per);
  const T* base = reinterpret_cast<const T*>(x.ptr) + (int64_t)n * hw * x.ld;
  if (blockDim.x % oct == 0) {
    const int o = (threadIdx.x % oct) * 8, lane = threadIdx.x / oct, step = blockDim.x / oct;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = 0.f;
    int64_t p = p0 + lane;
    for (; p + 3 * step < p1; p += 4 * step) {  // four packed loads in flight
      Raw8<T> r[4];
#pragma unroll
      for (int u = 0; u < 4; u++) raw_load(base + (p + u * step) * x.ld + o, r[u]);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        float v[8];
        raw_unpack(r[u], v);
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] += v[i];
      }
    }
    for (; p < p1; p += step) {
      float v[8];
      load8(base + p * x.ld + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) atomicAdd(&sm[o + i], s[i]);
  } else if (p1 > p0) {
    block_pixel_sum<T>(x, p1 - p0, [&](int64_t j) { return ((int64_t)n * hw + p0 + j) * x.ld; }, sm);
  }
  __syncthreads();
  const float inv = 1.0f / (float)hw;
  for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(&out[(int64_t)n * c + i], sm[i] * inv);
}

// LayerNorm over the channels of every pixel mixed with the input: y = LN(x) * gamma + x * gammax  (Mona.forward, nn/modules/mona.py:55 with
// LayerNorm2d :5-10).  g = lanes per pixel (power of two >= c / 8, at most 32): a warp normalises 32 / g pixels, each lane holds up to four
// 8-channel vectors in registers, mean and the centred second moment are reduced with xor shuffles inside the lane group.
template <typename T>
__global__ void __launch_bounds__(256) ln_mix_kernel(yad_tensor x, const float* __restrict__ w, const float* __restrict__ b,
                                                     const float* __restrict__ gamma, const float* __restrict__ gammax, float eps, yad_tensor y, int g) {
  pdl_sync();
  const int lane = threadIdx.x & 31, sub = lane & (g - 1), oct = x.c >> 3;
  const int64_t npix = (int64_t)x.n * x.h * x.w;
  const int64_t pix = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / g) + lane / g;
  const bool valid = pix < npix;
  const T* xp = reinterpret_cast<const T*>(x.ptr) + (valid ? pix : 0) * x.ld;
  float v[4][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int o = sub + k * g;
    if (valid && o < oct) {
      load8(xp + o * 8, v[k]);
#pragma unroll
      for (int i = 0; i < 8; i++) s += v[k][i];
    }
  }
  for (int d = g >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float mean = s / (float)x.c;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (valid && sub + k * g < oct) {
#pragma unroll
      for (int i = 0; i < 8; i++) { const float dlt = v[k][i] - mean; q = fmaf(dlt, dlt, q); }
    }
  }
  for (int d = g >> 1; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
  const float rstd = 1.0f / sqrtf(q / (float)x.c + eps);
  if (!valid) return;
  T* yp = reinterpret_cast<T*>(y.ptr) + pix * y.ld;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int o = sub + k * g;
    if (o < oct) {
      float wv[8], bv[8], gv[8], xv[8], r[8];
      load8(w + o * 8, wv); load8(b + o * 8, bv); load8(gamma + o * 8, gv); load8(gammax + o * 8, xv);
#pragma unroll
      for (int i = 0; i < 8; i++) r[i] = fmaf(fmaf((v[k][i] - mean) * rstd, wv[i], bv[i]), gv[i], v[k][i] * xv[i]);
      store8(yp + o * 8, r);
    }
  }
}

// grid (L, n, 2): z=0 -> row y=blockIdx.x: mean over x ; z=1 -> column x=blockIdx.x: mean over y
template <typename T>
__global__ void rowcol_mean_kernel(yad_tensor x, yad_tensor rowmean, yad_tensor colmean) {
  pdl_sync();
  extern __shared__ float sm[];
  const int n = blockIdx.y, c = x.c;
  const bool is_col = blockIdx.z == 1;
  const int L = is_col ? x.w : x.h;
  if ((int)blockIdx.x >= L) return;
  for (int i = threadIdx.x; i < c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int idx = blockIdx.x;
  const int count = is_col ? x.h : x.w;
  block_pixel_sum<T>(x, count, [&](int64_t j) { return is_col ? pix_off(x, n, (int)j, idx) : pix_off(x, n, idx, (int)j); }, sm);
  __syncthreads();
  const yad_tensor& o = is_col ? colmean : rowmean;  // views (n, L, 1, c)
  T* dst = reinterpret_cast<T*>(o.ptr) + ((int64_t)n * L + idx) * o.ld;
  const float inv = 1.0f / (float)count;
  for (int i = threadIdx.x; i < c; i += blockDim.x) st1(dst + i, sm[i] * inv);
}

// Single-pass variant (the map is read ONCE; rowcol_mean_kernel reads it once for the rows and once for the columns): one CTA per (image,
// slice of 16 channels), 8 warps = 2 channel octets x 4 row groups (rows y = group, group + 4, ...).  The 32 lanes of a warp sweep a whole row
// (x = lane + 32 k, NK positions), so a row costs NK 128-bit loads and ONE butterfly reduce-scatter (9 shuffles for 8 channels) whose result goes
// straight to the row table in shared memory; the column sums stay in registers (NK x 8 per thread) and the four row groups are merged through
// shared memory at the end.  History (profiles/r2_pointwise.md): lanes = 32 pixels of a quarter row needed one reduction per 32 pixels and was
// instruction-bound (33.6 M warp instructions, 60 us for the 105 MB map = the two-pass kernel's time); 32-channel slices with four rows in flight
// were latency-bound at the same 60 us.  Requires c % 16 == 0 and w <= 160 (the caller falls back otherwise).
template <typename T, int NK>
__global__ void __launch_bounds__(256, 2) rowcol_mean1_kernel(yad_tensor x, yad_tensor rowmean, yad_tensor colmean) {
  pdl_sync();
  extern __shared__ float sm[];  // rows[h][16], then cols[3][w][16] (row groups 1..3)
  float* smc = sm + x.h * 16;
  const int slices = x.c >> 4;
  const int n = blockIdx.x / slices, c0 = (blockIdx.x - n * slices) * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, oct = warp & 1, rg = warp >> 1;
  const T* xb = reinterpret_cast<const T*>(x.ptr) + (int64_t)n * x.h * x.w * x.ld + c0 + oct * 8;
  float col[NK][8];
#pragma unroll
  for (int k = 0; k < NK; k++)
#pragma unroll
    for (int i = 0; i < 8; i++) col[k][i] = 0.f;
  constexpr int RU = (sizeof(T) == 2 ? 8 : 4) / NK > 0 ? (sizeof(T) == 2 ? 8 : 4) / NK : 1;  // rows in flight per warp
  for (int y0 = rg; y0 < x.h; y0 += 4 * RU) {
    Raw8<T> r[RU][NK];
#pragma unroll
    for (int u = 0; u < RU; u++)
#pragma unroll
      for (int k = 0; k < NK; k++)
        if (y0 + 4 * u < x.h && lane + 32 * k < x.w) raw_load(xb + ((int64_t)(y0 + 4 * u) * x.w + lane + 32 * k) * x.ld, r[u][k]);
#pragma unroll
    for (int u = 0; u < RU; u++) {
      const int yy = y0 + 4 * u;
      if (yy >= x.h) break;  // warp-uniform
      float row[8];
#pragma unroll
      for (int i = 0; i < 8; i++) row[i] = 0.f;
#pragma unroll
      for (int k = 0; k < NK; k++)
        if (lane + 32 * k < x.w) {
          float v[8];
          raw_unpack(r[u][k], v);
#pragma unroll
          for (int i = 0; i < 8; i++) { col[k][i] += v[i]; row[i] += v[i]; }
        }
      // reduce-scatter over the 32 lanes: lane bits 16 / 8 / 4 each keep one half of the channels, bits 2 / 1 finish the sum
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const bool up = (lane & 16) != 0;
        const float got = __shfl_xor_sync(0xffffffffu, up ? row[i] : row[i + 4], 16);
        row[i] = (up ? row[i + 4] : row[i]) + got;
      }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const bool up = (lane & 8) != 0;
        const float got = __shfl_xor_sync(0xffffffffu, up ? row[i] : row[i + 2], 8);
        row[i] = (up ? row[i + 2] : row[i]) + got;
      }
      {
        const bool up = (lane & 4) != 0;
        const float got = __shfl_xor_sync(0xffffffffu, up ? row[0] : row[1], 4);
        row[0] = (up ? row[1] : row[0]) + got;
      }
      row[0] += __shfl_xor_sync(0xffffffffu, row[0], 2);
      row[0] += __shfl_xor_sync(0xffffffffu, row[0], 1);
      if ((lane & 3) == 0) sm[yy * 16 + oct * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = row[0];
    }
  }
  // columns: row groups 1..3 hand their partial sums to group 0
  if (rg > 0) {
#pragma unroll
    for (int k = 0; k < NK; k++)
      if (lane + 32 * k < x.w) {
        float* d = smc + ((size_t)(rg - 1) * x.w + lane + 32 * k) * 16 + oct * 8;
        *reinterpret_cast<float4*>(d) = make_float4(col[k][0], col[k][1], col[k][2], col[k][3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(col[k][4], col[k][5], col[k][6], col[k][7]);
      }
  }
  __syncthreads();
  if (rg == 0) {
    T* cb = reinterpret_cast<T*>(colmean.ptr) + (int64_t)n * x.w * colmean.ld + c0 + oct * 8;
    const float inv = 1.0f / (float)x.h;
#pragma unroll
    for (int k = 0; k < NK; k++) {
      const int xx = lane + 32 * k;
      if (xx < x.w) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = col[k][i];
        for (int g = 0; g < 3; g++) {
          const float* d = smc + ((size_t)g * x.w + xx) * 16 + oct * 8;
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] += d[i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] *= inv;
        store8(cb + (int64_t)xx * colmean.ld, v);
      }
    }
  }
  {
    T* rb = reinterpret_cast<T*>(rowmean.ptr) + (int64_t)n * x.h * rowmean.ld + c0;
    const float inv = 1.0f / (float)x.w;
    for (int i = threadIdx.x; i < x.h * 2; i += blockDim.x) {
      const int y = i >> 1, o = (i & 1) * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = sm[y * 16 + o + j] * inv;
      store8(rb + (int64_t)y * rowmean.ld + o, v);
    }
  }
}

// CoordAtt (nn/modules/head.py:694-703) between its pooling and its gating: y = hardswish(bn1(conv1(cat[x_h, x_w]))) (conv1 + eval-mode bn1 folded
// into w1 [MIP][c] / b1 [MIP]), then a_h = sigmoid(conv_h(y_h)), a_w = sigmoid(conv_w(y_w)) (wh, ww [co][MIP]).  One warp per position of the
// concatenated axis: lanes stride over the input channels, the MIP hidden sums are reduced with shuffles, then lanes stride over the output channels.
// Replaces four 1x1 yad_conv2d launches on (n, L, 1, c) views (each at the launch floor); fp32 weights and hidden activations.
template <typename T, int MIP>
__global__ void __launch_bounds__(TPB) coordatt_mlp_kernel(yad_tensor rows, yad_tensor cols, const float* __restrict__ w1, const float* __restrict__ b1,
                                                           const float* __restrict__ wh, const float* __restrict__ bh, const float* __restrict__ ww,
                                                           const float* __restrict__ bw, yad_tensor gh, yad_tensor gw) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t pos = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nrows = (int64_t)rows.n * rows.h, ncols = (int64_t)cols.n * cols.h;
  if (pos >= nrows + ncols) return;
  const bool is_row = pos < nrows;
  const int64_t idx = is_row ? pos : pos - nrows;
  const yad_tensor& src = is_row ? rows : cols;
  const yad_tensor& dst = is_row ? gh : gw;
  const T* xp = reinterpret_cast<const T*>(src.ptr) + idx * src.ld;
  float hid[MIP];
#pragma unroll
  for (int j = 0; j < MIP; j++) hid[j] = 0.f;
  const int c = src.c;
  for (int ch = lane; ch < c; ch += 32) {
    const float v = ld1(xp + ch);
#pragma unroll
    for (int j = 0; j < MIP; j++) hid[j] = fmaf(__ldg(w1 + j * c + ch), v, hid[j]);
  }
#pragma unroll
  for (int j = 0; j < MIP; j++) {
    float t = hid[j];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    t += __ldg(b1 + j);
    hid[j] = t * fminf(fmaxf(t + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);  // nn.Hardswish
  }
  const float* wo = is_row ? wh : ww;
  const float* bo = is_row ? bh : bw;
  T* yp = reinterpret_cast<T*>(dst.ptr) + idx * dst.ld;
  for (int co = lane; co < dst.c; co += 32) {
    float t = __ldg(bo + co);
#pragma unroll
    for (int j = 0; j < MIP; j++) t = fmaf(__ldg(wo + co * MIP + j), hid[j], t);
    st1(yp + co, sigmoidf_(t));
  }
}

template <typename T>
__global__ void rowcol_gate_kernel(yad_tensor x, bool has_x, yad_tensor gh, yad_tensor gw, yad_tensor y) {
  pdl_sync();
  const int oct = y.c >> 3;
  const int64_t total = (int64_t)y.n * y.h * y.w * oct;
  if (has_x && blockDim.x % oct == 0 && total < ((int64_t)1 << 31)) {
    // as gn_apply_kernel: a thread keeps its channel octet, 32-bit pixel arithmetic, four packed loads of x in flight per trip
    const int o = (threadIdx.x % oct) * 8, pstep = (int)(gridDim.x * blockDim.x) / oct, npix = y.n * y.h * y.w, hw = y.h * y.w;
    const T* xp = reinterpret_cast<const T*>(x.ptr) + o;
    T* yp = reinterpret_cast<T*>(y.ptr) + o;
    auto gate = [&](int p, const float (&v)[8]) {
      const int n = p / hw, r = p - n * hw, py = r / y.w, px = r - py * y.w;
      float a[8], b[8];
      load8(reinterpret_cast<const T*>(gh.ptr) + ((int64_t)n * y.h + py) * gh.ld + o, a);
      load8(reinterpret_cast<const T*>(gw.ptr) + ((int64_t)n * y.w + px) * gw.ld + o, b);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = v[i] * a[i] * b[i];
      store8(yp + (int64_t)p * y.ld, a);
    };
    int p = (int)(blockIdx.x * blockDim.x + threadIdx.x) / oct;
    for (; p + 3 * pstep < npix; p += 4 * pstep) {
      Raw8<T> r[4];
#pragma unroll
      for (int u = 0; u < 4; u++) raw_load(xp + (int64_t)(p + u * pstep) * x.ld, r[u]);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        float v[8];
        raw_unpack(r[u], v);
        gate(p + u * pstep, v);
      }
    }
    for (; p < npix; p += pstep) {
      float v[8];
      load8(xp + (int64_t)p * x.ld, v);
      gate(p, v);
    }
    return;
  }
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    int o = (int)(it % oct) * 8;
    int64_t p = it / oct;
    int px = (int)(p % y.w);
    int py = (int)((p / y.w) % y.h);
    int n = (int)(p / ((int64_t)y.w * y.h));
    float a[8], b[8];
    load8(reinterpret_cast<const T*>(gh.ptr) + ((int64_t)n * y.h + py) * gh.ld + o, a);
    load8(reinterpret_cast<const T*>(gw.ptr) + ((int64_t)n * y.w + px) * gw.ld + o, b);
    if (has_x) {
      float v[8];
      load8(reinterpret_cast<const T*>(x.ptr) + p * x.ld + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = v[i] * a[i] * b[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = a[i] * b[i];
    }
    store8(reinterpret_cast<T*>(y.ptr) + p * y.ld + o, a);
  }
}

__device__ __forceinline__ int bin_start(int i, int in, int out) { return (int)(((int64_t)i * in) / out); }
__device__ __forceinline__ int bin_end(int i, int in, int out) { return (int)((((int64_t)(i + 1)) * in + out - 1) / out); }

// grid (25, n): local[n][bin][c] = adaptive average of bin (block.py:1559 local_arv_pool)
template <typename T>
__global__ void mlca_pool_kernel(yad_tensor x, float* __restrict__ local, int ls) {
  pdl_sync();
  extern __shared__ float sm[];
  const int n = blockIdx.y, c = x.c, bin = blockIdx.x, by = bin / ls, bx = bin % ls;
  for (int i = threadIdx.x; i < c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int y0 = bin_start(by, x.h, ls), y1 = bin_end(by, x.h, ls), x0 = bin_start(bx, x.w, ls), x1 = bin_end(bx, x.w, ls);
  const int bw = x1 - x0, cnt = (y1 - y0) * bw;
  block_pixel_sum<T>(x, cnt, [&](int64_t j) { return pix_off(x, n, y0 + (int)(j / bw), x0 + (int)(j % bw)); }, sm);
  __syncthreads();
  const float inv = 1.0f / (float)cnt;
  for (int i = threadIdx.x; i < c; i += blockDim.x) local[((int64_t)n * ls * ls + bin) * c + i] = sm[i] * inv;
}

// one block per image: att[n][bin][c] = (1-lw)*G[row(bin)][c] + lw*sigmoid(conv1d_k(local sequence))  (block.py:1565-1581).
// G: the reference reshapes the global branch to a 3-D (c, b, 1) tensor before `adaptive_avg_pool2d(.., [ls, ls])` (block.py:1575-1579),
// so the pool runs over the BATCH axis: G[i][c] = mean over images b in [floor(i*B/ls), ceil((i+1)*B/ls)) of sigmoid(conv1d_k(glob_b))[c],
// shared by every image (for B == 1 this is the plain per-image broadcast).  Reproduced as is.
// pass 1 (one CTA per image): sg[b][ch] = sigmoid(conv1d_k(mean of the image's ls x ls bins))[ch]
__global__ void mlca_glob_kernel(const float* __restrict__ local, const float* __restrict__ wg, int k, int c, int ls, float* __restrict__ sg) {
  pdl_sync();
  extern __shared__ float sm[];  // glob[c]
  const int nb = ls * ls, b = blockIdx.x, len = nb * c, r = (k - 1) / 2;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < nb; q++) s += local[(int64_t)b * len + q * c + ch];
    sm[ch] = s / (float)nb;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < k; j++) {
      int q = ch + j - r;
      if (q >= 0 && q < c) s = fmaf(wg[j], sm[q], s);
    }
    sg[(int64_t)b * c + ch] = sigmoidf_(s);
  }
}

// pass 2 (one CTA per image): G from sg (batch-axis bins, in image order as the single-kernel version summed them), local branch, blend
__global__ void mlca_att_kernel(const float* __restrict__ local, const float* __restrict__ sg, const float* __restrict__ wl, int k, float lw,
                                int c, int ls, int batch, float* __restrict__ att) {
  pdl_sync();
  extern __shared__ float sm[];  // seq[nb*c], G[ls*c]
  const int nb = ls * ls;
  float* seq = sm;
  float* G = sm + nb * c;
  const int n = blockIdx.x, len = nb * c, r = (k - 1) / 2;
  for (int i = threadIdx.x; i < len; i += blockDim.x) seq[i] = local[(int64_t)n * len + i];
  for (int i = threadIdx.x; i < ls * c; i += blockDim.x) {
    const int bi = i / c, ch = i - bi * c;
    const int b0 = bin_start(bi, batch, ls), b1 = bin_end(bi, batch, ls);
    float s = 0.f;
    for (int b = b0; b < b1; b++) s += sg[(int64_t)b * c + ch] / (float)(b1 - b0);
    G[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < k; j++) {
      int q = i + j - r;
      if (q >= 0 && q < len) s = fmaf(wl[j], seq[q], s);
    }
    att[(int64_t)n * len + i] = G[((i / c) / ls) * c + i % c] * (1.0f - lw) + sigmoidf_(s) * lw;
  }
}

// y = x * adaptive_avg_pool2d(att (ls x ls) -> (h, w)) (+ add)   (block.py:1581-1583)
template <typename T>
__global__ void mlca_apply_kernel(yad_tensor x, const float* __restrict__ att, int ls, const T* __restrict__ add, int add_ld, yad_tensor y) {
  pdl_sync();
  const int oct = x.c >> 3, c = x.c;
  const int64_t total = (int64_t)x.n * x.h * x.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    int o = (int)(it % oct) * 8;
    int64_t p = it / oct;
    int px = (int)(p % x.w);
    int py = (int)((p / x.w) % x.h);
    int n = (int)(p / ((int64_t)x.w * x.h));
    const int y0 = bin_start(py, ls, x.h), y1 = bin_end(py, ls, x.h), x0 = bin_start(px, ls, x.w), x1 = bin_end(px, ls, x.w);
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = 0.f;
    for (int by = y0; by < y1; by++)
      for (int bx = x0; bx < x1; bx++) {
        float t[8];
        load8(att + ((int64_t)n * ls * ls + by * ls + bx) * c + o, t);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] += t[i];
      }
    const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
    float v[8];
    load8(reinterpret_cast<const T*>(x.ptr) + p * x.ld + o, v);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] *= a[i] * inv;
    if (add) {
      float r[8];
      load8(add + p * add_ld + o, r);
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] += r[i];
    }
    store8(reinterpret_cast<T*>(y.ptr) + p * y.ld + o, v);
  }
}

// Persistent variant in the style of rowcol_gate_kernel: a thread keeps its channel octet and walks pixels with a grid stride, four packed loads of
// x (and of the residual) in flight per trip, 32-bit index arithmetic; the adaptive-pool bin bounds of every output row / column come from a
// table built once per CTA in shared memory, the attention vectors (ls x ls x c floats per image) are read through L1.  History
// (profiles/r2_pointwise.md): mlca_apply_kernel spends nine 64-bit integer divisions per 16 bytes (72 us for the 52 MB map of the P3 block); one
// short CTA per output row with the y-reduced bins in shared memory was no faster (64 us: 5120 CTAs of ~7 us, each a chain of dependent loads).
template <typename T>
__global__ void __launch_bounds__(256) mlca_apply2_kernel(yad_tensor x, const float* __restrict__ att, int ls, const T* __restrict__ add, int add_ld,
                                                          yad_tensor y) {
  extern __shared__ int tb[];  // [h] y bins, [w] x bins: first | (last + 1) << 16
  for (int i = threadIdx.x; i < x.h; i += blockDim.x) tb[i] = ((i * ls) / x.h) | ((((i + 1) * ls + x.h - 1) / x.h) << 16);
  for (int i = threadIdx.x; i < x.w; i += blockDim.x) tb[x.h + i] = ((i * ls) / x.w) | ((((i + 1) * ls + x.w - 1) / x.w) << 16);
  pdl_sync();
  __syncthreads();
  const int c = x.c, oct = c >> 3;
  const int o = ((int)threadIdx.x % oct) * 8, pstep = (int)(gridDim.x * blockDim.x) / oct, hw = x.h * x.w, npix = x.n * hw;
  const T* xp = reinterpret_cast<const T*>(x.ptr) + o;
  T* yp = reinterpret_cast<T*>(y.ptr) + o;
  const T* ap = add ? add + o : nullptr;
  auto apply = [&](int p, float (&v)[8], const float (&r)[8]) {
    const int n = p / hw, rem = p - n * hw, py = rem / x.w, px = rem - py * x.w;
    const int by = tb[py], bx = tb[x.h + px], y0 = by & 0xFFFF, y1 = by >> 16, x0 = bx & 0xFFFF, x1 = bx >> 16;
    const float* an = att + (int64_t)n * ls * ls * c + o;
    float4 t0 = __ldg(reinterpret_cast<const float4*>(an + (y0 * ls + x0) * c)), t1 = __ldg(reinterpret_cast<const float4*>(an + (y0 * ls + x0) * c + 4));
    const int cnt = (y1 - y0) * (x1 - x0);
    if (cnt != 1) {  // several bins under this pixel (maps that ls does not divide, or smaller than ls)
      t0 = make_float4(0.f, 0.f, 0.f, 0.f);
      t1 = t0;
      for (int yy = y0; yy < y1; yy++)
        for (int xx = x0; xx < x1; xx++) {
          const float4 u0 = __ldg(reinterpret_cast<const float4*>(an + (yy * ls + xx) * c)), u1 = __ldg(reinterpret_cast<const float4*>(an + (yy * ls + xx) * c + 4));
          t0.x += u0.x; t0.y += u0.y; t0.z += u0.z; t0.w += u0.w; t1.x += u1.x; t1.y += u1.y; t1.z += u1.z; t1.w += u1.w;
        }
      const float inv = 1.0f / (float)cnt;
      t0.x *= inv; t0.y *= inv; t0.z *= inv; t0.w *= inv; t1.x *= inv; t1.y *= inv; t1.z *= inv; t1.w *= inv;
    }
    v[0] = fmaf(v[0], t0.x, r[0]); v[1] = fmaf(v[1], t0.y, r[1]); v[2] = fmaf(v[2], t0.z, r[2]); v[3] = fmaf(v[3], t0.w, r[3]);
    v[4] = fmaf(v[4], t1.x, r[4]); v[5] = fmaf(v[5], t1.y, r[5]); v[6] = fmaf(v[6], t1.z, r[6]); v[7] = fmaf(v[7], t1.w, r[7]);
    store8(yp + (int64_t)p * y.ld, v);
  };
  int p = (int)(blockIdx.x * blockDim.x + threadIdx.x) / oct;
  for (; p + 3 * pstep < npix; p += 4 * pstep) {
    Raw8<T> rx[4], ra[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      raw_load(xp + (int64_t)(p + u * pstep) * x.ld, rx[u]);
      if (ap) raw_load(ap + (int64_t)(p + u * pstep) * add_ld, ra[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      float v[8], r[8];
      raw_unpack(rx[u], v);
      if (ap) raw_unpack(ra[u], r);
      else {
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = 0.f;
      }
      apply(p + u * pstep, v, r);
    }
  }
  for (; p < npix; p += pstep) {
    float v[8], r[8];
    load8(xp + (int64_t)p * x.ld, v);
    if (ap) load8(ap + (int64_t)p * add_ld, r);
    else {
#pragma unroll
      for (int i = 0; i < 8; i++) r[i] = 0.f;
    }
    apply(p, v, r);
  }
}

// adaptive_avg_pool2d(x, (h/s, w/s)) -> bilinear upsample to (h, w), align_corners=False (block.py:2451-2457)
template <typename T>
__global__ void pool_upsample_kernel(yad_tensor x, int s, yad_tensor y) {
  pdl_sync();
  const int oct = x.c >> 3;
  const int hp = x.h / s, wp = x.w / s;
  const float sy = (float)hp / (float)x.h, sx = (float)wp / (float)x.w;
  const int64_t total = (int64_t)x.n * x.h * x.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    int o = (int)(it % oct) * 8;
    int64_t p = it / oct;
    int px = (int)(p % x.w);
    int py = (int)((p / x.w) % x.h);
    int n = (int)(p / ((int64_t)x.w * x.h));
    float fy = fmaxf(sy * ((float)py + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)px + 0.5f) - 0.5f, 0.f);
    int iy0 = (int)fy, ix0 = (int)fx;
    int iy1 = min(iy0 + 1, hp - 1), ix1 = min(ix0 + 1, wp - 1);
    float ly = fy - (float)iy0, lx = fx - (float)ix0;
    float out[8];
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = 0.f;
    for (int q = 0; q < 4; q++) {
      int by = (q >> 1) ? iy1 : iy0, bx = (q & 1) ? ix1 : ix0;
      float wgt = ((q >> 1) ? ly : 1.f - ly) * ((q & 1) ? lx : 1.f - lx);
      if (wgt == 0.f) continue;
      const int y0 = bin_start(by, x.h, hp), y1 = bin_end(by, x.h, hp), x0 = bin_start(bx, x.w, wp), x1 = bin_end(bx, x.w, wp);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = 0.f;
      for (int yy = y0; yy < y1; yy++)
        for (int xx = x0; xx < x1; xx++) {
          float v[8];
          load8(reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, yy, xx) + o, v);
#pragma unroll
          for (int i = 0; i < 8; i++) acc[i] += v[i];
        }
      const float sc = wgt / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
      for (int i = 0; i < 8; i++) out[i] = fmaf(acc[i], sc, out[i]);
    }
    store8(reinterpret_cast<T*>(y.ptr) + p * y.ld + o, out);
  }
}

// The same operator in two phases inside one CTA = (image, 32 channels): bin sums of the pooled map into shared memory (every input pixel read
// once), then the bilinear blend from shared memory.  pool_upsample_kernel re-sums up to four s x s bins for every output pixel (16 - 64 global
// loads each).  Same arithmetic in the same order (raw bin sums, weight / count applied in the blend): bit-identical results.
constexpr int PU_CC = 32;
template <typename T>
__global__ void __launch_bounds__(256) pool_upsample_smem_kernel(yad_tensor x, int s, yad_tensor y) {
  extern __shared__ __align__(16) float pu_sm[];  // [hp * wp][32] raw bin sums
  pdl_sync();
  const int chunks = x.c / PU_CC, n = blockIdx.x / chunks, c0 = (blockIdx.x - n * chunks) * PU_CC;
  const int hp = x.h / s, wp = x.w / s;
  for (int it = threadIdx.x; it < hp * wp * 4; it += 256) {
    const int oi = it & 3, bin = it >> 2, by = bin / wp, bx = bin - by * wp;
    const int y0 = bin_start(by, x.h, hp), y1 = bin_end(by, x.h, hp), x0 = bin_start(bx, x.w, wp), x1 = bin_end(bx, x.w, wp);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    for (int yy = y0; yy < y1; yy++)
      for (int xx = x0; xx < x1; xx++) {
        float v[8];
        load8(reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, yy, xx) + c0 + oi * 8, v);
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] += v[i];
      }
    store8(pu_sm + bin * PU_CC + oi * 8, acc);
  }
  __syncthreads();
  const float sy = (float)hp / (float)x.h, sx = (float)wp / (float)x.w;
  for (int it = threadIdx.x; it < x.h * x.w * 4; it += 256) {
    const int oi = it & 3, p = it >> 2, py = p / x.w, px = p - py * x.w;
    const float fy = fmaxf(sy * ((float)py + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)px + 0.5f) - 0.5f, 0.f);
    const int iy0 = (int)fy, ix0 = (int)fx;
    const int iy1 = min(iy0 + 1, hp - 1), ix1 = min(ix0 + 1, wp - 1);
    const float ly = fy - (float)iy0, lx = fx - (float)ix0;
    float out[8];
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int by = (q >> 1) ? iy1 : iy0, bx = (q & 1) ? ix1 : ix0;
      const float wgt = ((q >> 1) ? ly : 1.f - ly) * ((q & 1) ? lx : 1.f - lx);
      if (wgt == 0.f) continue;
      const int cnt = (bin_end(by, x.h, hp) - bin_start(by, x.h, hp)) * (bin_end(bx, x.w, wp) - bin_start(bx, x.w, wp));
      const float sc = wgt / (float)cnt;
      float acc[8];
      load8(pu_sm + (by * wp + bx) * PU_CC + oi * 8, acc);
#pragma unroll
      for (int i = 0; i < 8; i++) out[i] = fmaf(acc[i], sc, out[i]);
    }
    store8(reinterpret_cast<T*>(y.ptr) + pix_off(y, n, py, px) + c0 + oi * 8, out);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// tiny per-image MLPs and AdaptiveDynamicTanh
// ------------------------------------------------------------------------------------------------------------------
__global__ void gate_mlp_kernel(const float* __restrict__ g, const float* __restrict__ w1, const float* __restrict__ b1,
                                const float* __restrict__ w2, const float* __restrict__ b2, int c, int hidden, int nout, int kind,
                                float* __restrict__ out) {
  pdl_sync();
  extern __shared__ float sm[];  // hid[hidden], o[nout]
  float* hid = sm;
  float* o = sm + hidden;
  const int n = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int h = wid; h < hidden; h += nw) {
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s = fmaf(w1[h * c + i], g[(int64_t)n * c + i], s);
    s = warp_sum(s);
    if (lane == 0) hid[h] = fmaxf(s + b1[h], 0.f);
  }
  __syncthreads();
  for (int j = wid; j < nout; j += nw) {
    float s = 0.f;
    for (int i = lane; i < hidden; i += 32) s = fmaf(w2[j * hidden + i], hid[i], s);
    s = warp_sum(s);
    if (lane == 0) o[j] = s + b2[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (kind == 0) {
      for (int j = 0; j < nout; j++) out[(int64_t)n * nout + j] = sigmoidf_(o[j]);
    } else {
      float m = -INFINITY, sum = 0.f;
      for (int j = 0; j < nout; j++) m = fmaxf(m, o[j]);
      for (int j = 0; j < nout; j++) sum += expf(o[j] - m);
      for (int j = 0; j < nout; j++) out[(int64_t)n * nout + j] = expf(o[j] - m) / sum;
    }
  }
}

template <typename T>
__global__ void adt_apply_kernel(yad_tensor x, const float* __restrict__ imp, const float* __restrict__ alphas, const float* __restrict__ weight,
                                 const float* __restrict__ bias, yad_tensor y) {
  pdl_sync();
  const int oct = x.c >> 3;
  const int64_t hw = (int64_t)x.h * x.w, total = (int64_t)x.n * hw * oct;
  const float a0 = alphas[0], a1 = alphas[1], a2 = alphas[2];
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    int o = (int)(it % oct) * 8;
    int64_t p = it / oct;
    int n = (int)(p / hw);
    const float i0 = imp[n * 3], i1 = imp[n * 3 + 1], i2 = imp[n * 3 + 2];
    float v[8], wv[8], bv[8];
    load8(reinterpret_cast<const T*>(x.ptr) + p * x.ld + o, v);
    load8(weight + o, wv);
    load8(bias + o, bv);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      float t = tanhf(a0 * v[i]) * i0 + tanhf(a1 * v[i]) * i1 + tanhf(a2 * v[i]) * i2;
      v[i] = fmaf(t, wv[i], bv[i]);
    }
    store8(reinterpret_cast<T*>(y.ptr) + p * y.ld + o, v);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// generic elementwise combine
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void eltwise_kernel(int op, yad_tensor a, const T* __restrict__ b, int b_ld, const T* __restrict__ c3, int c3_ld,
                               const T* __restrict__ d4, int d4_ld, float alpha, float beta, float gamma, const float* __restrict__ pa,
                               const float* __restrict__ pb, const float* __restrict__ pg, yad_tensor y) {
  pdl_sync();
  if (pa) alpha = *pa;  // coefficients that are model parameters stay on the device (training path)
  if (pb) beta = *pb;
  if (pg) gamma = *pg;
  const int oct = a.c >> 3;
  const int64_t total = (int64_t)a.n * a.h * a.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    int o = (int)(it % oct) * 8;
    int64_t p = it / oct;
    float va[8], vb[8], vc[8];
    load8(reinterpret_cast<const T*>(a.ptr) + p * a.ld + o, va);
    if (op == 2 || op == 8) {
      float s = ld1(b + p * b_ld);
#pragma unroll
      for (int i = 0; i < 8; i++) va[i] *= s;
      if (op == 8) {
        load8(c3 + p * c3_ld + o, vc);
#pragma unroll
        for (int i = 0; i < 8; i++) va[i] += vc[i];
      }
    } else if (op == 5) {
#pragma unroll
      for (int i = 0; i < 8; i++) va[i] *= alpha;
      if (c3) {
        load8(c3 + p * c3_ld + o, vc);
#pragma unroll
        for (int i = 0; i < 8; i++) va[i] += vc[i];
      }
    } else {
      load8(b + p * b_ld + o, vb);
      if (op == 1) {
#pragma unroll
        for (int i = 0; i < 8; i++) va[i] *= vb[i];
      } else if (op == 4 || op == 6) {
        const float sc = op == 6 ? alpha : 1.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) va[i] *= vb[i] * sc;
        if (c3) {
          load8(c3 + p * c3_ld + o, vc);
#pragma unroll
          for (int i = 0; i < 8; i++) va[i] += vc[i];
        }
      } else if (op == 7) {
#pragma unroll
        for (int i = 0; i < 8; i++) va[i] = apply_act(va[i], YAD_ACT_GELU) * vb[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) va[i] = alpha * va[i] + beta * vb[i];
        if (op == 3) {
          float vd[8];
          load8(c3 + p * c3_ld + o, vc);
          load8(d4 + p * d4_ld + o, vd);
#pragma unroll
          for (int i = 0; i < 8; i++) va[i] += gamma * vc[i] + vd[i];
        }
      }
    }
    store8(reinterpret_cast<T*>(y.ptr) + p * y.ld + o, va);
  }
}

// mean over s token groups: x (n,1,s*T,c) -> y (n,1,T,c)
template <typename T>
__global__ void group_mean_kernel(yad_tensor x, int s, yad_tensor y) {
  pdl_sync();
  const int oct = y.c >> 3, Tn = y.w * y.h;
  const int64_t total = (int64_t)y.n * Tn * oct;
  const float inv = 1.0f / (float)s;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    int o = (int)(it % oct) * 8;
    int64_t p = it / oct;
    int64_t n = p / Tn, t = p - n * Tn;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    for (int g = 0; g < s; g++) {
      float v[8];
      load8(reinterpret_cast<const T*>(x.ptr) + ((n * s + g) * Tn + t) * x.ld + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] *= inv;
    store8(reinterpret_cast<T*>(y.ptr) + p * y.ld + o, acc);
  }
}

// EDFFN 8x8 patch spectral filter as a per-channel 64x64 matrix (block.py:2398-2413); grid (patches, n), block = c threads
template <typename T>
__global__ void patch_filter_kernel(yad_tensor x, const float* __restrict__ m, float alpha, const T* __restrict__ add, int add_ld, yad_tensor y) {
  pdl_sync();
  const int c = x.c, n = blockIdx.y;
  const int wp = (x.w + 7) / 8;
  const int pr = blockIdx.x / wp, pc = blockIdx.x % wp;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float in[64];
#pragma unroll
    for (int i = 0; i < 64; i++) {
      int yy = pr * 8 + (i >> 3), xx = pc * 8 + (i & 7);
      if (yy >= x.h) yy = 2 * x.h - 2 - yy;  // F.pad(..., mode='reflect') on the bottom / right edge
      if (xx >= x.w) xx = 2 * x.w - 2 - xx;
      in[i] = ld1(reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, yy, xx) + ch);
    }
    for (int o = 0; o < 64; o++) {
      int yy = pr * 8 + (o >> 3), xx = pc * 8 + (o & 7);
      if (yy >= x.h || xx >= x.w) continue;
      const float* mo = m + (int64_t)o * 64 * c + ch;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i++) s = fmaf(mo[(int64_t)i * c], in[i], s);
      int64_t p = ((int64_t)n * x.h + yy) * x.w + xx;
      float r = alpha * s;
      if (add) r += ld1(add + p * add_ld + ch);
      st1(reinterpret_cast<T*>(y.ptr) + p * y.ld + ch, r);
    }
  }
}

// The same filter with the matrix resident in shared memory: a CTA owns 8 channels (M[:, :, 8 ch] = 128 KB fp32, read from L2 once) and 64 patches
// taken across the whole batch; a thread = one patch x one channel pair, its 64 x 2 inputs in registers, the matrix read as broadcast LDS.64 (one
// wavefront per warp).  patch_filter_kernel re-reads the full 2 MB matrix for every (image, patch): 1.2 GB through L2 at batch 64 -- the L2
// bandwidth, not the 0.6 GFLOP, set its 168 us.  grid (ceil(n * patches / 64), c / 8), 256 threads.
template <typename T>
__global__ void __launch_bounds__(256) patch_filter_smem_kernel(yad_tensor x, const float* __restrict__ m, float alpha, const T* __restrict__ add,
                                                                int add_ld, yad_tensor y, int patches_total) {
  extern __shared__ __align__(16) float pf_ms[];  // [64 * 64][8]
  const int c = x.c, c0 = blockIdx.y * 8;
  for (int idx = threadIdx.x; idx < 64 * 64 * 2; idx += 256) {
    const int oi = idx >> 1, half = idx & 1;
    reinterpret_cast<float4*>(pf_ms)[idx] = *reinterpret_cast<const float4*>(m + (int64_t)oi * c + c0 + half * 4);
  }
  pdl_sync();
  const int cp = threadIdx.x & 3, gp = blockIdx.x * 64 + (threadIdx.x >> 2);
  const int wp = (x.w + 7) / 8, ppi = ((x.h + 7) / 8) * wp;
  const bool live = gp < patches_total;
  const int n = live ? gp / ppi : 0, pt = live ? gp - n * ppi : 0, pr = pt / wp, pc = pt - pr * wp;
  const int ch = c0 + 2 * cp;
  float in[64][2];
  if (live) {
#pragma unroll
    for (int i = 0; i < 64; i++) {
      int yy = pr * 8 + (i >> 3), xx = pc * 8 + (i & 7);
      if (yy >= x.h) yy = 2 * x.h - 2 - yy;  // F.pad(..., mode='reflect') on the bottom / right edge
      if (xx >= x.w) xx = 2 * x.w - 2 - xx;
      const T* src = reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, yy, xx) + ch;
      in[i][0] = ld1(src);
      in[i][1] = ld1(src + 1);
    }
  }
  __syncthreads();
  if (!live) return;
  const float* ms = pf_ms + 2 * cp;
  for (int o = 0; o < 64; o++) {
    const int yy = pr * 8 + (o >> 3), xx = pc * 8 + (o & 7);
    if (yy >= x.h || xx >= x.w) continue;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 64; i++) {
      const float2 mv = *reinterpret_cast<const float2*>(ms + (o * 64 + i) * 8);
      s0 = fmaf(mv.x, in[i][0], s0);
      s1 = fmaf(mv.y, in[i][1], s1);
    }
    const int64_t p = ((int64_t)n * x.h + yy) * x.w + xx;
    float r0 = alpha * s0, r1 = alpha * s1;
    if (add) { r0 += ld1(add + p * add_ld + ch); r1 += ld1(add + p * add_ld + ch + 1); }
    T* dst = reinterpret_cast<T*>(y.ptr) + p * y.ld + ch;
    st1(dst, r0);
    st1(dst + 1, r1);
  }
}

int grid_for(int64_t items, int tpb = TPB) {
  int64_t g = (items + tpb - 1) / tpb;
  const int64_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int check_view(const yad_tensor* t, const char* what) {
  YAD_CHECK(t && t->ptr, "%s: null tensor", what);
  YAD_CHECK(t->c % 8 == 0 && t->ld % 8 == 0 && t->ld >= t->c, "%s: channels (%d) and ld (%d) must be multiples of 8, ld >= c", what, t->c, t->ld);
  YAD_CHECK(((uintptr_t)t->ptr & 15) == 0, "%s: pointer must be 16-byte aligned", what);
  return 0;
}
#define CHECK_VIEW(t, what) \
  do {                      \
    if (check_view(t, what)) return 1; \
  } while (0)
#define SAME_SHAPE(a, b, what) YAD_CHECK((a)->n == (b)->n && (a)->h == (b)->h && (a)->w == (b)->w, "%s: shape mismatch", what)

}  // namespace

extern "C" {

int yad_nchw_to_nhwc(const float* src, int c_src, const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(y, "nchw_to_nhwc");
  YAD_CHECK(c_src <= y->c, "nchw_to_nhwc: c_src %d > c %d", c_src, y->c);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)y->n * y->h * y->w;
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(nchw_to_nhwc_kernel<T>, grid_for(total), TPB, 0, st, src, c_src, *y);)
  YAD_LAUNCH_CHECK("nchw_to_nhwc");
  return 0;
}

int yad_u8_to_nhwc(const uint8_t* src, int c_src, const yad_tensor* y, float scale, int dtype, void* stream) {
  CHECK_VIEW(y, "u8_to_nhwc");
  YAD_CHECK(c_src <= y->c, "u8_to_nhwc: c_src %d > c %d", c_src, y->c);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)y->n * y->h * y->w;
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(u8_to_nhwc_kernel<T>, grid_for(total), TPB, 0, st, src, c_src, scale, *y);)
  YAD_LAUNCH_CHECK("u8_to_nhwc");
  return 0;
}

int yad_stem_conv(const void* img, int img_is_u8, int n, int h, int w, int cin, const float* wgt, const float* bias, int act,
                  const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(y, "stem_conv");
  YAD_CHECK(y->c == 16, "stem_conv: built for 16 output channels (layer 0 of the yaml at scale n), got %d", y->c);
  YAD_CHECK(y->n == n && y->h == (h + 2 - 3) / 2 + 1 && y->w == (w + 2 - 3) / 2 + 1, "stem_conv: output shape does not match a k3 s2 p1 convolution");
  YAD_CHECK(cin >= 1 && cin <= 4, "stem_conv: cin %d unsupported", cin);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(9 * cin * 16 + 16 + cin * 3 * 257) * sizeof(float);
  const dim3 grid((y->w + 127) / 128, y->h, n);
  YAD_CHECK(y->h <= 65535 && n <= 65535, "stem_conv: grid too large");
  static int stem_tc = -1;
  if (stem_tc < 0) { const char* ev = getenv("YAD_STEM_TC"); stem_tc = (ev && ev[0] == '0') ? 0 : 1; }
  if (stem_tc && img_is_u8 && dtype == YAD_BF16 && cin == 3 && (w & 15) == 0 && ((uintptr_t)img & 15) == 0 && (y->ld % 8) == 0) {
    const dim3 grid_tc((y->w + 127) / 128, (y->h + STEM_ROWS - 1) / STEM_ROWS, n);
    YAD_LAUNCH(stem_tc_kernel, grid_tc, 128, 0, st, (const uint8_t*)img, h, w, wgt, bias, act, *y);
    YAD_LAUNCH_CHECK("stem_conv");
    return 0;
  }
  YAD_DISPATCH_DTYPE(dtype, {
    if (img_is_u8)
      YAD_LAUNCH((stem_conv_kernel<T, uint8_t, 16>), grid, 128, smem, st, (const uint8_t*)img, n, h, w, cin, wgt, bias, act, *y);
    else
      YAD_LAUNCH((stem_conv_kernel<T, float, 16>), grid, 128, smem, st, (const float*)img, n, h, w, cin, wgt, bias, act, *y);
  })
  YAD_LAUNCH_CHECK("stem_conv");
  return 0;
}

int yad_gn_stats(const yad_tensor* x, int groups, double* stats, int dtype, void* stream) {
  CHECK_VIEW(x, "gn_stats");
  YAD_CHECK(groups > 0 && x->c % groups == 0, "gn_stats: %d channels not divisible into %d groups", x->c, groups);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(stats, 0, sizeof(double) * 2 * groups * x->n, st);
  int64_t hw = (int64_t)x->h * x->w;
  int chunks = (int)((hw * (x->c / 8) + TPB * 8 - 1) / (TPB * 8));
  // the grid is sized for the whole launch: a batch-statistics BatchNorm arrives as ONE image (n = 1) and still has to fill 148 SMs
  const int cap_s = 1184 / x->n > 1 ? 1184 / x->n : 1;
  chunks = chunks < 1 ? 1 : (chunks > cap_s ? cap_s : chunks);
  dim3 grid(chunks, x->n);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(gn_stats_kernel<T>, grid, TPB, 2 * x->c * sizeof(float), st, *x, groups, stats);)
  YAD_LAUNCH_CHECK("gn_stats");
  return 0;
}

int yad_gn_apply(const yad_tensor* x, const double* stats, int groups, const float* gamma, const float* beta, float eps, int act,
                 const void* add, int add_ld, const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(x, "gn_apply x");
  CHECK_VIEW(y, "gn_apply y");
  SAME_SHAPE(x, y, "gn_apply");
  YAD_CHECK(x->c == y->c && x->c % groups == 0, "gn_apply: channel mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t items = (int64_t)x->h * x->w * (x->c / 8);
  int gx = (int)((items + TPB * 4 - 1) / (TPB * 4));
  static int gn_cap = 0;
  if (!gn_cap) { const char* ev = getenv("YAD_GN_CAP"); gn_cap = ev ? atoi(ev) : 592; }  // 4 CTAs per SM: with four packed loads in flight per thread that saturates HBM, and the per-CTA prologue is paid once
  const int cap_a = gn_cap / x->n > 1 ? gn_cap / x->n : 1;
  gx = gx < 1 ? 1 : (gx > cap_a ? cap_a : gx);
  dim3 grid(gx, x->n);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(gn_apply_kernel<T>, grid, TPB, 2 * x->c * sizeof(float), st, *x, stats, groups, gamma, beta, eps, act,
                                                                                               (const T*)add, add_ld, *y);)
  YAD_LAUNCH_CHECK("gn_apply");
  return 0;
}

int yad_dwconv(const yad_tensor* x, const float* w, const float* bias, const float* scale, const float* shift, int k, int act,
               int gate_split, const void* add, int add_ld, const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(x, "dwconv x");
  CHECK_VIEW(y, "dwconv y");
  SAME_SHAPE(x, y, "dwconv");
  YAD_CHECK(k == 3 || k == 5 || k == 7, "dwconv: k=%d unsupported", k);
  YAD_CHECK(gate_split > 0 ? (x->c == 2 * gate_split && y->c == gate_split) : (x->c == y->c), "dwconv: channel mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)y->n * y->h * y->w * (y->c / 8);
  static int row_env = -1;
  if (row_env < 0) { const char* ev = getenv("YAD_DWCONV_ROW"); row_env = ev ? atoi(ev) : 1; }  // 0: dwconv_kernel, 1: shared-memory tiles, 2: dwconv_row_kernel
  if (row_env && gate_split == 0 && total > 0 && ((((uintptr_t)w) | ((uintptr_t)bias) | ((uintptr_t)scale) | ((uintptr_t)shift)) & 15) == 0) {
    const int cc = 64 / (dtype == YAD_BF16 ? 2 : 4);
    if (row_env != 2 && y->c % cc == 0) {
      const int th = dws_tile(y->h), tw = dws_tile(y->w);
      const int tiles_y = (y->h + th - 1) / th, tiles_x = (y->w + tw - 1) / tw;
      const int64_t ctas = (int64_t)y->n * tiles_y * tiles_x * (y->c / cc);
      const size_t smem = (size_t)k * k * cc * sizeof(float) + (size_t)(th + k - 1) * ((tw + DWS_PX - 1) / DWS_PX * DWS_PX + k - 1) * 80;
      // threads: the tile's items (8 channels x 4 pixels) spread evenly over the fewest passes of at most 256 threads
      const int items = th * ((tw + DWS_PX - 1) / DWS_PX) * (cc / 8), passes = (items + DWS_THREADS - 1) / DWS_THREADS;
      const int threads = ((items + passes - 1) / passes + 31) / 32 * 32;
      if (ctas < ((int64_t)1 << 31)) {
#define DW_SMEM(KK)                                                                                                               \
  YAD_DISPATCH_DTYPE(dtype, {                                                                                                     \
    static bool attr = false;                                                                                                     \
    if (!attr) { cudaFuncSetAttribute(dwconv_smem_kernel<T, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; } \
    YAD_LAUNCH((dwconv_smem_kernel<T, KK>), (unsigned)ctas, threads, smem, st, *x, w, bias, scale, shift, act, (const T*)add, add_ld, *y, th, \
               tw, tiles_x, tiles_y);                                                                                             \
  })
        if (k == 3) { DW_SMEM(3) } else if (k == 5) { DW_SMEM(5) } else { DW_SMEM(7) }
#undef DW_SMEM
        YAD_LAUNCH_CHECK("dwconv");
        return 0;
      }
    }
    // PX adjacent outputs per thread, as many as still leave >= 16 warps per SM (small maps keep their parallelism)
    const int64_t rows = (int64_t)y->n * y->h * (y->c / 8);
    int px = 4;
    while (px > 1 && rows * ((y->w + px - 1) / px) < (int64_t)148 * 16 * 32) px >>= 1;
    const int groups_x = (y->w + px - 1) / px;
    const int64_t items = rows * groups_x;
    if (px > 1 && items < ((int64_t)1 << 31)) {
#define DW_ROW(KK, PP)                                                                                                          \
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH((dwconv_row_kernel<T, KK, PP>), (unsigned)((items + 127) / 128), 128, 0, st, *x, w, bias, scale, shift, \
                                       act, (const T*)add, add_ld, *y, groups_x, (int)items);)
      if (px == 4) {
        if (k == 3) { DW_ROW(3, 4) } else if (k == 5) { DW_ROW(5, 4) } else { DW_ROW(7, 4) }
      } else {
        if (k == 3) { DW_ROW(3, 2) } else if (k == 5) { DW_ROW(5, 2) } else { DW_ROW(7, 2) }
      }
#undef DW_ROW
      YAD_LAUNCH_CHECK("dwconv");
      return 0;
    }
  }
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(dwconv_kernel<T>, grid_for(total, 128), 128, 0, st, *x, w, bias, scale, shift, k, act, gate_split,
                                                                                    (const T*)add, add_ld, *y);)
  YAD_LAUNCH_CHECK("dwconv");
  return 0;
}

int yad_sppf_pool(const yad_tensor* x, const yad_tensor* y1, const yad_tensor* y2, const yad_tensor* y3, int dtype, void* stream) {
  CHECK_VIEW(x, "sppf x");
  CHECK_VIEW(y1, "sppf y1");
  CHECK_VIEW(y2, "sppf y2");
  CHECK_VIEW(y3, "sppf y3");
  SAME_SHAPE(x, y1, "sppf");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)x->n * x->h * x->w * (x->c / 8);
  static int sp_tile = -1;
  if (sp_tile < 0) { const char* ev = getenv("YAD_SPPF_TILE"); sp_tile = (ev && ev[0] == '0') ? 0 : 1; }
  if (sp_tile && dtype == YAD_BF16 && x->c % SP_CC == 0 && total > 0) {
    // tile edge: the map edge when it fits, else the edge <= 20 with the fewest tiles
    auto edge = [](int n) { const int t = (n + SP_TMAX - 1) / SP_TMAX; return (n + t - 1) / t; };
    const int T = edge(x->h) > edge(x->w) ? edge(x->h) : edge(x->w);
    const int tiles_y = (x->h + T - 1) / T, tiles_x = (x->w + T - 1) / T, halo = (tiles_y == 1 && tiles_x == 1) ? 0 : SP_HALO, TD = T + 2 * halo;
    const int64_t ctas = (int64_t)x->n * tiles_y * tiles_x * (x->c / SP_CC);
    const size_t smem = (size_t)TD * TD * 2 * 2 * sizeof(uint4);
    if (ctas < ((int64_t)1 << 31)) {
      static bool attr = false;
      if (!attr) { cudaFuncSetAttribute(sppf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 32 * 2 * 2 * 16); attr = true; }
      YAD_LAUNCH(sppf_tile_kernel, (unsigned)ctas, 256, smem, st, *x, *y1, *y2, *y3, T, tiles_x, tiles_y, halo);
      YAD_LAUNCH_CHECK("sppf_pool");
      return 0;
    }
  }
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(sppf_pool_kernel<T>, grid_for(total, 128), 128, 0, st, *x, *y1, *y2, *y3);)
  YAD_LAUNCH_CHECK("sppf_pool");
  return 0;
}

int yad_gap(const yad_tensor* x, float* out, int dtype, void* stream) {
  CHECK_VIEW(x, "gap");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(float) * x->n * x->c, st);
  int64_t hw = (int64_t)x->h * x->w;
  int splits = (int)((hw * (x->c / 8) + TPB * 8 - 1) / (TPB * 8));
  splits = splits < 1 ? 1 : (splits > 32 ? 32 : splits);
  dim3 grid(splits, x->n);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(gap_kernel<T>, grid, TPB, x->c * sizeof(float), st, *x, out);)
  YAD_LAUNCH_CHECK("gap");
  return 0;
}

int yad_ln_mix(const yad_tensor* x, const float* weight, const float* bias, const float* gamma, const float* gammax, float eps,
               const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(x, "ln_mix x");
  CHECK_VIEW(y, "ln_mix y");
  SAME_SHAPE(x, y, "ln_mix");
  YAD_CHECK(weight && bias && gamma && gammax, "ln_mix: null parameter");
  YAD_CHECK(x->c <= 1024, "ln_mix: at most 1024 channels (got %d)", x->c);
  int g = 1;
  while (g < 32 && g * 8 < x->c) g <<= 1;
  const int64_t npix = (int64_t)x->n * x->h * x->w;
  if (npix == 0) return 0;
  const int64_t per_block = 8 * (32 / g);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(ln_mix_kernel<T>, cdiv(npix, per_block), 256, 0, (cudaStream_t)stream, *x, weight, bias, gamma, gammax, eps, *y, g);)
  YAD_LAUNCH_CHECK("ln_mix");
  return 0;
}

int yad_rowcol_mean(const yad_tensor* x, const yad_tensor* rowmean, const yad_tensor* colmean, int dtype, void* stream) {
  CHECK_VIEW(x, "rowcol_mean x");
  CHECK_VIEW(rowmean, "rowcol_mean rows");
  CHECK_VIEW(colmean, "rowcol_mean cols");
  YAD_CHECK(rowmean->c == x->c && colmean->c == x->c, "rowcol_mean: channel mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  static int rc1 = -1;
  if (rc1 < 0) { const char* ev = getenv("YAD_ROWCOL_1PASS"); rc1 = (ev && ev[0] == '0') ? 0 : 1; }
  const size_t rc_smem = ((size_t)x->h * 16 + (size_t)3 * x->w * 16) * sizeof(float);
  if (rc1 && x->c % 16 == 0 && x->w <= 160 && rc_smem <= 48 * 1024 && (rowmean->ld % 8) == 0 && (colmean->ld % 8) == 0) {
    const unsigned g1 = (unsigned)(x->n * (x->c / 16));
    const int nk = (x->w + 31) / 32;
    YAD_DISPATCH_DTYPE(dtype, {
      if (nk <= 1) YAD_LAUNCH((rowcol_mean1_kernel<T, 1>), g1, 256, rc_smem, st, *x, *rowmean, *colmean);
      else if (nk == 2) YAD_LAUNCH((rowcol_mean1_kernel<T, 2>), g1, 256, rc_smem, st, *x, *rowmean, *colmean);
      else if (nk == 3) YAD_LAUNCH((rowcol_mean1_kernel<T, 3>), g1, 256, rc_smem, st, *x, *rowmean, *colmean);
      else YAD_LAUNCH((rowcol_mean1_kernel<T, 5>), g1, 256, rc_smem, st, *x, *rowmean, *colmean);
    })
    YAD_LAUNCH_CHECK("rowcol_mean");
    return 0;
  }
  dim3 grid(x->h > x->w ? x->h : x->w, x->n, 2);
  int tpb = (x->c / 8) * 8;  // 8 pixel lanes per octet
  tpb = tpb < 64 ? 64 : (tpb > 256 ? 256 : tpb);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(rowcol_mean_kernel<T>, grid, tpb, x->c * sizeof(float), st, *x, *rowmean, *colmean);)
  YAD_LAUNCH_CHECK("rowcol_mean");
  return 0;
}

int yad_rowcol_gate(const yad_tensor* x, const yad_tensor* gh, const yad_tensor* gw, const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(y, "rowcol_gate y");
  CHECK_VIEW(gh, "rowcol_gate gh");
  CHECK_VIEW(gw, "rowcol_gate gw");
  if (x) { CHECK_VIEW(x, "rowcol_gate x"); SAME_SHAPE(x, y, "rowcol_gate"); }
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)y->n * y->h * y->w * (y->c / 8);
  yad_tensor xx = x ? *x : *y;
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(rowcol_gate_kernel<T>, grid_for(total), TPB, 0, st, xx, x != nullptr, *gh, *gw, *y);)
  YAD_LAUNCH_CHECK("rowcol_gate");
  return 0;
}

// CoordAtt's gate MLP on the pooled rows / columns (nn/modules/head.py:694-703): one warp per position of the concatenated (h + w) axis.
int yad_coordatt_mlp(const yad_tensor* rows, const yad_tensor* cols, const float* w1, const float* b1, int mip, const float* wh, const float* bh,
                     const float* ww, const float* bw, const yad_tensor* gh, const yad_tensor* gw, int dtype, void* stream) {
  CHECK_VIEW(rows, "coordatt_mlp rows");
  CHECK_VIEW(cols, "coordatt_mlp cols");
  CHECK_VIEW(gh, "coordatt_mlp gh");
  CHECK_VIEW(gw, "coordatt_mlp gw");
  YAD_CHECK(w1 && b1 && wh && bh && ww && bw, "coordatt_mlp: null weights");
  YAD_CHECK(rows->w == 1 && cols->w == 1 && gh->w == 1 && gw->w == 1, "coordatt_mlp: (n, L, 1, c) views expected");
  YAD_CHECK(rows->n == cols->n && rows->c == cols->c && gh->n == rows->n && gh->h == rows->h && gw->n == cols->n && gw->h == cols->h &&
                gh->c == gw->c,
            "coordatt_mlp: shape mismatch");
  YAD_CHECK(mip == 8 || mip == 16 || mip == 32, "coordatt_mlp: %d hidden channels (8, 16 or 32 supported)", mip);
  YAD_CHECK(rows->c <= 1024 && gh->c <= 1024, "coordatt_mlp: at most 1024 channels");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t positions = (int64_t)rows->n * (rows->h + cols->h);
  const int wpb = TPB / 32;
  const int grid = (int)((positions + wpb - 1) / wpb);
  if (grid == 0) return 0;
#define CA_LAUNCH(M) YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH((coordatt_mlp_kernel<T, M>), grid, TPB, 0, st, *rows, *cols, w1, b1, wh, bh, ww, bw, *gh, *gw);)
  if (mip == 8) { CA_LAUNCH(8) } else if (mip == 16) { CA_LAUNCH(16) } else { CA_LAUNCH(32) }
#undef CA_LAUNCH
  YAD_LAUNCH_CHECK("coordatt_mlp");
  return 0;
}

int yad_pool_upsample(const yad_tensor* x, int s, const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(x, "pool_upsample x");
  CHECK_VIEW(y, "pool_upsample y");
  SAME_SHAPE(x, y, "pool_upsample");
  YAD_CHECK(s >= 1 && x->h / s >= 1 && x->w / s >= 1, "pool_upsample: scale %d too large for %dx%d", s, x->h, x->w);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)x->n * x->h * x->w * (x->c / 8);
  const size_t pooled = (size_t)(x->h / s) * (x->w / s) * PU_CC * sizeof(float);
  static int pu_smem = -1;
  if (pu_smem < 0) { const char* ev = getenv("YAD_POOL_UPSAMPLE_SMEM"); pu_smem = (ev && ev[0] == '0') ? 0 : 1; }
  if (pu_smem && x->c % PU_CC == 0 && pooled <= 64 * 1024 && (int64_t)x->n * (x->c / PU_CC) >= 64 && total > 0) {  // enough CTAs to fill the GPU
    YAD_DISPATCH_DTYPE(dtype, {
      static bool attr = false;
      if (!attr) { cudaFuncSetAttribute(pool_upsample_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); attr = true; }
      YAD_LAUNCH(pool_upsample_smem_kernel<T>, (unsigned)(x->n * (x->c / PU_CC)), 256, pooled, st, *x, s, *y);
    })
    YAD_LAUNCH_CHECK("pool_upsample");
    return 0;
  }
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(pool_upsample_kernel<T>, grid_for(total, 128), 128, 0, st, *x, s, *y);)
  YAD_LAUNCH_CHECK("pool_upsample");
  return 0;
}

int yad_mlca_pool(const yad_tensor* x, float* local, int local_size, int dtype, void* stream) {
  CHECK_VIEW(x, "mlca_pool");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(local_size * local_size, x->n);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(mlca_pool_kernel<T>, grid, 128, x->c * sizeof(float), st, *x, local, local_size);)
  YAD_LAUNCH_CHECK("mlca_pool");
  return 0;
}

int yad_mlca_att(const float* local, const float* w_global, const float* w_local, int ksize, float local_weight, int n, int c,
                 int local_size, float* att, float* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int nb = local_size * local_size;
  size_t smem = (size_t)(nb * c + local_size * c) * sizeof(float);
  YAD_CHECK(smem <= 48 * 1024, "mlca_att: %d channels need %zu B of shared memory", c, smem);
  YAD_CHECK(scratch != nullptr, "mlca_att: scratch (fp32 [n][c]) is required");
  YAD_LAUNCH(mlca_glob_kernel, n, 128, c * sizeof(float), st, local, w_global, ksize, c, local_size, scratch);
  YAD_LAUNCH(mlca_att_kernel, n, 256, smem, st, local, scratch, w_local, ksize, local_weight, c, local_size, n, att);
  YAD_LAUNCH_CHECK("mlca_att");
  return 0;
}

int yad_mlca_apply(const yad_tensor* x, const float* att, int local_size, const void* add, int add_ld, const yad_tensor* y, int dtype,
                   void* stream) {
  CHECK_VIEW(x, "mlca_apply x");
  CHECK_VIEW(y, "mlca_apply y");
  SAME_SHAPE(x, y, "mlca_apply");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)x->n * x->h * x->w * (x->c / 8);
  const int oct_ = x->c / 8;
  const size_t tb_smem = (size_t)(x->h + x->w) * sizeof(int);
  if (oct_ > 0 && 256 % oct_ == 0 && total < ((int64_t)1 << 31) && x->h < 32768 && x->w < 32768 && local_size <= 1024 && tb_smem <= 48 * 1024 && total > 0 &&
      (!add || ((uintptr_t)add & 15) == 0)) {
    YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(mlca_apply2_kernel<T>, grid_for(total), 256, tb_smem, st, *x, att, local_size, (const T*)add, add_ld, *y);)
    YAD_LAUNCH_CHECK("mlca_apply");
    return 0;
  }
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(mlca_apply_kernel<T>, grid_for(total), TPB, 0, st, *x, att, local_size, (const T*)add, add_ld, *y);)
  YAD_LAUNCH_CHECK("mlca_apply");
  return 0;
}

int yad_gate_mlp(const float* g, const float* w1, const float* b1, const float* w2, const float* b2, int n, int c, int hidden,
                 int nout, int kind, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  YAD_LAUNCH(gate_mlp_kernel, n, 128, (hidden + nout) * sizeof(float), st, g, w1, b1, w2, b2, c, hidden, nout, kind, out);
  YAD_LAUNCH_CHECK("gate_mlp");
  return 0;
}

int yad_adt_apply(const yad_tensor* x, const float* imp, const float* alphas, const float* weight, const float* bias,
                  const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(x, "adt x");
  CHECK_VIEW(y, "adt y");
  SAME_SHAPE(x, y, "adt_apply");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)x->n * x->h * x->w * (x->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(adt_apply_kernel<T>, grid_for(total), TPB, 0, st, *x, imp, alphas, weight, bias, *y);)
  YAD_LAUNCH_CHECK("adt_apply");
  return 0;
}

int yad_eltwise_dev(int op, const yad_tensor* a, const void* b, int b_ld, const void* c3, int c3_ld, const void* d4, int d4_ld, float alpha,
                    float beta, float gamma, const float* alpha_dev, const float* beta_dev, const float* gamma_dev, const yad_tensor* y, int dtype,
                    void* stream) {
  CHECK_VIEW(a, "eltwise a");
  CHECK_VIEW(y, "eltwise y");
  SAME_SHAPE(a, y, "eltwise");
  YAD_CHECK(op >= 0 && op <= 8 && (b != nullptr || op == 5), "eltwise: bad op %d or null operand", op);
  YAD_CHECK(op != 3 || (c3 && d4), "eltwise: op 3 needs four operands");
  YAD_CHECK(op != 8 || c3, "eltwise: op 8 needs c3");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)a->n * a->h * a->w * (a->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(eltwise_kernel<T>, grid_for(total), TPB, 0, st, op, *a, (const T*)b, b_ld, (const T*)c3, c3_ld, (const T*)d4, d4_ld,
                                                                                alpha, beta, gamma, alpha_dev, beta_dev, gamma_dev, *y);)
  YAD_LAUNCH_CHECK("eltwise");
  return 0;
}

int yad_eltwise(int op, const yad_tensor* a, const void* b, int b_ld, const void* c3, int c3_ld, const void* d4, int d4_ld,
                float alpha, float beta, float gamma, const yad_tensor* y, int dtype, void* stream) {
  return yad_eltwise_dev(op, a, b, b_ld, c3, c3_ld, d4, d4_ld, alpha, beta, gamma, nullptr, nullptr, nullptr, y, dtype, stream);
}

int yad_group_mean(const yad_tensor* x, int s, const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(x, "group_mean x");
  CHECK_VIEW(y, "group_mean y");
  YAD_CHECK(x->n == y->n && x->c == y->c && (int64_t)x->h * x->w == (int64_t)s * y->h * y->w, "group_mean: shape mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)y->n * y->h * y->w * (y->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(group_mean_kernel<T>, grid_for(total), TPB, 0, st, *x, s, *y);)
  YAD_LAUNCH_CHECK("group_mean");
  return 0;
}

int yad_patch_filter(const yad_tensor* x, const float* m, float alpha, const void* add, int add_ld, const yad_tensor* y, int dtype,
                     void* stream) {
  CHECK_VIEW(x, "patch_filter x");
  CHECK_VIEW(y, "patch_filter y");
  SAME_SHAPE(x, y, "patch_filter");
  YAD_CHECK(x->h >= 2 && x->w >= 2, "patch_filter: reflect padding needs h, w >= 2");
  YAD_CHECK(((8 - x->h % 8) % 8) < x->h && ((8 - x->w % 8) % 8) < x->w, "patch_filter: reflect pad wider than the map (%dx%d)", x->h, x->w);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(((x->h + 7) / 8) * ((x->w + 7) / 8), x->n);
  static int pf_smem = -1;
  if (pf_smem < 0) { const char* ev = getenv("YAD_PATCH_FILTER_SMEM"); pf_smem = (ev && ev[0] == '0') ? 0 : 1; }
  const int64_t patches_total = (int64_t)grid.x * x->n;
  // matrix-resident kernel once there are enough patches to amortise the 128 KB fill of a CTA (the batched inference / training shapes)
  if (pf_smem && patches_total >= 64 && patches_total < ((int64_t)1 << 30) && (((uintptr_t)m) & 15) == 0 && x->c % 8 == 0) {
    const size_t smem = 64 * 64 * 8 * sizeof(float);
    const dim3 g2((unsigned)((patches_total + 63) / 64), x->c / 8);
    YAD_DISPATCH_DTYPE(dtype, {
      static bool attr = false;
      if (!attr) { cudaFuncSetAttribute(patch_filter_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
      YAD_LAUNCH(patch_filter_smem_kernel<T>, g2, 256, smem, st, *x, m, alpha, (const T*)add, add_ld, *y, (int)patches_total);
    })
    YAD_LAUNCH_CHECK("patch_filter");
    return 0;
  }
  int tpb = x->c < 128 ? ((x->c + 31) / 32) * 32 : 128;
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(patch_filter_kernel<T>, grid, tpb, 0, st, *x, m, alpha, (const T*)add, add_ld, *y);)
  YAD_LAUNCH_CHECK("patch_filter");
  return 0;
}

}  // extern "C"
