// SIMT implicit-GEMM convolution (fp32 accumulate) -- the precision-reference GPU path (fp32 build) and the path for
// shapes the tcgen05 kernel does not take.  NHWC activations, weights [cout][k*k][cin].
// GEMM view: M = n*ho*wo output pixels, N = cout, K = k*k*cin with K index = tap*cin + ci.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 32, NTHREADS = 256;

struct ConvGeom {
  int n, hi, wi, cin, ho, wo, cout;
  int kh, kw, stride, pad_h, pad_w, mode;
  int x_ld, y_ld, om_ld;
};

// Gather 8 consecutive input channels (ci..ci+7) of the input pixel that tap (ky,kx) of output pixel (img,oy,ox) reads.
template <typename T>
__device__ __forceinline__ void gather8(const T* __restrict__ x, const T* __restrict__ om, const ConvGeom& g, int img, int oy, int ox,
                                        int tap, int ci, float (&v)[8]) {
  const int ky = tap / g.kw, kx = tap - ky * g.kw;
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = 0.f;
  if (g.mode == YAD_CONV_NORMAL) {
    int iy = oy * g.stride - g.pad_h + ky, ix = ox * g.stride - g.pad_w + kx;
    if (iy >= 0 && iy < g.hi && ix >= 0 && ix < g.wi) load8(x + ((int64_t)(img * g.hi + iy) * g.wi + ix) * g.x_ld + ci, v);
  } else if (g.mode == YAD_CONV_TRANSPOSED) {
    // y[oy] += x[iy] * w[ky] with oy = iy*stride - pad + ky
    int ty = oy + g.pad_h - ky, tx = ox + g.pad_w - kx;
    if (ty >= 0 && tx >= 0 && (ty % g.stride) == 0 && (tx % g.stride) == 0) {
      int iy = ty / g.stride, ix = tx / g.stride;
      if (iy < g.hi && ix < g.wi) load8(x + ((int64_t)(img * g.hi + iy) * g.wi + ix) * g.x_ld + ci, v);
    }
  } else {  // modulated deformable 3x3, stride 1, pad 1
    const T* o = om + ((int64_t)(img * g.ho + oy) * g.wo + ox) * g.om_ld;
    float dy = ld1(o + 2 * tap), dx = ld1(o + 2 * tap + 1);
    float mk = sigmoidf_(ld1(o + 18 + tap));
    float py = (float)(oy - g.pad_h + ky) + dy, px = (float)(ox - g.pad_w + kx) + dx;
    if (py > -1.f && px > -1.f && py < (float)g.hi && px < (float)g.wi) {
      float fy = floorf(py), fx = floorf(px);
      int y0 = (int)fy, x0 = (int)fx;
      float ly = py - fy, lx = px - fx;
      float wgt[4] = {(1.f - ly) * (1.f - lx), (1.f - ly) * lx, ly * (1.f - lx), ly * lx};
#pragma unroll
      for (int c4 = 0; c4 < 4; c4++) {
        int yy = y0 + (c4 >> 1), xx = x0 + (c4 & 1);
        if (yy >= 0 && yy < g.hi && xx >= 0 && xx < g.wi) {
          float t[8];
          load8(x + ((int64_t)(img * g.hi + yy) * g.wi + xx) * g.x_ld + ci, t);
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] += wgt[c4] * t[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] *= mk;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(NTHREADS) conv_simt_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ om,
                                                             T* __restrict__ y, ConvGeom g, yad_epilogue e) {
  pdl_sync();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.n * g.ho * g.wo;
  const int K = g.kh * g.kw * g.cin;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // loader role: one 8-wide K chunk of one row per thread
  const int lrow = tid >> 2, lkc = tid & 3;
  int64_t lm = m0 + lrow;
  int l_img = 0, l_oy = 0, l_ox = 0;
  const bool l_valid = lm < M;
  if (l_valid) {
    l_img = (int)(lm / ((int64_t)g.ho * g.wo));
    int r = (int)(lm - (int64_t)l_img * g.ho * g.wo);
    l_oy = r / g.wo;
    l_ox = r - l_oy * g.wo;
  }
  const int ln = n0 + lrow;  // weight row

  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    const int kk = k0 + lkc * 8;
    float va[8], vb[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { va[i] = 0.f; vb[i] = 0.f; }
    if (kk < K) {
      if (l_valid) {
        int tap = kk / g.cin, ci = kk - tap * g.cin;
        gather8<T>(x, om, g, l_img, l_oy, l_ox, tap, ci, va);
      }
      if (ln < g.cout) load8(w + (int64_t)ln * K + kk, vb);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; i++) {
      As[lkc * 8 + i][lrow] = va[i];
      Bs[lkc * 8 + i][lrow] = vb[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  const int co = n0 + tx * 4;
  if (co >= g.cout) return;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int img = (int)(m / ((int64_t)g.ho * g.wo));
    float r[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    epilogue4<T>(r, e, img, m, co);
    store4(y + m * g.y_ld + co, r);
  }
}

}  // namespace

int yad_conv2d_simt(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, int dtype,
                    void* stream) {
  ConvGeom g;
  g.n = x->n; g.hi = x->h; g.wi = x->w; g.cin = x->c;
  g.ho = y->h; g.wo = y->w; g.cout = y->c;
  g.kh = d->kh; g.kw = d->kw; g.stride = d->stride; g.pad_h = d->pad_h; g.pad_w = d->pad_w; g.mode = d->mode;
  g.x_ld = x->ld; g.y_ld = y->ld; g.om_ld = d->offmask_ld;
  YAD_CHECK(x->n == y->n, "conv2d: batch mismatch %d vs %d", x->n, y->n);
  YAD_CHECK(g.cin % 8 == 0 && g.cout % 4 == 0, "conv2d: cin (%d) must be a multiple of 8 and cout (%d) of 4", g.cin, g.cout);
  YAD_CHECK(g.x_ld % 8 == 0 && g.y_ld % 4 == 0, "conv2d: ld must be aligned (x %d, y %d)", g.x_ld, g.y_ld);
  if (d->mode == YAD_CONV_NORMAL) {
    YAD_CHECK(g.ho == (g.hi + 2 * g.pad_h - g.kh) / g.stride + 1 && g.wo == (g.wi + 2 * g.pad_w - g.kw) / g.stride + 1,
              "conv2d: output shape %dx%d does not match input %dx%d k%dx%d s%d p%d,%d", g.ho, g.wo, g.hi, g.wi, g.kh, g.kw, g.stride, g.pad_h,
              g.pad_w);
  } else if (d->mode == YAD_CONV_TRANSPOSED) {
    YAD_CHECK(g.kh == 3 && g.kw == 3 && g.stride == 2 && g.pad_h == 1 && g.pad_w == 1 && g.ho == 2 * g.hi && g.wo == 2 * g.wi,
              "conv2d: transposed mode is k3 s2 p1 op1 only");
  } else if (d->mode == YAD_CONV_DEFORM) {
    YAD_CHECK(g.kh == 3 && g.kw == 3 && g.stride == 1 && g.pad_h == 1 && g.pad_w == 1 && g.ho == g.hi && g.wo == g.wi && d->offmask != nullptr,
              "conv2d: deformable mode is 3x3 s1 p1 with an offset/mask view");
  } else {
    YAD_CHECK(false, "conv2d: bad mode %d", d->mode);
  }
  int64_t M = (int64_t)g.n * g.ho * g.wo;
  dim3 grid(cdiv(M, BM), cdiv(g.cout, BN));
  cudaStream_t st = (cudaStream_t)stream;
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(conv_simt_kernel<T>, grid, NTHREADS, 0, st, (const T*)x->ptr, (const T*)w, (const T*)d->offmask, (T*)y->ptr, g, *e);)
  YAD_LAUNCH_CHECK("conv2d_simt");
  return 0;
}
