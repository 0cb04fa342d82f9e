// Backward kernels of the memory-bound blocks (training path): normalisation (GroupNorm and
// batch-statistics BatchNorm share one pair of kernels), activation, pooling / broadcast / gating blocks, MLCA, SPPF max-pools,
// AdaptiveDynamicTanh, tiny gate MLPs, the EDFFN patch filter and the head's loss gather / scatter.
// Conventions: NHWC views, 8-channel 128-bit vectors, fp32 arithmetic; parameter gradients are ACCUMULATED into fp32 buffers the
// caller zeroes once per step; activation gradients are written (acc = 0) or accumulated (acc = 1: dx += ...).
#include "common.cuh"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ int64_t pix_off(const yad_tensor& t, int n, int y, int x) { return ((int64_t)(n * t.h + y) * t.w + x) * t.ld; }
__device__ __forceinline__ int bin_start(int i, int in, int out) { return (int)(((int64_t)i * in) / out); }
__device__ __forceinline__ int bin_end(int i, int in, int out) { return (int)((((int64_t)(i + 1)) * in + out - 1) / out); }

template <typename T>
__device__ __forceinline__ void store8_acc(T* p, float (&v)[8], int acc) {
  if (acc) {
    float o[8];
    load8(p, o);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] += o[i];
  }
  store8(p, v);
}

// derivative of the activation with respect to its input u
__device__ __forceinline__ float act_grad(float u, int act) {
  switch (act) {
    case YAD_ACT_SILU: {
      float s = sigmoidf_(u);
      return s * (1.0f + u * (1.0f - s));
    }
    case YAD_ACT_RELU: return u > 0.f ? 1.0f : 0.0f;
    case YAD_ACT_SIGMOID: {
      float s = sigmoidf_(u);
      return s * (1.0f - s);
    }
    case YAD_ACT_GELU: return 0.5f * (1.0f + erff(u * 0.70710678118654752440f)) + u * 0.39894228040143267794f * expf(-0.5f * u * u);
    case YAD_ACT_HARDSWISH: return u < -3.0f ? 0.0f : (u > 3.0f ? 1.0f : (2.0f * u + 3.0f) * (1.0f / 6.0f));
    default: return 1.0f;
  }
}

// compile-time activation variant for the hot normalisation kernels: no per-element branch, ex2.approx / rcp.approx sigmoid (~2 ulp)
template <int ACT>
__device__ __forceinline__ float act_grad_ct(float u) {
  if constexpr (ACT == YAD_ACT_SILU) {
    const float s = sigmoid_fast(u);
    return s * fmaf(u, 1.0f - s, 1.0f);
  } else if constexpr (ACT == YAD_ACT_SIGMOID) {
    const float s = sigmoid_fast(u);
    return s * (1.0f - s);
  } else if constexpr (ACT == YAD_ACT_NONE) {
    return 1.0f;
  } else {
    return act_grad(u, ACT);
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
#define SAME_SHAPE(a, b, what) \
  YAD_CHECK((a)->n == (b)->n && (a)->h == (b)->h && (a)->w == (b)->w && (a)->c == (b)->c, "%s: shape mismatch", what)

// ------------------------------------------------------------------------------------------------------------------
// normalisation backward.  y = act(u), u = xhat * gamma + beta, xhat = (x - mean) * rstd over (pixels of image n) x (channels of group g).
// BatchNorm with batch statistics = the same with the whole batch viewed as one image and one channel per group.
// pass 1: g = dy * act'(u);  sums[n][grp] += (sum g*gamma, sum g*gamma*xhat);  dgamma[c] += sum g*xhat;  dbeta[c] += sum g
// pass 2: dx = rstd * (g*gamma - S1/cnt - xhat * S2/cnt)
// ------------------------------------------------------------------------------------------------------------------
template <typename T, int ACT>
__global__ void __launch_bounds__(TPB, 3) norm_bwd_reduce_kernel(yad_tensor x, yad_tensor dy, const double* __restrict__ stats, int groups, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, int act, double* __restrict__ sums,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_sync();
  extern __shared__ float sm[];  // mean[c], rstd[c], A[c], B[c]
  const int c = x.c, n = blockIdx.y, oct = c >> 3, cpg = c / groups;
  const int64_t hw = (int64_t)x.h * x.w;
  const double cnt = (double)hw * cpg;
  float* smean = sm;
  float* srstd = sm + c;
  float* sA = sm + 2 * c;
  float* sB = sm + 3 * c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const int g = ch / cpg;
    const double mean = stats[((int64_t)n * groups + g) * 2] / cnt;
    const double var = stats[((int64_t)n * groups + g) * 2 + 1] / cnt - mean * mean;
    smean[ch] = (float)mean;
    srstd[ch] = (float)(1.0 / sqrt(fmax(var, 0.0) + (double)eps));
    sA[ch] = 0.f;
    sB[ch] = 0.f;
  }
  __syncthreads();
  const int64_t per = (hw + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * per, p1 = min(hw, p0 + per);
  const T* xb = reinterpret_cast<const T*>(x.ptr) + (int64_t)n * hw * x.ld;
  const T* gb = reinterpret_cast<const T*>(dy.ptr) + (int64_t)n * hw * dy.ld;
  const int64_t items = (p1 - p0) * oct;
  // thread -> fixed octet when the block size allows it (register accumulation), generic otherwise
  const bool fixed = (blockDim.x % oct) == 0;
  const int step = fixed ? blockDim.x / oct : 1;
  if (fixed) {
    const int o = (threadIdx.x % oct) * 8, lane = threadIdx.x / oct;
    float a[8], b[8], c1[8], c2[8], ga[8], be[8];  // per-channel constants live in registers: xhat = v*c1 - c2, u = xhat*ga + be
#pragma unroll
    for (int i = 0; i < 8; i++) {
      a[i] = 0.f; b[i] = 0.f;
      c1[i] = srstd[o + i]; c2[i] = smean[o + i] * srstd[o + i]; ga[i] = gamma[o + i]; be[i] = beta[o + i];
    }
    int64_t p = p0 + lane;
    for (; p + 3 * step < p1; p += 4 * step) {  // four pixels per trip: eight 128-bit loads in flight (latency-bound otherwise)
      float v[4][8], g[4][8];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        load8(xb + (p + u * step) * x.ld + o, v[u]);
        load8(gb + (p + u * step) * dy.ld + o, g[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float xh = fmaf(v[u][i], c1[i], -c2[i]);
          const float gg = g[u][i] * act_grad_ct<ACT>(fmaf(xh, ga[i], be[i]));
          a[i] += gg;
          b[i] = fmaf(gg, xh, b[i]);
        }
      }
    }
    for (; p < p1; p += step) {
      float v[8], g[8];
      load8(xb + p * x.ld + o, v);
      load8(gb + p * dy.ld + o, g);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float xh = fmaf(v[i], c1[i], -c2[i]);
        const float gg = g[i] * act_grad_ct<ACT>(fmaf(xh, ga[i], be[i]));
        a[i] += gg;
        b[i] = fmaf(gg, xh, b[i]);
      }
    }
    // lanes of a warp that own the same octet (lane % oct equal) combine with shuffles first: 32 / oct times fewer shared-memory atomics
    if (oct <= 16 && (32 % oct) == 0) {
      for (int d = oct; d < 32; d <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          a[i] += __shfl_xor_sync(0xffffffffu, a[i], d);
          b[i] += __shfl_xor_sync(0xffffffffu, b[i], d);
        }
      }
      if ((threadIdx.x & 31) < oct) {
#pragma unroll
        for (int i = 0; i < 8; i++) { atomicAdd(&sA[o + i], a[i]); atomicAdd(&sB[o + i], b[i]); }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) { atomicAdd(&sA[o + i], a[i]); atomicAdd(&sB[o + i], b[i]); }
    }
  } else {
    for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
      const int64_t p = p0 + it / oct;
      const int o = (int)(it % oct) * 8;
      float v[8], g[8];
      load8(xb + p * x.ld + o, v);
      load8(gb + p * dy.ld + o, g);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float xh = (v[i] - smean[o + i]) * srstd[o + i];
        const float gg = g[i] * act_grad_ct<ACT>(fmaf(xh, gamma[o + i], beta[o + i]));
        atomicAdd(&sA[o + i], gg);
        atomicAdd(&sB[o + i], gg * xh);
      }
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[ch], sB[ch]);
    if (dbeta) atomicAdd(&dbeta[ch], sA[ch]);
  }
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < cpg; i++) {
      const int ch = g * cpg + i;
      s1 += (double)sA[ch] * gamma[ch];
      s2 += (double)sB[ch] * gamma[ch];
    }
    atomicAdd(&sums[((int64_t)n * groups + g) * 2], s1);
    atomicAdd(&sums[((int64_t)n * groups + g) * 2 + 1], s2);
  }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(TPB, 3) norm_bwd_apply_kernel(yad_tensor x, yad_tensor dy, const double* __restrict__ stats, const double* __restrict__ sums, int groups,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act, yad_tensor dx, int acc) {
  pdl_sync();
  extern __shared__ float sm[];  // mean[c], rstd[c], k1[c] = S1/cnt, k2[c] = S2/cnt
  const int c = x.c, n = blockIdx.y, oct = c >> 3, cpg = c / groups;
  const int64_t hw = (int64_t)x.h * x.w;
  const double cnt = (double)hw * cpg;
  float* smean = sm;
  float* srstd = sm + c;
  float* k1 = sm + 2 * c;
  float* k2 = sm + 3 * c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const int g = ch / cpg;
    const double mean = stats[((int64_t)n * groups + g) * 2] / cnt;
    const double var = stats[((int64_t)n * groups + g) * 2 + 1] / cnt - mean * mean;
    smean[ch] = (float)mean;
    srstd[ch] = (float)(1.0 / sqrt(fmax(var, 0.0) + (double)eps));
    k1[ch] = (float)(sums[((int64_t)n * groups + g) * 2] / cnt);
    k2[ch] = (float)(sums[((int64_t)n * groups + g) * 2 + 1] / cnt);
  }
  __syncthreads();
  const int64_t items = hw * oct;
  const T* xb = reinterpret_cast<const T*>(x.ptr) + (int64_t)n * hw * x.ld;
  const T* gb = reinterpret_cast<const T*>(dy.ptr) + (int64_t)n * hw * dy.ld;
  T* ob = reinterpret_cast<T*>(dx.ptr) + (int64_t)n * hw * dx.ld;
  if ((blockDim.x % oct) == 0) {
    // thread -> fixed octet; per-channel constants in registers: xhat = v*c1 - c2, u = xhat*ga + be, dx = gg*A - E - xhat*F
    const int o = (threadIdx.x % oct) * 8, lane = threadIdx.x / oct, ppb = blockDim.x / oct;
    float c1[8], c2[8], ga[8], be[8], A[8], E[8], F[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      c1[i] = srstd[o + i]; c2[i] = smean[o + i] * srstd[o + i]; ga[i] = gamma[o + i]; be[i] = beta[o + i];
      A[i] = srstd[o + i] * ga[i]; E[i] = srstd[o + i] * k1[o + i]; F[i] = srstd[o + i] * k2[o + i];
    }
    const int64_t stride = (int64_t)gridDim.x * ppb;
    int64_t p = (int64_t)blockIdx.x * ppb + lane;
    for (; p + stride < hw; p += 2 * stride) {  // two pixels per trip: four 128-bit loads in flight (the kernel is latency-bound otherwise)
      float v[8], g[8], v2[8], g2[8];
      load8(xb + p * x.ld + o, v);
      load8(gb + p * dy.ld + o, g);
      load8(xb + (p + stride) * x.ld + o, v2);
      load8(gb + (p + stride) * dy.ld + o, g2);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float xh = fmaf(v[i], c1[i], -c2[i]), xh2 = fmaf(v2[i], c1[i], -c2[i]);
        const float gg = g[i] * act_grad_ct<ACT>(fmaf(xh, ga[i], be[i]));
        const float gg2 = g2[i] * act_grad_ct<ACT>(fmaf(xh2, ga[i], be[i]));
        v[i] = fmaf(gg, A[i], -fmaf(xh, F[i], E[i]));
        v2[i] = fmaf(gg2, A[i], -fmaf(xh2, F[i], E[i]));
      }
      store8_acc(ob + p * dx.ld + o, v, acc);
      store8_acc(ob + (p + stride) * dx.ld + o, v2, acc);
    }
    for (; p < hw; p += stride) {
      float v[8], g[8];
      load8(xb + p * x.ld + o, v);
      load8(gb + p * dy.ld + o, g);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float xh = fmaf(v[i], c1[i], -c2[i]);
        const float gg = g[i] * act_grad_ct<ACT>(fmaf(xh, ga[i], be[i]));
        v[i] = fmaf(gg, A[i], -fmaf(xh, F[i], E[i]));
      }
      store8_acc(ob + p * dx.ld + o, v, acc);
    }
    return;
  }
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = it / oct;
    const int o = (int)(it - p * oct) * 8;
    float v[8], g[8];
    load8(xb + p * x.ld + o, v);
    load8(gb + p * dy.ld + o, g);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float xh = (v[i] - smean[o + i]) * srstd[o + i];
      const float gg = g[i] * act_grad_ct<ACT>(fmaf(xh, gamma[o + i], beta[o + i]));
      v[i] = srstd[o + i] * (gg * gamma[o + i] - k1[o + i] - xh * k2[o + i]);
    }
    store8_acc(ob + p * dx.ld + o, v, acc);
  }
}

// BatchNorm running statistics (nn.BatchNorm2d in train(): momentum update with the unbiased batch variance)
__global__ void bn_running_kernel(const double* __restrict__ stats, int c, double cnt, float momentum, float* __restrict__ rmean,
                                  float* __restrict__ rvar) {
  pdl_sync();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double mean = stats[ch * 2] / cnt;
  const double var = fmax(stats[ch * 2 + 1] / cnt - mean * mean, 0.0);
  const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
  rmean[ch] = (float)((1.0 - momentum) * rmean[ch] + momentum * mean);
  rvar[ch] = (float)((1.0 - momentum) * rvar[ch] + momentum * unbiased);
}

// dx = dy * f'(.) expressed with the OUTPUT y (sigmoid: y(1-y); relu: y > 0) -- conv epilogue activations keep only y
template <typename T>
__global__ void act_bwd_kernel(yad_tensor y, yad_tensor dy, int act, yad_tensor dx, int acc) {
  pdl_sync();
  const int oct = y.c >> 3;
  const int64_t total = (int64_t)y.n * y.h * y.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    float v[8], g[8];
    load8(reinterpret_cast<const T*>(y.ptr) + p * y.ld + o, v);
    load8(reinterpret_cast<const T*>(dy.ptr) + p * dy.ld + o, g);
#pragma unroll
    for (int i = 0; i < 8; i++) g[i] *= (act == YAD_ACT_SIGMOID) ? v[i] * (1.0f - v[i]) : (v[i] > 0.f ? 1.0f : 0.0f);
    store8_acc(reinterpret_cast<T*>(dx.ptr) + p * dx.ld + o, g, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// reductions: per-channel sums, dot products
// ------------------------------------------------------------------------------------------------------------------
// out[c] += sum over all pixels of a[p][c] (* b[p][c] when b != null).  Bias gradients, per-channel parameter gradients.
template <typename T>
__global__ void colsum_kernel(yad_tensor a, const T* __restrict__ b, int b_ld, float* __restrict__ out) {
  pdl_sync();
  extern __shared__ float sm[];
  const int c = a.c, oct = c >> 3;
  for (int i = threadIdx.x; i < c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int64_t npix = (int64_t)a.n * a.h * a.w;
  const int64_t per = (npix + gridDim.x - 1) / gridDim.x, p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  const int64_t items = (p1 - p0) * oct;
  if (blockDim.x % oct == 0) {
    const int o = (threadIdx.x % oct) * 8, lane = threadIdx.x / oct, step = blockDim.x / oct;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = 0.f;
    for (int64_t p = p0 + lane; p < p1; p += step) {
      float v[8];
      load8(reinterpret_cast<const T*>(a.ptr) + p * a.ld + o, v);
      if (b) {
        float w[8];
        load8(b + p * b_ld + o, w);
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] *= w[i];
      }
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) atomicAdd(&sm[o + i], s[i]);
  } else {
    for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
      const int64_t p = p0 + it / oct;
      const int o = (int)(it % oct) * 8;
      float v[8];
      load8(reinterpret_cast<const T*>(a.ptr) + p * a.ld + o, v);
      if (b) {
        float w[8];
        load8(b + p * b_ld + o, w);
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] *= w[i];
      }
#pragma unroll
      for (int i = 0; i < 8; i++) atomicAdd(&sm[o + i], v[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(&out[i], sm[i]);
}

// mode 0: out[0] += scale * sum_all a*b ; mode 1: out[n] += scale * inv[n]^-1... (see host) sum over image n of a*b
template <typename T>
__global__ void dot_kernel(yad_tensor a, const T* __restrict__ b, int b_ld, int per_image, float scale, const float* __restrict__ img_div,
                           float* __restrict__ out) {
  pdl_sync();
  __shared__ float red[32];
  const int oct = a.c >> 3, n = blockIdx.y;
  const int64_t hw = (int64_t)a.h * a.w;
  const int64_t npix = per_image ? hw : (int64_t)a.n * hw;
  const int64_t base = per_image ? (int64_t)n * hw : 0;
  const int64_t items = npix * oct;
  float s = 0.f;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = base + it / oct;
    const int o = (int)(it % oct) * 8;
    float v[8], w[8];
    load8(reinterpret_cast<const T*>(a.ptr) + p * a.ld + o, v);
    load8(b + p * b_ld + o, w);
#pragma unroll
    for (int i = 0; i < 8; i++) s = fmaf(v[i], w[i], s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    if (per_image && img_div) s /= img_div[n];
    atomicAdd(&out[per_image ? n : 0], s * scale);
  }
}

// y[p][0] = sum_c a[p][c] * b[p][c], y[p][1..7] = 0   (gradient of a per-pixel scalar gate; y is an 8-channel view)
template <typename T>
__global__ void dot_pixel_kernel(yad_tensor a, const T* __restrict__ b, int b_ld, yad_tensor y) {
  pdl_sync();
  const int64_t npix = (int64_t)a.n * a.h * a.w;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t p = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); p < npix; p += (int64_t)gridDim.x * wpb) {
    float s = 0.f;
    for (int o = lane * 8; o < a.c; o += 256) {
      float v[8], w[8];
      load8(reinterpret_cast<const T*>(a.ptr) + p * a.ld + o, v);
      load8(b + p * b_ld + o, w);
#pragma unroll
      for (int i = 0; i < 8; i++) s = fmaf(v[i], w[i], s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      float out[8] = {s, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      store8(reinterpret_cast<T*>(y.ptr) + p * y.ld, out);
    }
  }
}

// dx[n,y,x,:] (+)= img[n][:] * s_img + row[n,y,:] * s_row + col[n,x,:] * s_col   (backward of global / row / column means)
template <typename T>
__global__ void bcast_add_kernel(yad_tensor dx, const float* __restrict__ img, float s_img, const T* __restrict__ row, int row_ld, float s_row,
                                 const T* __restrict__ col, int col_ld, float s_col, int acc) {
  pdl_sync();
  const int oct = dx.c >> 3;
  const int64_t total = (int64_t)dx.n * dx.h * dx.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    const int px = (int)(p % dx.w), py = (int)((p / dx.w) % dx.h), n = (int)(p / ((int64_t)dx.w * dx.h));
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 0.f;
    if (img) {
      float t[8];
      load8(img + (int64_t)n * dx.c + o, t);
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = fmaf(t[i], s_img, v[i]);
    }
    if (row) {
      float t[8];
      load8(row + ((int64_t)n * dx.h + py) * row_ld + o, t);
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = fmaf(t[i], s_row, v[i]);
    }
    if (col) {
      float t[8];
      load8(col + ((int64_t)n * dx.w + px) * col_ld + o, t);
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = fmaf(t[i], s_col, v[i]);
    }
    store8_acc(reinterpret_cast<T*>(dx.ptr) + p * dx.ld + o, v, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// rowcol_gate backward: y = x * gh[row] * gw[col]  (x == null: y = gh * gw)
// grid (L, n, 2): z = 0 -> dgh[n][row] = sum_x dy * x * gw ; z = 1 -> dgw[n][col] = sum_y dy * x * gh
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void rowcol_gate_bwd_g_kernel(yad_tensor x, bool has_x, yad_tensor gh, yad_tensor gw, yad_tensor dy, yad_tensor dgh, yad_tensor dgw) {
  pdl_sync();
  extern __shared__ float sm[];
  const int n = blockIdx.y, c = dy.c, oct = c >> 3;
  const bool is_col = blockIdx.z == 1;
  const int L = is_col ? dy.w : dy.h;
  if ((int)blockIdx.x >= L) return;
  for (int i = threadIdx.x; i < c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int idx = blockIdx.x, count = is_col ? dy.h : dy.w;
  const int64_t items = (int64_t)count * oct;
  if ((blockDim.x % oct) == 0) {  // thread -> fixed octet: register accumulation, one shared-memory atomic per channel per thread
    const int o = (threadIdx.x % oct) * 8, step = blockDim.x / oct;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = 0.f;
    for (int j = threadIdx.x / oct; j < count; j += step) {
      const int py = is_col ? j : idx, px = is_col ? idx : j;
      float g[8], w[8];
      load8(reinterpret_cast<const T*>(dy.ptr) + pix_off(dy, n, py, px) + o, g);
      if (is_col)
        load8(reinterpret_cast<const T*>(gh.ptr) + ((int64_t)n * dy.h + py) * gh.ld + o, w);
      else
        load8(reinterpret_cast<const T*>(gw.ptr) + ((int64_t)n * dy.w + px) * gw.ld + o, w);
      if (has_x) {
        float v[8];
        load8(reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, py, px) + o, v);
#pragma unroll
        for (int i = 0; i < 8; i++) g[i] *= v[i];
      }
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = fmaf(g[i], w[i], s[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) atomicAdd(&sm[o + i], s[i]);
  } else
  for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
    const int j = (int)(it / oct), o = (int)(it % oct) * 8;
    const int py = is_col ? j : idx, px = is_col ? idx : j;
    float g[8], w[8];
    load8(reinterpret_cast<const T*>(dy.ptr) + pix_off(dy, n, py, px) + o, g);
    if (is_col)
      load8(reinterpret_cast<const T*>(gh.ptr) + ((int64_t)n * dy.h + py) * gh.ld + o, w);
    else
      load8(reinterpret_cast<const T*>(gw.ptr) + ((int64_t)n * dy.w + px) * gw.ld + o, w);
    if (has_x) {
      float v[8];
      load8(reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, py, px) + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) g[i] *= v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) atomicAdd(&sm[o + i], g[i] * w[i]);
  }
  __syncthreads();
  const yad_tensor& d = is_col ? dgw : dgh;
  T* dst = reinterpret_cast<T*>(d.ptr) + ((int64_t)n * L + idx) * d.ld;
  for (int i = threadIdx.x; i < c; i += blockDim.x) st1(dst + i, sm[i]);
}

template <typename T>
__global__ void rowcol_gate_bwd_x_kernel(yad_tensor gh, yad_tensor gw, yad_tensor dy, yad_tensor dx, int acc) {
  pdl_sync();
  const int oct = dy.c >> 3;
  const int64_t total = (int64_t)dy.n * dy.h * dy.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    const int px = (int)(p % dy.w), py = (int)((p / dy.w) % dy.h), n = (int)(p / ((int64_t)dy.w * dy.h));
    float a[8], b[8], g[8];
    load8(reinterpret_cast<const T*>(gh.ptr) + ((int64_t)n * dy.h + py) * gh.ld + o, a);
    load8(reinterpret_cast<const T*>(gw.ptr) + ((int64_t)n * dy.w + px) * gw.ld + o, b);
    load8(reinterpret_cast<const T*>(dy.ptr) + p * dy.ld + o, g);
#pragma unroll
    for (int i = 0; i < 8; i++) g[i] *= a[i] * b[i];
    store8_acc(reinterpret_cast<T*>(dx.ptr) + p * dx.ld + o, g, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// MLCA backward (block.py:1540-1584).  Forward: local = pool5x5(x); att = f(local); a[p] = mean of att over the bins that the
// adaptive "un-pool" assigns to pixel p; y = x * a (+ add).
// ------------------------------------------------------------------------------------------------------------------
// grid (25, n): datt[n][bin][c] = sum over pixels p whose un-pool range contains `bin` of dy[p] * x[p] / cnt(p)
template <typename T>
__global__ void mlca_bwd_datt_kernel(yad_tensor x, yad_tensor dy, int ls, float* __restrict__ datt) {
  pdl_sync();
  extern __shared__ float sm[];
  const int n = blockIdx.y, c = x.c, oct = c >> 3, bin = blockIdx.x, by = bin / ls, bx = bin % ls;
  for (int i = threadIdx.x; i < c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  // pixels py with bin_start(py, ls, h) <= by < bin_end(py, ls, h) form a contiguous range
  int y0 = x.h, y1 = 0, x0 = x.w, x1 = 0;
  for (int py = 0; py < x.h; py++)
    if (bin_start(py, ls, x.h) <= by && by < bin_end(py, ls, x.h)) { y0 = min(y0, py); y1 = max(y1, py + 1); }
  for (int px = 0; px < x.w; px++)
    if (bin_start(px, ls, x.w) <= bx && bx < bin_end(px, ls, x.w)) { x0 = min(x0, px); x1 = max(x1, px + 1); }
  const int bw = max(x1 - x0, 0), bh = max(y1 - y0, 0);
  const int64_t items = (int64_t)bw * bh * oct;
  for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int j = (int)(it / oct);
    const int py = y0 + j / bw, px = x0 + j % bw;
    const int cy = bin_end(py, ls, x.h) - bin_start(py, ls, x.h), cx = bin_end(px, ls, x.w) - bin_start(px, ls, x.w);
    const float inv = 1.0f / (float)(cy * cx);
    float v[8], g[8];
    load8(reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, py, px) + o, v);
    load8(reinterpret_cast<const T*>(dy.ptr) + pix_off(dy, n, py, px) + o, g);
#pragma unroll
    for (int i = 0; i < 8; i++) atomicAdd(&sm[o + i], v[i] * g[i] * inv);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) datt[((int64_t)n * ls * ls + bin) * c + i] = sm[i];
}

// Tiny part, phase A (grid n): local branch.  zl = conv1d_k(seq), att_l = sigmoid(zl); dzl = lw * datt * s(1-s);
// dwl[j] += sum dzl[i] * seq[i+j-r]; dlocal[n][q] = sum_j wl[j] dzl[q-j+r];  dG[by][c] += (1-lw) * sum_bx datt[n][by,bx][c]
__global__ void mlca_att_bwd_a_kernel(const float* __restrict__ local, const float* __restrict__ datt, const float* __restrict__ wl, int k, float lw,
                                      int c, int ls, float* __restrict__ dlocal, float* __restrict__ dwl, float* __restrict__ dG) {
  pdl_sync();
  extern __shared__ float sm[];  // seq[len], dzl[len], red[32]
  const int nb = ls * ls, len = nb * c, n = blockIdx.x, r = (k - 1) / 2;
  float* seq = sm;
  float* dzl = sm + len;
  float* red = dzl + len;
  for (int i = threadIdx.x; i < len; i += blockDim.x) seq[i] = local[(int64_t)n * len + i];
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < k; j++) {
      const int q = i + j - r;
      if (q >= 0 && q < len) s = fmaf(wl[j], seq[q], s);
    }
    const float sg = sigmoidf_(s);
    dzl[i] = lw * datt[(int64_t)n * len + i] * sg * (1.0f - sg);
  }
  __syncthreads();
  for (int j = 0; j < k; j++) {
    float s = 0.f;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      const int q = i + j - r;
      if (q >= 0 && q < len) s = fmaf(dzl[i], seq[q], s);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(&dwl[j], s);
  }
  for (int q = threadIdx.x; q < len; q += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < k; j++) {
      const int i = q - j + r;
      if (i >= 0 && i < len) s = fmaf(wl[j], dzl[i], s);
    }
    dlocal[(int64_t)n * len + q] = s;
  }
  for (int i = threadIdx.x; i < ls * c; i += blockDim.x) {
    const int by = i / c, ch = i % c;
    float s = 0.f;
    for (int bx = 0; bx < ls; bx++) s += datt[(int64_t)n * len + (by * ls + bx) * c + ch];
    atomicAdd(&dG[i], (1.0f - lw) * s);
  }
}

// phase B (grid = batch): global branch of image b.  G[by][c] = mean over b in R(by) of sigmoid(conv1d_k(glob_b)) where R(by) is the
// adaptive-pool range over the BATCH axis (reference behaviour, block.py:1575-1579).  dsg[b][c] = sum_{by: b in R(by)} dG[by][c]/|R(by)|
__global__ void mlca_att_bwd_b_kernel(const float* __restrict__ local, const float* __restrict__ dG, const float* __restrict__ wg, int k, int c, int ls,
                                      int batch, float* __restrict__ dlocal, float* __restrict__ dwg) {
  pdl_sync();
  extern __shared__ float sm[];  // glob[c], dzg[c], red[32]
  const int nb = ls * ls, len = nb * c, b = blockIdx.x, r = (k - 1) / 2;
  float* glob = sm;
  float* dzg = sm + c;
  float* red = dzg + c;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < nb; q++) s += local[(int64_t)b * len + q * c + ch];
    glob[ch] = s / (float)nb;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < k; j++) {
      const int q = ch + j - r;
      if (q >= 0 && q < c) s = fmaf(wg[j], glob[q], s);
    }
    const float sg = sigmoidf_(s);
    float d = 0.f;
    for (int by = 0; by < ls; by++) {
      const int b0 = bin_start(by, batch, ls), b1 = bin_end(by, batch, ls);
      if (b >= b0 && b < b1) d += dG[by * c + ch] / (float)(b1 - b0);
    }
    dzg[ch] = d * sg * (1.0f - sg);
  }
  __syncthreads();
  for (int j = 0; j < k; j++) {
    float s = 0.f;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const int q = ch + j - r;
      if (q >= 0 && q < c) s = fmaf(dzg[ch], glob[q], s);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(&dwg[j], s);
  }
  for (int q = threadIdx.x; q < c; q += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < k; j++) {
      const int ch = q - j + r;
      if (ch >= 0 && ch < c) s = fmaf(wg[j], dzg[ch], s);
    }
    s /= (float)nb;
    for (int bin = 0; bin < nb; bin++) dlocal[(int64_t)b * len + bin * c + q] += s;
  }
}

// dx (+)= dy * a[p] + sum over pool bins containing p of dlocal[bin] / cnt(bin)
template <typename T>
__global__ void mlca_bwd_apply_kernel(yad_tensor dy, const float* __restrict__ att, const float* __restrict__ dlocal, int ls, yad_tensor dx, int acc) {
  pdl_sync();
  const int oct = dy.c >> 3, c = dy.c;
  const int64_t total = (int64_t)dy.n * dy.h * dy.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    const int px = (int)(p % dy.w), py = (int)((p / dy.w) % dy.h), n = (int)(p / ((int64_t)dy.w * dy.h));
    const int y0 = bin_start(py, ls, dy.h), y1 = bin_end(py, ls, dy.h), x0 = bin_start(px, ls, dy.w), x1 = bin_end(px, ls, dy.w);
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
    float g[8];
    load8(reinterpret_cast<const T*>(dy.ptr) + p * dy.ld + o, g);
#pragma unroll
    for (int i = 0; i < 8; i++) g[i] *= a[i] * inv;
    for (int by = 0; by < ls; by++) {
      const int s0 = bin_start(by, dy.h, ls), s1 = bin_end(by, dy.h, ls);
      if (py < s0 || py >= s1) continue;
      for (int bx = 0; bx < ls; bx++) {
        const int t0 = bin_start(bx, dy.w, ls), t1 = bin_end(bx, dy.w, ls);
        if (px < t0 || px >= t1) continue;
        float t[8];
        load8(dlocal + ((int64_t)n * ls * ls + by * ls + bx) * c + o, t);
        const float iv = 1.0f / (float)((s1 - s0) * (t1 - t0));
#pragma unroll
        for (int i = 0; i < 8; i++) g[i] = fmaf(t[i], iv, g[i]);
      }
    }
    store8_acc(reinterpret_cast<T*>(dx.ptr) + p * dx.ld + o, g, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// 5x5 / s1 / p2 max-pool backward: the gradient of output pixel p goes to the FIRST maximum of its window in row-major order
// (torch max_pool2d index rule).  Output gradient = dy_t (activation dtype view, may be null) + dy_f (fp32 NHWC dense, may be null).
// dx_f: fp32 dense (n,h,w,c), accumulated with atomics (the caller zeroes it).
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool5_bwd_kernel(yad_tensor x, const T* __restrict__ dy_t, int dy_ld, const float* __restrict__ dy_f, float* __restrict__ dx_f) {
  pdl_sync();
  const int c = x.c;
  const int64_t total = (int64_t)x.n * x.h * x.w * c;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(it % c);
    const int64_t p = it / c;
    const int px = (int)(p % x.w), py = (int)((p / x.w) % x.h), n = (int)(p / ((int64_t)x.w * x.h));
    float g = 0.f;
    if (dy_t) g += ld1(dy_t + p * dy_ld + ch);
    if (dy_f) g += dy_f[p * c + ch];
    if (g == 0.f) continue;
    float best = -INFINITY;
    int by = -1, bx = -1;
    for (int yy = max(py - 2, 0); yy <= min(py + 2, x.h - 1); yy++)
      for (int xx = max(px - 2, 0); xx <= min(px + 2, x.w - 1); xx++) {
        const float v = ld1(reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, yy, xx) + ch);
        if (v > best || by < 0) { best = v; by = yy; bx = xx; }
      }
    atomicAdd(&dx_f[(((int64_t)n * x.h + by) * x.w + bx) * c + ch], g);
  }
}

// y (+)= scale * x_f32 (dense fp32 NHWC with c channels) -> activation dtype view
template <typename T>
__global__ void cast_acc_kernel(const float* __restrict__ src, float scale, yad_tensor y, int acc) {
  pdl_sync();
  const int oct = y.c >> 3;
  const int64_t total = (int64_t)y.n * y.h * y.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    float v[8];
    load8(src + p * y.c + o, v);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] *= scale;
    store8_acc(reinterpret_cast<T*>(y.ptr) + p * y.ld + o, v, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// pool_upsample backward (adaptive avg pool to (h/s, w/s) then bilinear upsample): scatter into the pooled grid, then gather
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pool_upsample_bwd_scatter_kernel(yad_tensor dy, int s, float* __restrict__ dpool) {
  pdl_sync();
  const int c = dy.c, oct = c >> 3;
  const int hp = dy.h / s, wp = dy.w / s;
  const float sy = (float)hp / (float)dy.h, sx = (float)wp / (float)dy.w;
  const int64_t total = (int64_t)dy.n * dy.h * dy.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    const int px = (int)(p % dy.w), py = (int)((p / dy.w) % dy.h), n = (int)(p / ((int64_t)dy.w * dy.h));
    const float fy = fmaxf(sy * ((float)py + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)px + 0.5f) - 0.5f, 0.f);
    const int iy0 = (int)fy, ix0 = (int)fx;
    const int iy1 = min(iy0 + 1, hp - 1), ix1 = min(ix0 + 1, wp - 1);
    const float ly = fy - (float)iy0, lx = fx - (float)ix0;
    float g[8];
    load8(reinterpret_cast<const T*>(dy.ptr) + p * dy.ld + o, g);
    for (int q = 0; q < 4; q++) {
      const int by = (q >> 1) ? iy1 : iy0, bx = (q & 1) ? ix1 : ix0;
      const float wgt = ((q >> 1) ? ly : 1.f - ly) * ((q & 1) ? lx : 1.f - lx);
      if (wgt == 0.f) continue;
      float* d = dpool + (((int64_t)n * hp + by) * wp + bx) * c + o;
#pragma unroll
      for (int i = 0; i < 8; i++) atomicAdd(d + i, g[i] * wgt);
    }
  }
}

template <typename T>
__global__ void pool_upsample_bwd_gather_kernel(const float* __restrict__ dpool, int s, yad_tensor dx, int acc) {
  pdl_sync();
  const int c = dx.c, oct = c >> 3;
  const int hp = dx.h / s, wp = dx.w / s;
  const int64_t total = (int64_t)dx.n * dx.h * dx.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    const int px = (int)(p % dx.w), py = (int)((p / dx.w) % dx.h), n = (int)(p / ((int64_t)dx.w * dx.h));
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; i++) g[i] = 0.f;
    // adaptive bins may overlap when h is not a multiple of hp: test the (few) candidate bins around py*hp/h
    for (int by = max(0, py * hp / dx.h - 1); by < min(hp, py * hp / dx.h + 2); by++) {
      const int y0 = bin_start(by, dx.h, hp), y1 = bin_end(by, dx.h, hp);
      if (py < y0 || py >= y1) continue;
      for (int bx = max(0, px * wp / dx.w - 1); bx < min(wp, px * wp / dx.w + 2); bx++) {
        const int x0 = bin_start(bx, dx.w, wp), x1 = bin_end(bx, dx.w, wp);
        if (px < x0 || px >= x1) continue;
        float t[8];
        load8(dpool + (((int64_t)n * hp + by) * wp + bx) * c + o, t);
        const float iv = 1.0f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
        for (int i = 0; i < 8; i++) g[i] = fmaf(t[i], iv, g[i]);
      }
    }
    store8_acc(reinterpret_cast<T*>(dx.ptr) + p * dx.ld + o, g, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// tiny gate MLP backward (one block per image): out = f(w2 relu(w1 g + b1) + b2), f = sigmoid (kind 0) / softmax (kind 1)
// ------------------------------------------------------------------------------------------------------------------
__global__ void gate_mlp_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w1, const float* __restrict__ b1,
                                    const float* __restrict__ w2, const float* __restrict__ b2, int c, int hidden, int nout, int kind,
                                    const float* __restrict__ dout, float* __restrict__ dg, float* __restrict__ dw1, float* __restrict__ db1,
                                    float* __restrict__ dw2, float* __restrict__ db2) {
  pdl_sync();
  extern __shared__ float sm[];  // hid[hidden], o[nout], dz[nout], dhid[hidden]
  float* hid = sm;
  float* o = sm + hidden;
  float* dz = o + nout;
  float* dhid = dz + nout;
  const int n = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int h = wid; h < hidden; h += nw) {
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s = fmaf(w1[h * c + i], g[(int64_t)n * c + i], s);
    s = warp_sum(s);
    if (lane == 0) hid[h] = s + b1[h];  // pre-activation
  }
  __syncthreads();
  for (int j = wid; j < nout; j += nw) {
    float s = 0.f;
    for (int i = lane; i < hidden; i += 32) s = fmaf(w2[j * hidden + i], fmaxf(hid[i], 0.f), s);
    s = warp_sum(s);
    if (lane == 0) o[j] = s + b2[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (kind == 0) {
      for (int j = 0; j < nout; j++) {
        const float sg = sigmoidf_(o[j]);
        dz[j] = dout[(int64_t)n * nout + j] * sg * (1.0f - sg);
      }
    } else {
      float m = -INFINITY, sum = 0.f, dotp = 0.f;
      for (int j = 0; j < nout; j++) m = fmaxf(m, o[j]);
      for (int j = 0; j < nout; j++) sum += expf(o[j] - m);
      for (int j = 0; j < nout; j++) dotp += expf(o[j] - m) / sum * dout[(int64_t)n * nout + j];
      for (int j = 0; j < nout; j++) dz[j] = expf(o[j] - m) / sum * (dout[(int64_t)n * nout + j] - dotp);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < nout; j++) {
      s = fmaf(w2[j * hidden + i], dz[j], s);
      atomicAdd(&dw2[j * hidden + i], dz[j] * fmaxf(hid[i], 0.f));
    }
    dhid[i] = hid[i] > 0.f ? s : 0.f;
    atomicAdd(&db1[i], dhid[i]);
  }
  for (int j = threadIdx.x; j < nout; j += blockDim.x) atomicAdd(&db2[j], dz[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    float s = 0.f;
    const float gi = g[(int64_t)n * c + i];
    for (int h = 0; h < hidden; h++) {
      s = fmaf(w1[h * c + i], dhid[h], s);
      atomicAdd(&dw1[h * c + i], dhid[h] * gi);
    }
    dg[(int64_t)n * c + i] = s;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// AdaptiveDynamicTanh backward: y = w_c * t + b_c, t = sum_i imp[n][i] tanh(alpha_i x).  grid (chunks, n)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void adt_bwd_kernel(yad_tensor x, yad_tensor dy, const float* __restrict__ imp, const float* __restrict__ alphas,
                               const float* __restrict__ weight, yad_tensor dx, int acc, float* __restrict__ dimp, float* __restrict__ dalpha,
                               float* __restrict__ dweight, float* __restrict__ dbias) {
  pdl_sync();
  extern __shared__ float sm[];  // dw[c], db[c], red[32]
  __shared__ float red[32];
  const int c = x.c, oct = c >> 3, n = blockIdx.y;
  float* sdw = sm;
  float* sdb = sm + c;
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int64_t hw = (int64_t)x.h * x.w;
  const int64_t per = (hw + gridDim.x - 1) / gridDim.x, p0 = blockIdx.x * per, p1 = min(hw, p0 + per);
  const int64_t items = (p1 - p0) * oct;
  const float a0 = alphas[0], a1 = alphas[1], a2 = alphas[2];
  const float i0 = imp[n * 3], i1 = imp[n * 3 + 1], i2 = imp[n * 3 + 2];
  float di[3] = {0.f, 0.f, 0.f}, da[3] = {0.f, 0.f, 0.f};
  for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
    const int64_t p = (int64_t)n * hw + p0 + it / oct;
    const int o = (int)(it % oct) * 8;
    float v[8], g[8], wv[8], out[8];
    load8(reinterpret_cast<const T*>(x.ptr) + p * x.ld + o, v);
    load8(reinterpret_cast<const T*>(dy.ptr) + p * dy.ld + o, g);
    load8(weight + o, wv);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float t0 = tanhf(a0 * v[i]), t1 = tanhf(a1 * v[i]), t2 = tanhf(a2 * v[i]);
      const float t = t0 * i0 + t1 * i1 + t2 * i2;
      const float dt = g[i] * wv[i];
      atomicAdd(&sdw[o + i], g[i] * t);
      atomicAdd(&sdb[o + i], g[i]);
      const float s0 = 1.f - t0 * t0, s1 = 1.f - t1 * t1, s2 = 1.f - t2 * t2;
      out[i] = dt * (i0 * a0 * s0 + i1 * a1 * s1 + i2 * a2 * s2);
      di[0] = fmaf(dt, t0, di[0]); di[1] = fmaf(dt, t1, di[1]); di[2] = fmaf(dt, t2, di[2]);
      da[0] = fmaf(dt * i0 * v[i], s0, da[0]); da[1] = fmaf(dt * i1 * v[i], s1, da[1]); da[2] = fmaf(dt * i2 * v[i], s2, da[2]);
    }
    store8_acc(reinterpret_cast<T*>(dx.ptr) + p * dx.ld + o, out, acc);
  }
  for (int i = 0; i < 3; i++) {
    const float s = block_sum(di[i], red);
    const float t = block_sum(da[i], red);
    if (threadIdx.x == 0) { atomicAdd(&dimp[n * 3 + i], s); atomicAdd(&dalpha[i], t); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) { atomicAdd(&dweight[i], sdw[i]); atomicAdd(&dbias[i], sdb[i]); }
}

// ------------------------------------------------------------------------------------------------------------------
// elementwise helpers of the backward pass
// ------------------------------------------------------------------------------------------------------------------
// gelu gate (EDFFN): forward y = gelu(a) * b is yad_eltwise op 7; backward: da (+)= dy * b * gelu'(a), db (+)= dy * gelu(a)
template <typename T>
__global__ void gelu_gate_bwd_kernel(yad_tensor a, const T* __restrict__ b, int b_ld, yad_tensor dy, yad_tensor da, const T* db_c, int db_ld, int acc) {
  pdl_sync();
  T* db = const_cast<T*>(db_c);
  const int oct = a.c >> 3;
  const int64_t total = (int64_t)a.n * a.h * a.w * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    float va[8], vb[8], g[8], ga[8], gb[8];
    load8(reinterpret_cast<const T*>(a.ptr) + p * a.ld + o, va);
    load8(b + p * b_ld + o, vb);
    load8(reinterpret_cast<const T*>(dy.ptr) + p * dy.ld + o, g);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      ga[i] = g[i] * vb[i] * act_grad(va[i], YAD_ACT_GELU);
      gb[i] = g[i] * apply_act(va[i], YAD_ACT_GELU);
    }
    store8_acc(reinterpret_cast<T*>(da.ptr) + p * da.ld + o, ga, acc);
    store8_acc(db + p * db_ld + o, gb, acc);
  }
}

// y (+)= a * s[n]  (per-image fp32 scale)
template <typename T>
__global__ void scale_img_kernel(yad_tensor a, const float* __restrict__ s, yad_tensor y, int acc) {
  pdl_sync();
  const int oct = a.c >> 3;
  const int64_t hw = (int64_t)a.h * a.w, total = (int64_t)a.n * hw * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    const float sc = s[p / hw];
    float v[8];
    load8(reinterpret_cast<const T*>(a.ptr) + p * a.ld + o, v);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] *= sc;
    store8_acc(reinterpret_cast<T*>(y.ptr) + p * y.ld + o, v, acc);
  }
}

// dx[(n*s + g)*T + t] (+)= dy[n*T + t] / s
template <typename T>
__global__ void group_mean_bwd_kernel(yad_tensor dy, int s, yad_tensor dx, int acc) {
  pdl_sync();
  const int oct = dy.c >> 3, Tn = dy.w * dy.h;
  const int64_t total = (int64_t)dx.n * dx.h * dx.w * oct;
  const float inv = 1.0f / (float)s;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;  // token index in dx: (n*s + g)*T + t
    const int64_t t = p % Tn, n = p / ((int64_t)Tn * s);
    float v[8];
    load8(reinterpret_cast<const T*>(dy.ptr) + (n * Tn + t) * dy.ld + o, v);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] *= inv;
    store8_acc(reinterpret_cast<T*>(dx.ptr) + p * dx.ld + o, v, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// EDFFN patch filter backward.  Forward: out_patch = M_c in_patch, in = reflect-padded x, out cropped.
// dx: dx_f (fp32 dense (n,h,w,c), atomics; reflected sources may belong to another patch).  grid (patches, n), block = c threads
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void patch_filter_bwd_x_kernel(yad_tensor dy, const float* __restrict__ m, float alpha, float* __restrict__ dx_f) {
  pdl_sync();
  const int c = dy.c, n = blockIdx.y;
  const int wp = (dy.w + 7) / 8;
  const int pr = blockIdx.x / wp, pc = blockIdx.x % wp;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float g[64];
#pragma unroll
    for (int o = 0; o < 64; o++) {
      const int yy = pr * 8 + (o >> 3), xx = pc * 8 + (o & 7);
      g[o] = (yy < dy.h && xx < dy.w) ? alpha * ld1(reinterpret_cast<const T*>(dy.ptr) + pix_off(dy, n, yy, xx) + ch) : 0.f;
    }
    for (int i = 0; i < 64; i++) {
      float s = 0.f;
#pragma unroll
      for (int o = 0; o < 64; o++) s = fmaf(m[((int64_t)o * 64 + i) * c + ch], g[o], s);
      int yy = pr * 8 + (i >> 3), xx = pc * 8 + (i & 7);
      if (yy >= dy.h) yy = 2 * dy.h - 2 - yy;
      if (xx >= dy.w) xx = 2 * dy.w - 2 - xx;
      atomicAdd(&dx_f[(((int64_t)n * dy.h + yy) * dy.w + xx) * c + ch], s);
    }
  }
}

// dM[o][i][ch] += alpha * sum over (n, patch) of dy[o] * in[i].  grid (c), block 256: thread t owns (o, i) pairs t, t+256, ...
template <typename T>
__global__ void patch_filter_bwd_m_kernel(yad_tensor x, yad_tensor dy, float alpha, float* __restrict__ dm) {
  pdl_sync();
  __shared__ float sin[64], sg[64];
  const int ch = blockIdx.x, c = x.c;
  const int hp = (x.h + 7) / 8, wp = (x.w + 7) / 8;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = 0.f;
  const int samples = x.n * hp * wp;
  for (int s = 0; s < samples; s++) {
    const int n = s / (hp * wp), pr = (s / wp) % hp, pc = s % wp;
    __syncthreads();
    if (threadIdx.x < 64) {
      int yy = pr * 8 + (threadIdx.x >> 3), xx = pc * 8 + (threadIdx.x & 7);
      if (yy >= x.h) yy = 2 * x.h - 2 - yy;
      if (xx >= x.w) xx = 2 * x.w - 2 - xx;
      sin[threadIdx.x] = ld1(reinterpret_cast<const T*>(x.ptr) + pix_off(x, n, yy, xx) + ch);
    } else if (threadIdx.x < 128) {
      const int o = threadIdx.x - 64;
      const int yy = pr * 8 + (o >> 3), xx = pc * 8 + (o & 7);
      sg[o] = (yy < x.h && xx < x.w) ? ld1(reinterpret_cast<const T*>(dy.ptr) + pix_off(dy, n, yy, xx) + ch) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int idx = threadIdx.x + k * 256;  // = o*64 + i
      acc[k] = fmaf(sg[idx >> 6], sin[idx & 63], acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int idx = threadIdx.x + k * 256;
    atomicAdd(&dm[(int64_t)idx * c + ch], alpha * acc[k]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// head outputs <-> loss layout.  levels: NHWC views (n, h_l, w_l, 4*reg_max + nc); distri fp32 (B, N, 4*reg_max), logits fp32 (B, N, nc)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void head_pack_kernel(yad_tensor lv, int a0, int N, int nd, int nc, float* __restrict__ distri, float* __restrict__ logits) {
  pdl_sync();
  const int oct = lv.c >> 3, hw = lv.h * lv.w;
  const int64_t total = (int64_t)lv.n * hw * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    const int64_t b = p / hw, a = a0 + p % hw;
    float v[8];
    load8(reinterpret_cast<const T*>(lv.ptr) + p * lv.ld + o, v);
    if (o < nd) {
      store8(distri + (b * N + a) * nd + o, v);
    } else if ((nc & 7) == 0) {
      store8(logits + (b * N + a) * nc + (o - nd), v);
    } else {  // class count not a multiple of 8 (custom datasets): the level carries pad8(nc) class channels, the loss sees the first nc
      float* dst = logits + (b * N + a) * nc;
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (o - nd + i < nc) dst[o - nd + i] = v[i];
    }
  }
}

template <typename T>
__global__ void head_unpack_kernel(const float* __restrict__ gd, const float* __restrict__ gl, float scale, int a0, int N, int nd, int nc,
                                   yad_tensor lv) {
  pdl_sync();
  const int oct = lv.c >> 3, hw = lv.h * lv.w;
  const int64_t total = (int64_t)lv.n * hw * oct;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(it % oct) * 8;
    const int64_t p = it / oct;
    const int64_t b = p / hw, a = a0 + p % hw;
    float v[8];
    if (o < nd) {
      load8(gd + (b * N + a) * nd + o, v);
    } else if ((nc & 7) == 0) {
      load8(gl + (b * N + a) * nc + (o - nd), v);
    } else {  // padding class channels get a zero gradient
      const float* src = gl + (b * N + a) * nc;
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = (o - nd + i < nc) ? src[o - nd + i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] *= scale;
    store8(reinterpret_cast<T*>(lv.ptr) + p * lv.ld + o, v);
  }
}

// Fusion('bifpn') weights (block.py:1532-1534): w = relu(p) / (sum relu(p) + 1e-4).  Single thread; k <= 8.
__global__ void fusion_weights_kernel(const float* __restrict__ p, int k, float* __restrict__ w, const float* __restrict__ dw, float* __restrict__ dp) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float r[8], s = 1e-4f;
  for (int i = 0; i < k; i++) { r[i] = fmaxf(p[i], 0.f); s += r[i]; }
  if (w) for (int i = 0; i < k; i++) w[i] = r[i] / s;
  if (dp) {
    float dotp = 0.f;
    for (int i = 0; i < k; i++) dotp += dw[i] * r[i] / s;
    for (int i = 0; i < k; i++) dp[i] += p[i] > 0.f ? (dw[i] - dotp) / s : 0.f;
  }
}

// C[m][n] (+)= sum_k A[m][k] * B[n][k] (trans_a = 0)   or   sum_k A[k][m] * B[k][n] (trans_a = 1); tiny parameter-sized matrices only.
// trans_a = 0: thread per output, n fastest (B rows are short: K <= a few dozen).  trans_a = 1 (long K): grid (n, K chunks), threads over m so
// that A[k][m] is read coalesced; partial sums meet in C through atomics (the caller passes acc = 1 semantics: C is accumulated).
__global__ void small_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K, int acc) {
  pdl_sync();
  const int64_t total = (int64_t)M * N;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(it / N), n = (int)(it % N);
    float s = 0.f;
    for (int k = 0; k < K; k++) s = fmaf(A[(int64_t)m * K + k], B[(int64_t)n * K + k], s);
    C[it] = acc ? C[it] + s : s;
  }
}
__global__ void small_gemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K, int chunk) {
  pdl_sync();
  const int n = blockIdx.x, k0 = blockIdx.y * chunk, k1 = min(K, k0 + chunk);
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    float s = 0.f;
    for (int k = k0; k < k1; k++) s = fmaf(A[(int64_t)k * M + m], B[(int64_t)k * N + n], s);
    atomicAdd(&C[(int64_t)m * N + n], s);
  }
}

}  // namespace

extern "C" {

int yad_norm_bwd(const yad_tensor* x, const yad_tensor* dy, const double* stats, int groups, const float* gamma, const float* beta, float eps,
                 int act, double* sums, float* dgamma, float* dbeta, const yad_tensor* dx, int acc, int dtype, void* stream) {
  CHECK_VIEW(x, "norm_bwd x");
  CHECK_VIEW(dy, "norm_bwd dy");
  CHECK_VIEW(dx, "norm_bwd dx");
  SAME_SHAPE(x, dy, "norm_bwd");
  SAME_SHAPE(x, dx, "norm_bwd");
  YAD_CHECK(groups > 0 && x->c % groups == 0, "norm_bwd: %d channels not divisible into %d groups", x->c, groups);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * groups * x->n, st);
  const int64_t hw = (int64_t)x->h * x->w;
  // grids are sized for the whole launch (about 8 CTAs per SM in total), not per image: the per-CTA prologue (statistics -> mean / rstd per
  // channel) and epilogue (shared -> global atomics) are amortised over >= 16 pixel-octets per thread
  int chunks = (int)((hw * (x->c / 8) + TPB * 16 - 1) / (TPB * 16));
  const int cap1 = 1184 / x->n > 1 ? 1184 / x->n : 1;
  chunks = chunks < 1 ? 1 : (chunks > cap1 ? cap1 : chunks);
  const size_t smem = 4 * x->c * sizeof(float);
  YAD_CHECK(smem <= 48 * 1024, "norm_bwd: too many channels (%d)", x->c);
  dim3 g1(chunks, x->n);
  int gx = (int)((hw * (x->c / 8) + TPB * 8 - 1) / (TPB * 8));
  const int cap2 = 2368 / x->n > 1 ? 2368 / x->n : 1;
  gx = gx < 1 ? 1 : (gx > cap2 ? cap2 : gx);
  dim3 g2(gx, x->n);
#define NORM_BWD_LAUNCH(A)                                                                                                          \
  case A:                                                                                                                           \
    YAD_LAUNCH((norm_bwd_reduce_kernel<T, A>), g1, TPB, smem, st, *x, *dy, stats, groups, gamma, beta, eps, act, sums, dgamma, dbeta);         \
    YAD_LAUNCH((norm_bwd_apply_kernel<T, A>), g2, TPB, smem, st, *x, *dy, stats, sums, groups, gamma, beta, eps, act, *dx, acc);               \
    break;
  YAD_DISPATCH_DTYPE(dtype, {
    switch (act) {
      NORM_BWD_LAUNCH(YAD_ACT_NONE)
      NORM_BWD_LAUNCH(YAD_ACT_SILU)
      NORM_BWD_LAUNCH(YAD_ACT_RELU)
      NORM_BWD_LAUNCH(YAD_ACT_SIGMOID)
      NORM_BWD_LAUNCH(YAD_ACT_GELU)
      NORM_BWD_LAUNCH(YAD_ACT_HARDSWISH)
      default: yad_set_error("norm_bwd: unknown activation %d", act); return 1;
    }
  })
#undef NORM_BWD_LAUNCH
  YAD_LAUNCH_CHECK("norm_bwd");
  return 0;
}

int yad_bn_running_update(const double* stats, int c, double count, float momentum, float* running_mean, float* running_var, void* stream) {
  YAD_LAUNCH(bn_running_kernel, cdiv(c, 128), 128, 0, (cudaStream_t)stream, stats, c, count, momentum, running_mean, running_var);
  YAD_LAUNCH_CHECK("bn_running_update");
  return 0;
}

int yad_act_bwd(const yad_tensor* y, const yad_tensor* dy, int act, const yad_tensor* dx, int acc, int dtype, void* stream) {
  CHECK_VIEW(y, "act_bwd y");
  CHECK_VIEW(dy, "act_bwd dy");
  CHECK_VIEW(dx, "act_bwd dx");
  SAME_SHAPE(y, dy, "act_bwd");
  SAME_SHAPE(y, dx, "act_bwd");
  YAD_CHECK(act == YAD_ACT_SIGMOID || act == YAD_ACT_RELU, "act_bwd: only sigmoid / relu are expressible from the output (act %d)", act);
  const int64_t total = (int64_t)y->n * y->h * y->w * (y->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(act_bwd_kernel<T>, grid_for(total), TPB, 0, (cudaStream_t)stream, *y, *dy, act, *dx, acc);)
  YAD_LAUNCH_CHECK("act_bwd");
  return 0;
}

int yad_colsum(const yad_tensor* a, const void* b, int b_ld, float* out, int dtype, void* stream) {
  CHECK_VIEW(a, "colsum");
  const int64_t npix = (int64_t)a->n * a->h * a->w;
  int blocks = (int)((npix * (a->c / 8) + TPB * 8 - 1) / (TPB * 8));
  blocks = blocks < 1 ? 1 : (blocks > 592 ? 592 : blocks);
  const int oct = a->c / 8;
  const int tpb = oct <= TPB ? oct * (TPB / oct) : TPB;  // a multiple of the octet count: every thread keeps one octet in registers
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(colsum_kernel<T>, blocks, tpb, a->c * sizeof(float), (cudaStream_t)stream, *a, (const T*)b, b_ld, out);)
  YAD_LAUNCH_CHECK("colsum");
  return 0;
}

int yad_dot(const yad_tensor* a, const void* b, int b_ld, int per_image, float scale, const float* img_div, float* out, int dtype, void* stream) {
  CHECK_VIEW(a, "dot");
  const int64_t items = (int64_t)(per_image ? 1 : a->n) * a->h * a->w * (a->c / 8);
  int gx = (int)((items + TPB * 8 - 1) / (TPB * 8));
  gx = gx < 1 ? 1 : (gx > 296 ? 296 : gx);
  dim3 grid(gx, per_image ? a->n : 1);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(dot_kernel<T>, grid, TPB, 0, (cudaStream_t)stream, *a, (const T*)b, b_ld, per_image, scale, img_div, out);)
  YAD_LAUNCH_CHECK("dot");
  return 0;
}

int yad_dot_pixel(const yad_tensor* a, const void* b, int b_ld, const yad_tensor* y, int dtype, void* stream) {
  CHECK_VIEW(a, "dot_pixel a");
  CHECK_VIEW(y, "dot_pixel y");
  YAD_CHECK(y->c == 8 && y->n == a->n && y->h == a->h && y->w == a->w, "dot_pixel: y must be an 8-channel view of the same pixels");
  const int64_t npix = (int64_t)a->n * a->h * a->w;
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(dot_pixel_kernel<T>, grid_for(npix, 8), TPB, 0, (cudaStream_t)stream, *a, (const T*)b, b_ld, *y);)
  YAD_LAUNCH_CHECK("dot_pixel");
  return 0;
}

int yad_bcast_add(const yad_tensor* dx, const float* img, float s_img, const yad_tensor* row, float s_row, const yad_tensor* col, float s_col,
                  int acc, int dtype, void* stream) {
  CHECK_VIEW(dx, "bcast_add");
  const int64_t total = (int64_t)dx->n * dx->h * dx->w * (dx->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(bcast_add_kernel<T>, grid_for(total), TPB, 0, (cudaStream_t)stream, 
      *dx, img, s_img, row ? (const T*)row->ptr : nullptr, row ? row->ld : 0, s_row, col ? (const T*)col->ptr : nullptr, col ? col->ld : 0, s_col, acc);)
  YAD_LAUNCH_CHECK("bcast_add");
  return 0;
}

int yad_rowcol_gate_bwd(const yad_tensor* x, const yad_tensor* gh, const yad_tensor* gw, const yad_tensor* dy, const yad_tensor* dx, int acc,
                        const yad_tensor* dgh, const yad_tensor* dgw, int dtype, void* stream) {
  CHECK_VIEW(dy, "rowcol_gate_bwd dy");
  CHECK_VIEW(gh, "rowcol_gate_bwd gh");
  CHECK_VIEW(gw, "rowcol_gate_bwd gw");
  CHECK_VIEW(dgh, "rowcol_gate_bwd dgh");
  CHECK_VIEW(dgw, "rowcol_gate_bwd dgw");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(dy->h > dy->w ? dy->h : dy->w, dy->n, 2);
  yad_tensor xx = x ? *x : *dy;
  const int64_t total = (int64_t)dy->n * dy->h * dy->w * (dy->c / 8);
  YAD_DISPATCH_DTYPE(dtype, {
    YAD_LAUNCH(rowcol_gate_bwd_g_kernel<T>, grid, 128, dy->c * sizeof(float), st, xx, x != nullptr, *gh, *gw, *dy, *dgh, *dgw);
    if (x && dx) YAD_LAUNCH(rowcol_gate_bwd_x_kernel<T>, grid_for(total), TPB, 0, st, *gh, *gw, *dy, *dx, acc);
  })
  YAD_LAUNCH_CHECK("rowcol_gate_bwd");
  return 0;
}

/* MLCA backward: datt / dlocal fp32 [n][ls*ls][c] scratch, dG fp32 [ls][c] scratch */
int yad_mlca_bwd(const yad_tensor* x, const yad_tensor* dy, const float* local, const float* att, const float* w_global, const float* w_local,
                 int ksize, float local_weight, int local_size, float* datt, float* dlocal, float* dG, float* dw_global, float* dw_local,
                 const yad_tensor* dx, int acc, int dtype, void* stream) {
  CHECK_VIEW(x, "mlca_bwd x");
  CHECK_VIEW(dy, "mlca_bwd dy");
  CHECK_VIEW(dx, "mlca_bwd dx");
  SAME_SHAPE(x, dy, "mlca_bwd");
  SAME_SHAPE(x, dx, "mlca_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = local_size * local_size, c = x->c;
  cudaMemsetAsync(dG, 0, sizeof(float) * local_size * c, st);
  dim3 g1(nb, x->n);
  const size_t smem_a = (size_t)(2 * nb * c + 32) * sizeof(float);
  YAD_CHECK(smem_a <= 48 * 1024, "mlca_bwd: %d channels need %zu B of shared memory", c, smem_a);
  const int64_t total = (int64_t)x->n * x->h * x->w * (c / 8);
  YAD_DISPATCH_DTYPE(dtype, {
    YAD_LAUNCH(mlca_bwd_datt_kernel<T>, g1, 128, c * sizeof(float), st, *x, *dy, local_size, datt);
    YAD_LAUNCH(mlca_att_bwd_a_kernel, x->n, 256, smem_a, st, local, datt, w_local, ksize, local_weight, c, local_size, dlocal, dw_local, dG);
    YAD_LAUNCH(mlca_att_bwd_b_kernel, x->n, 128, (2 * c + 32) * sizeof(float), st, local, dG, w_global, ksize, c, local_size, x->n, dlocal, dw_global);
    YAD_LAUNCH(mlca_bwd_apply_kernel<T>, grid_for(total), TPB, 0, st, *dy, att, dlocal, local_size, *dx, acc);
  })
  YAD_LAUNCH_CHECK("mlca_bwd");
  return 0;
}

int yad_maxpool5_bwd(const yad_tensor* x, const void* dy_t, int dy_ld, const float* dy_f, float* dx_f, int dtype, void* stream) {
  CHECK_VIEW(x, "maxpool5_bwd");
  const int64_t total = (int64_t)x->n * x->h * x->w * x->c;
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(maxpool5_bwd_kernel<T>, grid_for(total), TPB, 0, (cudaStream_t)stream, *x, (const T*)dy_t, dy_ld, dy_f, dx_f);)
  YAD_LAUNCH_CHECK("maxpool5_bwd");
  return 0;
}

int yad_cast_acc(const float* src, float scale, const yad_tensor* y, int acc, int dtype, void* stream) {
  CHECK_VIEW(y, "cast_acc");
  const int64_t total = (int64_t)y->n * y->h * y->w * (y->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(cast_acc_kernel<T>, grid_for(total), TPB, 0, (cudaStream_t)stream, src, scale, *y, acc);)
  YAD_LAUNCH_CHECK("cast_acc");
  return 0;
}

/* dpool: fp32 [n][h/s][w/s][c] scratch */
int yad_pool_upsample_bwd(const yad_tensor* dy, int s, float* dpool, const yad_tensor* dx, int acc, int dtype, void* stream) {
  CHECK_VIEW(dy, "pool_upsample_bwd dy");
  CHECK_VIEW(dx, "pool_upsample_bwd dx");
  SAME_SHAPE(dy, dx, "pool_upsample_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  const int hp = dy->h / s, wp = dy->w / s;
  cudaMemsetAsync(dpool, 0, sizeof(float) * dy->n * hp * wp * dy->c, st);
  const int64_t total = (int64_t)dy->n * dy->h * dy->w * (dy->c / 8);
  YAD_DISPATCH_DTYPE(dtype, {
    YAD_LAUNCH(pool_upsample_bwd_scatter_kernel<T>, grid_for(total), TPB, 0, st, *dy, s, dpool);
    YAD_LAUNCH(pool_upsample_bwd_gather_kernel<T>, grid_for(total), TPB, 0, st, dpool, s, *dx, acc);
  })
  YAD_LAUNCH_CHECK("pool_upsample_bwd");
  return 0;
}

int yad_gate_mlp_bwd(const float* g, const float* w1, const float* b1, const float* w2, const float* b2, int n, int c, int hidden, int nout,
                     int kind, const float* dout, float* dg, float* dw1, float* db1, float* dw2, float* db2, void* stream) {
  YAD_LAUNCH(gate_mlp_bwd_kernel, n, 128, (2 * hidden + 2 * nout) * sizeof(float), (cudaStream_t)stream, g, w1, b1, w2, b2, c, hidden, nout, kind, dout, dg,
                                                                                                  dw1, db1, dw2, db2);
  YAD_LAUNCH_CHECK("gate_mlp_bwd");
  return 0;
}

int yad_adt_bwd(const yad_tensor* x, const yad_tensor* dy, const float* imp, const float* alphas, const float* weight, const yad_tensor* dx, int acc,
                float* dimp, float* dalpha, float* dweight, float* dbias, int dtype, void* stream) {
  CHECK_VIEW(x, "adt_bwd x");
  CHECK_VIEW(dy, "adt_bwd dy");
  CHECK_VIEW(dx, "adt_bwd dx");
  SAME_SHAPE(x, dy, "adt_bwd");
  SAME_SHAPE(x, dx, "adt_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dimp, 0, sizeof(float) * 3 * x->n, st);
  const int64_t hw = (int64_t)x->h * x->w;
  int chunks = (int)((hw * (x->c / 8) + TPB * 4 - 1) / (TPB * 4));
  chunks = chunks < 1 ? 1 : (chunks > 32 ? 32 : chunks);
  dim3 grid(chunks, x->n);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(adt_bwd_kernel<T>, grid, TPB, 2 * x->c * sizeof(float), st, *x, *dy, imp, alphas, weight, *dx, acc, dimp, dalpha,
                                                                                            dweight, dbias);)
  YAD_LAUNCH_CHECK("adt_bwd");
  return 0;
}

int yad_gelu_gate_bwd(const yad_tensor* a, const void* b, int b_ld, const yad_tensor* dy, const yad_tensor* da, void* db, int db_ld, int acc,
                      int dtype, void* stream) {
  CHECK_VIEW(a, "gelu_gate_bwd a");
  CHECK_VIEW(dy, "gelu_gate_bwd dy");
  CHECK_VIEW(da, "gelu_gate_bwd da");
  const int64_t total = (int64_t)a->n * a->h * a->w * (a->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(gelu_gate_bwd_kernel<T>, grid_for(total), TPB, 0, (cudaStream_t)stream, *a, (const T*)b, b_ld, *dy, *da, (const T*)db,
                                                                                                        db_ld, acc);)
  YAD_LAUNCH_CHECK("gelu_gate_bwd");
  return 0;
}

int yad_scale_img(const yad_tensor* a, const float* s, const yad_tensor* y, int acc, int dtype, void* stream) {
  CHECK_VIEW(a, "scale_img a");
  CHECK_VIEW(y, "scale_img y");
  SAME_SHAPE(a, y, "scale_img");
  const int64_t total = (int64_t)a->n * a->h * a->w * (a->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(scale_img_kernel<T>, grid_for(total), TPB, 0, (cudaStream_t)stream, *a, s, *y, acc);)
  YAD_LAUNCH_CHECK("scale_img");
  return 0;
}

int yad_group_mean_bwd(const yad_tensor* dy, int s, const yad_tensor* dx, int acc, int dtype, void* stream) {
  CHECK_VIEW(dy, "group_mean_bwd dy");
  CHECK_VIEW(dx, "group_mean_bwd dx");
  YAD_CHECK(dx->n == dy->n && dx->c == dy->c && (int64_t)dx->h * dx->w == (int64_t)s * dy->h * dy->w, "group_mean_bwd: shape mismatch");
  const int64_t total = (int64_t)dx->n * dx->h * dx->w * (dx->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(group_mean_bwd_kernel<T>, grid_for(total), TPB, 0, (cudaStream_t)stream, *dy, s, *dx, acc);)
  YAD_LAUNCH_CHECK("group_mean_bwd");
  return 0;
}

/* dx_f: fp32 dense (n,h,w,c), zeroed by this call; dm: fp32 [64][64][c] accumulated (may be NULL) */
int yad_patch_filter_bwd(const yad_tensor* x, const yad_tensor* dy, const float* m, float alpha, float* dx_f, float* dm, int dtype, void* stream) {
  CHECK_VIEW(x, "patch_filter_bwd x");
  CHECK_VIEW(dy, "patch_filter_bwd dy");
  SAME_SHAPE(x, dy, "patch_filter_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dx_f, 0, sizeof(float) * (int64_t)x->n * x->h * x->w * x->c, st);
  dim3 grid(((x->h + 7) / 8) * ((x->w + 7) / 8), x->n);
  int tpb = x->c < 128 ? ((x->c + 31) / 32) * 32 : 128;
  YAD_DISPATCH_DTYPE(dtype, {
    YAD_LAUNCH(patch_filter_bwd_x_kernel<T>, grid, tpb, 0, st, *dy, m, alpha, dx_f);
    if (dm) YAD_LAUNCH(patch_filter_bwd_m_kernel<T>, x->c, 256, 0, st, *x, *dy, alpha, dm);
  })
  YAD_LAUNCH_CHECK("patch_filter_bwd");
  return 0;
}

int yad_head_pack(const yad_tensor* level, int anchor0, int n_anchors, int reg_ch, int nc, float* distri, float* logits, int dtype, void* stream) {
  CHECK_VIEW(level, "head_pack");
  YAD_CHECK(nc >= 1 && level->c == reg_ch + (nc + 7) / 8 * 8 && reg_ch % 8 == 0, "head_pack: level has %d channels, expected %d + %d classes padded to a multiple of 8",
            level->c, reg_ch, nc);
  const int64_t total = (int64_t)level->n * level->h * level->w * (level->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(head_pack_kernel<T>, grid_for(total), TPB, 0, (cudaStream_t)stream, *level, anchor0, n_anchors, reg_ch, nc, distri, logits);)
  YAD_LAUNCH_CHECK("head_pack");
  return 0;
}

int yad_head_unpack(const float* grad_distri, const float* grad_logits, float scale, int anchor0, int n_anchors, int reg_ch, int nc,
                    const yad_tensor* level, int dtype, void* stream) {
  CHECK_VIEW(level, "head_unpack");
  YAD_CHECK(nc >= 1 && level->c == reg_ch + (nc + 7) / 8 * 8 && reg_ch % 8 == 0, "head_unpack: level has %d channels, expected %d + %d classes padded to a multiple of 8",
            level->c, reg_ch, nc);
  const int64_t total = (int64_t)level->n * level->h * level->w * (level->c / 8);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(head_unpack_kernel<T>, grid_for(total), TPB, 0, (cudaStream_t)stream, grad_distri, grad_logits, scale, anchor0, n_anchors,
                                                                                                      reg_ch, nc, *level);)
  YAD_LAUNCH_CHECK("head_unpack");
  return 0;
}

int yad_fusion_weights(const float* p, int k, float* w, const float* dw, float* dp, void* stream) {
  YAD_CHECK(k >= 1 && k <= 8, "fusion_weights: k = %d", k);
  YAD_LAUNCH(fusion_weights_kernel, 1, 32, 0, (cudaStream_t)stream, p, k, w, dw, dp);
  YAD_LAUNCH_CHECK("fusion_weights");
  return 0;
}

int yad_small_gemm(const float* a, const float* b, float* c, int m, int n, int k, int trans_a, int acc, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (trans_a) {
    if (!acc) cudaMemsetAsync(c, 0, sizeof(float) * (size_t)m * n, st);
    const int chunk = 128;
    dim3 grid(n, (k + chunk - 1) / chunk);
    YAD_LAUNCH(small_gemm_tn_kernel, grid, 128, 0, st, a, b, c, m, n, k, chunk);
  } else {
    YAD_LAUNCH(small_gemm_kernel, grid_for((int64_t)m * n), TPB, 0, st, a, b, c, m, n, k, acc);
  }
  YAD_LAUNCH_CHECK("small_gemm");
  return 0;
}

}  // extern "C"
