// Layer-10 token kernels: CrossScaleAttentionTSSA statistics (nn/modules/block.py:2459-2474) and the core of
// nn.MultiheadAttention (block.py:2432-2434, 2479-2486) as a flash-style streaming-softmax kernel (fp32 accumulate).
#include "common.cuh"

namespace {

// ---- TSSA: one block per (image, head) ---------------------------------------------------------------------------
// qkv: (n, T, 3c) tokens; q | k | v each c wide, head h owns channels [h*d, (h+1)*d).
template <typename T>
__global__ void tssa_kernel(const T* __restrict__ qkv, int64_t ld, int Tn, int c, int heads, const float* __restrict__ temps,
                            T* __restrict__ out, int64_t out_ld, int Tout, int tok_off) {
  extern __shared__ float sm[];  // pi[Tn], dots[d], red[32]
  const int d = c / heads, n = blockIdx.x / heads, h = blockIdx.x % heads;
  float* pi = sm;
  float* dots = sm + Tn;
  float* red = dots + d;
  const T* base = qkv + (int64_t)n * Tn * ld + h * d;
  const float temp = temps[h];
  for (int i = threadIdx.x; i < d; i += blockDim.x) dots[i] = 0.f;
  // 1. logits[t] = sum_d normalize(q)^2 * temp
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    const T* q = base + (int64_t)t * ld;
    float ss = 0.f;
    for (int o = 0; o < d; o += 8) {
      float v[8];
      load8(q + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) ss = fmaf(v[i], v[i], ss);
    }
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
    float s = 0.f;
    for (int o = 0; o < d; o += 8) {
      float v[8];
      load8(q + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) { float w = v[i] * inv; s = fmaf(w, w, s); }
    }
    pi[t] = s * temp;
  }
  __syncthreads();
  // 2. softmax over tokens
  float m = -INFINITY;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) m = fmaxf(m, pi[t]);
  m = block_max(m, red);
  float sum = 0.f;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) { float e = expf(pi[t] - m); pi[t] = e; sum += e; }
  sum = block_sum(sum, red);
  __syncthreads();
  const float inv_sum = 1.0f / sum;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) pi[t] *= inv_sum;
  __syncthreads();
  // 3. dots[dd] = sum_t pi[t] * k[t][dd]^2
  const int oct = d >> 3;
  {
    const int o = (threadIdx.x % oct) * 8, lane = threadIdx.x / oct, step = blockDim.x / oct;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    if (lane < step) {
      for (int t = lane; t < Tn; t += step) {
        float v[8];
        load8(base + (int64_t)t * ld + c + o, v);
        const float p = pi[t];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fmaf(p * v[i], v[i], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) atomicAdd(&dots[o + i], acc[i]);
    }
  }
  __syncthreads();
  // 4. out = -(v * pi) * 1/(1+dots)
  const int64_t items = (int64_t)Tn * oct;
  for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
    int t = (int)(it / oct), o = (int)(it % oct) * 8;
    float v[8];
    load8(base + (int64_t)t * ld + 2 * c + o, v);
    const float p = pi[t];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = -(v[i] * p) * (1.0f / (1.0f + dots[o + i]));
    store8(out + ((int64_t)n * Tout + tok_off + t) * out_ld + h * d + o, v);
  }
}

// ---- multi-head self-attention core, head_dim 64 -------------------------------------------------------------------
constexpr int HD = 64, QB = 128, KB = 32;

template <typename T>
__global__ void __launch_bounds__(QB) mha_kernel(const T* __restrict__ qkv, int64_t ld, int Tn, int c, int heads, T* __restrict__ out,
                                                 int64_t out_ld, float scale) {
  __shared__ __align__(16) float Ks[KB][HD];
  __shared__ __align__(16) float Vs[KB][HD];
  const int n = blockIdx.y / heads, h = blockIdx.y % heads;
  const int tq = blockIdx.x * QB + threadIdx.x;
  const bool qvalid = tq < Tn;
  const T* base = qkv + (int64_t)n * Tn * ld + h * HD;
  float q[HD], acc[HD];
#pragma unroll
  for (int i = 0; i < HD; i++) { q[i] = 0.f; acc[i] = 0.f; }
  if (qvalid) {
#pragma unroll
    for (int o = 0; o < HD; o += 8) {
      float v[8];
      load8(base + (int64_t)tq * ld + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) q[o + i] = v[i] * scale;
    }
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < Tn; k0 += KB) {
    __syncthreads();
    for (int ch = threadIdx.x; ch < KB * (HD / 8); ch += QB) {
      int j = ch / (HD / 8), o = (ch % (HD / 8)) * 8;
      float kv[8], vv[8];
      if (k0 + j < Tn) {
        load8(base + (int64_t)(k0 + j) * ld + c + o, kv);
        load8(base + (int64_t)(k0 + j) * ld + 2 * c + o, vv);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) { kv[i] = 0.f; vv[i] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) { Ks[j][o + i] = kv[i]; Vs[j][o + i] = vv[i]; }
    }
    __syncthreads();
    float s[KB];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < KB; j++) {
      float a = 0.f;
#pragma unroll
      for (int o = 0; o < HD; o += 4) {
        float4 kk = *reinterpret_cast<const float4*>(&Ks[j][o]);
        a = fmaf(q[o], kk.x, a); a = fmaf(q[o + 1], kk.y, a); a = fmaf(q[o + 2], kk.z, a); a = fmaf(q[o + 3], kk.w, a);
      }
      s[j] = (k0 + j < Tn) ? a : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float mn = fmaxf(m, tmax);
    const float corr = expf(m - mn);  // m = -inf on the first tile -> 0
    l *= corr;
#pragma unroll
    for (int i = 0; i < HD; i++) acc[i] *= corr;
#pragma unroll
    for (int j = 0; j < KB; j++) {
      const float p = expf(s[j] - mn);
      l += p;
#pragma unroll
      for (int o = 0; o < HD; o += 4) {
        float4 vv = *reinterpret_cast<const float4*>(&Vs[j][o]);
        acc[o] = fmaf(p, vv.x, acc[o]); acc[o + 1] = fmaf(p, vv.y, acc[o + 1]);
        acc[o + 2] = fmaf(p, vv.z, acc[o + 2]); acc[o + 3] = fmaf(p, vv.w, acc[o + 3]);
      }
    }
    m = mn;
  }
  if (qvalid) {
    const float inv = 1.0f / l;
    T* dst = out + ((int64_t)n * Tn + tq) * out_ld + h * HD;
#pragma unroll
    for (int o = 0; o < HD; o += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = acc[o + i] * inv;
      store8(dst + o, v);
    }
  }
}

}  // namespace

extern "C" {

int yad_tssa(const yad_tensor* qkv, const float* temps, int heads, const yad_tensor* out, int out_token_offset, int dtype, void* stream) {
  const int Tn = qkv->h * qkv->w, c = out->c;
  YAD_CHECK(qkv->c == 3 * c && c % heads == 0 && (c / heads) % 8 == 0, "tssa: qkv has %d channels, out %d, heads %d", qkv->c, c, heads);
  const int Tout = out->h * out->w;
  YAD_CHECK(out->n == qkv->n && out_token_offset >= 0 && Tout >= Tn + out_token_offset, "tssa: token count mismatch");
  const int d = c / heads;
  size_t smem = (size_t)(Tn + d + 32) * sizeof(float);
  YAD_CHECK(smem <= 200 * 1024, "tssa: %d tokens do not fit in shared memory", Tn);
  cudaStream_t st = (cudaStream_t)stream;
  YAD_DISPATCH_DTYPE(dtype, {
    if (smem > 48 * 1024) cudaFuncSetAttribute(tssa_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tssa_kernel<T><<<qkv->n * heads, 256, smem, st>>>((const T*)qkv->ptr, qkv->ld, Tn, c, heads, temps, (T*)out->ptr, out->ld, Tout,
                                                     out_token_offset);
  })
  YAD_LAUNCH_CHECK("tssa");
  return 0;
}

int yad_mha(const yad_tensor* qkv, int heads, const yad_tensor* out, int dtype, void* stream) {
  const int Tn = qkv->h * qkv->w, c = out->c;
  YAD_CHECK(qkv->c == 3 * c && c == heads * HD, "mha: only head_dim 64 is built (c=%d, heads=%d)", c, heads);
  YAD_CHECK(out->n == qkv->n && out->h * out->w == Tn, "mha: token count mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((Tn + QB - 1) / QB, qkv->n * heads);
  const float scale = 1.0f / sqrtf((float)HD);
  YAD_DISPATCH_DTYPE(dtype, mha_kernel<T><<<grid, QB, 0, st>>>((const T*)qkv->ptr, qkv->ld, Tn, c, heads, (T*)out->ptr, out->ld, scale);)
  YAD_LAUNCH_CHECK("mha");
  return 0;
}

}  // extern "C"
