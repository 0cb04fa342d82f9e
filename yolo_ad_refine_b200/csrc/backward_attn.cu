// Backward kernels of the layer-10 token blocks (training path):
// CrossScaleAttentionTSSA statistics (nn/modules/block.py:2459-2474) and the nn.MultiheadAttention core (block.py:2432-2434, 2479-2486).
#include "common.cuh"

namespace {

// ---- TSSA backward: one block per (image, head) -------------------------------------------------------------------------------------------
// forward: s_t = sum_j normalize(q_t)_j^2 ; Pi = softmax_t(s_t * temp) ; dots_j = sum_t Pi_t k_tj^2 ; a_j = 1/(1+dots_j) ; out_tj = -v_tj Pi_t a_j.
// s_t == 1 for every non-zero q_t, so d(out)/dq vanishes identically: dq is written as zeros (the reference's autograd produces rounding
// noise of the order 1e-9 there); dtemp is computed as written (it sums to ~0 for the same reason).
template <typename T>
__global__ void tssa_bwd_kernel(const T* __restrict__ qkv, int64_t ld, int Tn, int c, int heads, const float* __restrict__ temps,
                                const T* __restrict__ dout, int64_t dout_ld, int Tout, int tok_off, T* __restrict__ dqkv, int64_t dld,
                                float* __restrict__ dtemps) {
  extern __shared__ float sm[];  // pi[Tn], st[Tn], dpi[Tn], dots[d], dattn[d], red[32]
  const int d = c / heads, n = blockIdx.x / heads, h = blockIdx.x % heads;
  float* pi = sm;
  float* st = pi + Tn;
  float* dpi = st + Tn;
  float* dots = dpi + Tn;
  float* dattn = dots + d;
  float* red = dattn + d;
  const T* base = qkv + (int64_t)n * Tn * ld + h * d;
  const T* gbase = dout + ((int64_t)n * Tout + tok_off) * dout_ld + h * d;
  T* obase = dqkv + (int64_t)n * Tn * dld + h * d;
  const float temp = temps[h];
  for (int i = threadIdx.x; i < d; i += blockDim.x) { dots[i] = 0.f; dattn[i] = 0.f; }
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    const T* q = base + (int64_t)t * ld;
    float ss = 0.f;
    for (int o = 0; o < d; o += 8) {
      float v[8];
      load8(q + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) ss = fmaf(v[i], v[i], ss);
    }
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    float s = 0.f;
    for (int o = 0; o < d; o += 8) {
      float v[8];
      load8(q + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) { float w = v[i] * inv; s = fmaf(w, w, s); }
    }
    st[t] = s;
    pi[t] = s * temp;
  }
  __syncthreads();
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
  // dots_j and dattn_j = sum_t -v_tj Pi_t dout_tj
  const int oct = d >> 3;
  {
    const int o = (threadIdx.x % oct) * 8, lane = threadIdx.x / oct, step = blockDim.x / oct;
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = 0.f; b[i] = 0.f; }
    if (lane < step) {
      for (int t = lane; t < Tn; t += step) {
        float kv[8], vv[8], g[8];
        load8(base + (int64_t)t * ld + c + o, kv);
        load8(base + (int64_t)t * ld + 2 * c + o, vv);
        load8(gbase + (int64_t)t * dout_ld + o, g);
        const float p = pi[t];
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = fmaf(p * kv[i], kv[i], a[i]); b[i] = fmaf(-vv[i] * p, g[i], b[i]); }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) { atomicAdd(&dots[o + i], a[i]); atomicAdd(&dattn[o + i], b[i]); }
    }
  }
  __syncthreads();
  // per token: dv, dk, dPi
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    const float p = pi[t];
    float acc = 0.f;
    for (int o = 0; o < d; o += 8) {
      float kv[8], vv[8], g[8], dk[8], dv[8], z[8];
      load8(base + (int64_t)t * ld + c + o, kv);
      load8(base + (int64_t)t * ld + 2 * c + o, vv);
      load8(gbase + (int64_t)t * dout_ld + o, g);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float at = 1.0f / (1.0f + dots[o + i]);
        const float ddots = -at * at * dattn[o + i];
        dv[i] = -p * at * g[i];
        dk[i] = 2.0f * ddots * p * kv[i];
        acc += -vv[i] * at * g[i] + ddots * kv[i] * kv[i];
        z[i] = 0.f;
      }
      store8(obase + (int64_t)t * dld + o, z);
      store8(obase + (int64_t)t * dld + c + o, dk);
      store8(obase + (int64_t)t * dld + 2 * c + o, dv);
    }
    dpi[t] = acc;
  }
  __syncthreads();
  float dotp = 0.f;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) dotp = fmaf(pi[t], dpi[t], dotp);
  dotp = block_sum(dotp, red);
  float dt = 0.f;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) dt = fmaf(pi[t] * (dpi[t] - dotp), st[t], dt);
  dt = block_sum(dt, red);
  if (threadIdx.x == 0 && dtemps) atomicAdd(&dtemps[h], dt);
}

// ---- multi-head attention backward, head_dim 64: 4 lanes per row, 16 dims per lane, fp32 arithmetic --------------------------------------
constexpr int HD = 64, RB = 32, KB = 32;

__device__ __forceinline__ float reduce4(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

template <typename T>
__device__ __forceinline__ void load16(const T* p, float (&v)[16]) {
  float a[8], b[8];
  load8(p, a);
  load8(p + 8, b);
#pragma unroll
  for (int i = 0; i < 8; i++) { v[i] = a[i]; v[8 + i] = b[i]; }
}
template <typename T>
__device__ __forceinline__ void store16(T* p, const float (&v)[16]) {
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = v[i]; b[i] = v[8 + i]; }
  store8(p, a);
  store8(p + 8, b);
}

// dQ + the per-row softmax statistics (lse, D = dO . O) for the dK/dV kernel.  grid (ceil(T/32), n*heads), block 128
template <typename T>
__global__ void __launch_bounds__(128) mha_bwd_dq_kernel(const T* __restrict__ qkv, int64_t ld, int Tn, int c, int heads, const T* __restrict__ out,
                                                         int64_t out_ld, const T* __restrict__ dout, int64_t dout_ld, float scale,
                                                         T* __restrict__ dqkv, int64_t dld, float* __restrict__ lse_d) {
  __shared__ __align__(16) float Ks[KB][HD];
  __shared__ __align__(16) float Vs[KB][HD];
  const int n = blockIdx.y / heads, h = blockIdx.y % heads;
  const int row = blockIdx.x * RB + (threadIdx.x >> 2), part = (threadIdx.x & 3) * 16;
  const bool valid = row < Tn;
  const int rr = valid ? row : Tn - 1;
  const T* base = qkv + (int64_t)n * Tn * ld + h * HD;
  float q[16], go[16], dq[16];
  load16(base + (int64_t)rr * ld + part, q);
  load16(dout + ((int64_t)n * Tn + rr) * dout_ld + h * HD + part, go);
  float D;
  {
    float o[16];
    load16(out + ((int64_t)n * Tn + rr) * out_ld + h * HD + part, o);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) { s = fmaf(go[i], o[i], s); q[i] *= scale; dq[i] = 0.f; }
    D = reduce4(s);
  }
  float m = -INFINITY, l = 0.f, lse = 0.f;
  for (int pass = 0; pass < 2; pass++) {
    for (int k0 = 0; k0 < Tn; k0 += KB) {
      __syncthreads();
      for (int ch = threadIdx.x; ch < KB * (HD / 8); ch += 128) {
        const int j = ch / (HD / 8), o = (ch % (HD / 8)) * 8;
        float kv[8], vv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { kv[i] = 0.f; vv[i] = 0.f; }
        if (k0 + j < Tn) {
          load8(base + (int64_t)(k0 + j) * ld + c + o, kv);
          if (pass) load8(base + (int64_t)(k0 + j) * ld + 2 * c + o, vv);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) { Ks[j][o + i] = kv[i]; Vs[j][o + i] = vv[i]; }
      }
      __syncthreads();
      const int kn = min(KB, Tn - k0);
      for (int j = 0; j < kn; j++) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) s = fmaf(q[i], Ks[j][part + i], s);
        s = reduce4(s);
        if (pass == 0) {
          const float mn = fmaxf(m, s);
          l = l * expf(m - mn) + expf(s - mn);
          m = mn;
        } else {
          const float p = expf(s - lse);
          float dp = 0.f;
#pragma unroll
          for (int i = 0; i < 16; i++) dp = fmaf(go[i], Vs[j][part + i], dp);
          dp = reduce4(dp);
          const float ds = p * (dp - D);
#pragma unroll
          for (int i = 0; i < 16; i++) dq[i] = fmaf(ds, Ks[j][part + i], dq[i]);
        }
      }
    }
    if (pass == 0) lse = m + logf(l);
  }
  if (valid) {
#pragma unroll
    for (int i = 0; i < 16; i++) dq[i] *= scale;
    store16(dqkv + ((int64_t)n * Tn + row) * dld + h * HD + part, dq);
    if (part == 0) {
      lse_d[((int64_t)blockIdx.y * Tn + row) * 2] = lse;
      lse_d[((int64_t)blockIdx.y * Tn + row) * 2 + 1] = D;
    }
  }
}

// dK, dV: one key row per 4 lanes; loops over query tiles.  grid (ceil(T/32), n*heads), block 128
template <typename T>
__global__ void __launch_bounds__(128) mha_bwd_dkv_kernel(const T* __restrict__ qkv, int64_t ld, int Tn, int c, int heads, const T* __restrict__ dout,
                                                          int64_t dout_ld, float scale, const float* __restrict__ lse_d, T* __restrict__ dqkv,
                                                          int64_t dld) {
  __shared__ __align__(16) float Qs[RB][HD];
  __shared__ __align__(16) float Gs[RB][HD];
  __shared__ float Ls[RB], Ds[RB];
  const int n = blockIdx.y / heads, h = blockIdx.y % heads;
  const int row = blockIdx.x * RB + (threadIdx.x >> 2), part = (threadIdx.x & 3) * 16;
  const bool valid = row < Tn;
  const int rr = valid ? row : Tn - 1;
  const T* base = qkv + (int64_t)n * Tn * ld + h * HD;
  float k[16], v[16], dk[16], dv[16];
  load16(base + (int64_t)rr * ld + c + part, k);
  load16(base + (int64_t)rr * ld + 2 * c + part, v);
#pragma unroll
  for (int i = 0; i < 16; i++) { dk[i] = 0.f; dv[i] = 0.f; }
  for (int q0 = 0; q0 < Tn; q0 += RB) {
    __syncthreads();
    for (int ch = threadIdx.x; ch < RB * (HD / 8); ch += 128) {
      const int j = ch / (HD / 8), o = (ch % (HD / 8)) * 8;
      float qv[8], gv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) { qv[i] = 0.f; gv[i] = 0.f; }
      if (q0 + j < Tn) {
        load8(base + (int64_t)(q0 + j) * ld + o, qv);
        load8(dout + ((int64_t)n * Tn + q0 + j) * dout_ld + h * HD + o, gv);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) { Qs[j][o + i] = qv[i] * scale; Gs[j][o + i] = gv[i]; }
    }
    if (threadIdx.x < RB) {
      const int t = q0 + threadIdx.x;
      Ls[threadIdx.x] = t < Tn ? lse_d[((int64_t)blockIdx.y * Tn + t) * 2] : INFINITY;
      Ds[threadIdx.x] = t < Tn ? lse_d[((int64_t)blockIdx.y * Tn + t) * 2 + 1] : 0.f;
    }
    __syncthreads();
    const int qn = min(RB, Tn - q0);
    for (int j = 0; j < qn; j++) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < 16; i++) { s = fmaf(Qs[j][part + i], k[i], s); dp = fmaf(Gs[j][part + i], v[i], dp); }
      s = reduce4(s);
      dp = reduce4(dp);
      const float p = expf(s - Ls[j]);
      const float ds = p * (dp - Ds[j]);
#pragma unroll
      for (int i = 0; i < 16; i++) { dv[i] = fmaf(p, Gs[j][part + i], dv[i]); dk[i] = fmaf(ds, Qs[j][part + i], dk[i]); }
    }
  }
  if (valid) {
    store16(dqkv + ((int64_t)n * Tn + row) * dld + c + h * HD + part, dk);
    store16(dqkv + ((int64_t)n * Tn + row) * dld + 2 * c + h * HD + part, dv);
  }
}

}  // namespace

extern "C" {

/* gradient of yad_tssa for one scale: dout = tokens [tok_off, tok_off + T) of the (n, T_out, 1, c) gradient; dqkv (n, h, w, 3c) is
 * overwritten (dq = 0, see the kernel comment); dtemps fp32 [heads] accumulated */
int yad_tssa_bwd(const yad_tensor* qkv, const float* temps, int heads, const yad_tensor* dout, int out_token_offset, const yad_tensor* dqkv,
                 float* dtemps, int dtype, void* stream) {
  const int Tn = qkv->h * qkv->w, c = dout->c;
  YAD_CHECK(qkv->c == 3 * c && dqkv->c == 3 * c && c % heads == 0 && (c / heads) % 8 == 0, "tssa_bwd: channel mismatch");
  const int Tout = dout->h * dout->w, d = c / heads;
  YAD_CHECK(dout->n == qkv->n && dqkv->n == qkv->n && dqkv->h * dqkv->w == Tn && Tout >= Tn + out_token_offset, "tssa_bwd: token count mismatch");
  YAD_CHECK(256 % (d / 8) == 0, "tssa_bwd: head_dim %d unsupported", d);
  size_t smem = (size_t)(3 * Tn + 2 * d + 32) * sizeof(float);
  YAD_CHECK(smem <= 200 * 1024, "tssa_bwd: %d tokens do not fit in shared memory", Tn);
  cudaStream_t st = (cudaStream_t)stream;
  YAD_DISPATCH_DTYPE(dtype, {
    if (smem > 48 * 1024) cudaFuncSetAttribute(tssa_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tssa_bwd_kernel<T><<<qkv->n * heads, 256, smem, st>>>((const T*)qkv->ptr, qkv->ld, Tn, c, heads, temps, (const T*)dout->ptr, dout->ld, Tout,
                                                         out_token_offset, (T*)dqkv->ptr, dqkv->ld, dtemps);
  })
  YAD_LAUNCH_CHECK("tssa_bwd");
  return 0;
}

/* gradient of yad_mha: qkv (n,1,T,3c), out / dout (n,1,T,c) -> dqkv (n,1,T,3c) overwritten.  lse_d: fp32 [n*heads][T][2] scratch */
int yad_mha_bwd(const yad_tensor* qkv, int heads, const yad_tensor* out, const yad_tensor* dout, const yad_tensor* dqkv, float* lse_d, int dtype,
                void* stream) {
  const int Tn = qkv->h * qkv->w, c = out->c;
  YAD_CHECK(qkv->c == 3 * c && dqkv->c == 3 * c && c == heads * HD, "mha_bwd: only head_dim 64 is built (c=%d, heads=%d)", c, heads);
  YAD_CHECK(out->n == qkv->n && out->h * out->w == Tn && dout->h * dout->w == Tn && dqkv->h * dqkv->w == Tn, "mha_bwd: token count mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = 1.0f / sqrtf((float)HD);
  dim3 grid((Tn + RB - 1) / RB, qkv->n * heads);
  YAD_DISPATCH_DTYPE(dtype, {
    mha_bwd_dq_kernel<T><<<grid, 128, 0, st>>>((const T*)qkv->ptr, qkv->ld, Tn, c, heads, (const T*)out->ptr, out->ld, (const T*)dout->ptr, dout->ld,
                                              scale, (T*)dqkv->ptr, dqkv->ld, lse_d);
    mha_bwd_dkv_kernel<T><<<grid, 128, 0, st>>>((const T*)qkv->ptr, qkv->ld, Tn, c, heads, (const T*)dout->ptr, dout->ld, scale, lse_d,
                                               (T*)dqkv->ptr, dqkv->ld);
  })
  YAD_LAUNCH_CHECK("mha_bwd");
  return 0;
}

}  // extern "C"
