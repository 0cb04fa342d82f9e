// Layer-10 token kernels: CrossScaleAttentionTSSA statistics (nn/modules/block.py:2459-2474) and the core of
// nn.MultiheadAttention (block.py:2432-2434, 2479-2486) as a flash-style streaming-softmax kernel (fp32 accumulate).
#include <stdlib.h>

#include "common.cuh"

namespace {

// ---- TSSA: one block per (image, head) ---------------------------------------------------------------------------
// qkv: (n, T, 3c) tokens; q | k | v each c wide, head h owns channels [h*d, (h+1)*d).
template <typename T>
__global__ void tssa_kernel(const T* __restrict__ qkv, int64_t ld, int Tn, int c, int heads, const float* __restrict__ temps,
                            T* __restrict__ out, int64_t out_ld, int Tout, int tok_off) {
  pdl_sync();
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
  pdl_sync();
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

// ---- multi-head self-attention core on tensor cores (bf16 storage): flash-style streaming softmax ---------------------------------
// One CTA = 64 queries of one (image, head): 4 warps x 16 query rows.  S = Q K^T and O += P V run on mma.sync.m16n8k16 (bf16 -> fp32);
// K / V tiles of 64 keys are staged in shared memory with a 144-byte row pitch (conflict-free ldmatrix), P never leaves registers
// (the S accumulator layout is re-used as the A fragment of the second GEMM).  [tcgen05 / TMEM version: next round.]
constexpr int FA_Q = 64, FA_K = 64, FA_PITCH = 72;  // pitch in bf16 elements

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(128) mha_mma_kernel(const bf16* __restrict__ qkv, int64_t ld, int Tn, int c, int heads, bf16* __restrict__ out,
                                                      int64_t out_ld, float scale_log2e) {
  pdl_sync();
  __shared__ __align__(16) bf16 Ks[2][FA_K * FA_PITCH];  // double buffer: tile k + 1 streams in (cp.async) while tile k is consumed
  __shared__ __align__(16) bf16 Vs[2][FA_K * FA_PITCH];
  const int n = blockIdx.y / heads, h = blockIdx.y % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
  const bf16* base = qkv + (int64_t)n * Tn * ld + h * HD;
  const int q0 = blockIdx.x * FA_Q + warp * 16;  // this warp's first query row
  // Q fragments (A operand, 16 x 64): rows g and g + 8, k = kt * 16 + 2 * q4 (+8)
  uint32_t qa[4][4];
#pragma unroll
  for (int kt = 0; kt < 4; kt++) {
    const int r0 = q0 + g, r1 = q0 + g + 8, col = kt * 16 + 2 * q4;
    qa[kt][0] = r0 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r0 * ld + col) : 0u;
    qa[kt][1] = r1 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r1 * ld + col) : 0u;
    qa[kt][2] = r0 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r0 * ld + col + 8) : 0u;
    qa[kt][3] = r1 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r1 * ld + col + 8) : 0u;
  }
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) o[j][i] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const uint32_t ks_base = (uint32_t)__cvta_generic_to_shared(&Ks[0][0]), vs_base = (uint32_t)__cvta_generic_to_shared(&Vs[0][0]);
  constexpr uint32_t BUF = FA_K * FA_PITCH * 2;
  auto stage = [&](int k0, int buf) {  // 64 rows x 8 chunks of 16 bytes, 8 lanes per row; rows beyond Tn are zero-filled (src-size 0)
    for (int ch = tid; ch < FA_K * 8; ch += 128) {
      const int row = ch >> 3, cc = (ch & 7) * 8;
      const bool ok = k0 + row < Tn;
      const bf16* g = base + (int64_t)(ok ? k0 + row : Tn - 1) * ld + cc;
      const uint32_t off = (uint32_t)((row * FA_PITCH + cc) * 2) + buf * BUF;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ks_base + off), "l"(g + c), "r"(ok ? 16 : 0) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vs_base + off), "l"(g + 2 * c), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0, 0);
  for (int k0 = 0, it = 0; k0 < Tn; k0 += FA_K, it++) {
    const int buf = it & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // tile `it` has landed; everybody is done with tile it - 1, whose buffer is refilled now
    if (k0 + FA_K < Tn) stage(k0 + FA_K, buf ^ 1);
    const uint32_t ks_addr = ks_base + buf * BUF, vs_addr = vs_base + buf * BUF;
    // S = Q K^T : 8 key tiles of 8
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
      for (int i = 0; i < 4; i++) s[j][i] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 4; kt++) {
        uint32_t b0, b1;  // B fragment: rows = keys j*8 + (lane & 7), two 8x8 matrices along d
        ldsm_x2(b0, b1, ks_addr + (uint32_t)(((j * 8 + (lane & 7)) * FA_PITCH + kt * 16 + ((lane >> 3) & 1) * 8) * 2));
        mma_bf16_16816(s[j], qa[kt], b0, b1);
      }
    }
    // mask the key tail, running max / sum (rows g and g + 8 of this warp)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int key = k0 + j * 8 + 2 * q4;
      if (key >= Tn) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (key + 1 >= Tn) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = exp2f((m0 - mn0) * scale_log2e), c1 = exp2f((m1 - mn1) * scale_log2e);  // m = -inf on the first tile -> 0
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int j = 0; j < 8; j++) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
    const float mb0 = mn0 * scale_log2e, mb1 = mn1 * scale_log2e;
    uint32_t pa[4][4];  // P as A fragments: key tile pairs (2kt, 2kt+1)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float p0 = exp2f(fmaf(s[j][0], scale_log2e, -mb0)), p1 = exp2f(fmaf(s[j][1], scale_log2e, -mb0));
      const float p2 = exp2f(fmaf(s[j][2], scale_log2e, -mb1)), p3 = exp2f(fmaf(s[j][3], scale_log2e, -mb1));
      l0 += p0 + p1; l1 += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    // O += P V : k = keys (4 tiles of 16), n = d (8 tiles of 8); V fragments via transposed ldmatrix
#pragma unroll
    for (int kt = 0; kt < 4; kt++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, vs_addr + (uint32_t)(((kt * 16 + (lane & 15)) * FA_PITCH + j * 8) * 2));
        mma_bf16_16816(o[j], pa[kt], b0, b1);
      }
    }
    m0 = mn0; m1 = mn1;
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int col = h * HD + j * 8 + 2 * q4;
    if (r0 < Tn) *reinterpret_cast<uint32_t*>(out + ((int64_t)n * Tn + r0) * out_ld + col) = pack_bf16(o[j][0] * i0, o[j][1] * i0);
    if (r1 < Tn) *reinterpret_cast<uint32_t*>(out + ((int64_t)n * Tn + r1) * out_ld + col) = pack_bf16(o[j][2] * i1, o[j][3] * i1);
  }
}


// ---- second version of the tensor-core attention core: 128 queries per CTA, 32 per warp -------------------------------------------
// ncu of mha_mma_kernel at the benchmark shape (batch 64, T = 1200, 2 heads x 64; profiles/r2_ncu_mha.md): tensor pipe 44 % active, issue slots
// 58 %, LSU 32 %, shared-memory wavefronts 48 % -- one ldmatrix.x2 per mma.sync, because a warp owns only one 16-row tile and every B fragment
// is used once.  Here a warp owns TWO 16-row query tiles, so each K / V fragment feeds two MMAs, and the fragments arrive four 8x8 matrices at a
// time (ldmatrix.x4: both k halves of a K fragment pair, two d tiles of a V fragment): 32 ldmatrix.x4 for 128 MMAs per 64-key tile instead of
// 128 ldmatrix.x2, half the shared-memory bytes per MMA, and half the L2 -> shared traffic for K / V (10 CTAs per (image, head) instead of 19).
// The key-tail mask is evaluated on the last tile only.
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

constexpr int FA2_Q = 128;

__global__ void __launch_bounds__(128, 2) mha_mma2_kernel(const bf16* __restrict__ qkv, int64_t ld, int Tn, int c, int heads, bf16* __restrict__ out,
                                                          int64_t out_ld, float scale_log2e) {
  pdl_sync();
  __shared__ __align__(16) bf16 Ks[2][FA_K * FA_PITCH];  // double buffer: tile k + 1 streams in (cp.async) while tile k is consumed
  __shared__ __align__(16) bf16 Vs[2][FA_K * FA_PITCH];
  const int n = blockIdx.y / heads, h = blockIdx.y % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
  const bf16* base = qkv + (int64_t)n * Tn * ld + h * HD;
  const int q0 = blockIdx.x * FA2_Q + warp * 32;  // this warp's first query row (two 16-row tiles)
  uint32_t qa[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; mt++) {
#pragma unroll
    for (int kt = 0; kt < 4; kt++) {
      const int r0 = q0 + mt * 16 + g, r1 = r0 + 8, col = kt * 16 + 2 * q4;
      qa[mt][kt][0] = r0 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r0 * ld + col) : 0u;
      qa[mt][kt][1] = r1 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r1 * ld + col) : 0u;
      qa[mt][kt][2] = r0 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r0 * ld + col + 8) : 0u;
      qa[mt][kt][3] = r1 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r1 * ld + col + 8) : 0u;
    }
  }
  float o[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int i = 0; i < 4; i++) o[mt][j][i] = 0.f;
  float mrun[2][2], lrun[2][2];  // [m tile][row g / row g + 8]
#pragma unroll
  for (int mt = 0; mt < 2; mt++) { mrun[mt][0] = mrun[mt][1] = -INFINITY; lrun[mt][0] = lrun[mt][1] = 0.f; }
  const uint32_t ks_base = (uint32_t)__cvta_generic_to_shared(&Ks[0][0]), vs_base = (uint32_t)__cvta_generic_to_shared(&Vs[0][0]);
  constexpr uint32_t BUF = FA_K * FA_PITCH * 2;
  auto stage = [&](int k0, int buf) {  // 64 rows x 8 chunks of 16 bytes, 8 lanes per row; rows beyond Tn are zero-filled (src-size 0)
    for (int ch = tid; ch < FA_K * 8; ch += 128) {
      const int row = ch >> 3, cc = (ch & 7) * 8;
      const bool ok = k0 + row < Tn;
      const bf16* gp = base + (int64_t)(ok ? k0 + row : Tn - 1) * ld + cc;
      const uint32_t off = (uint32_t)((row * FA_PITCH + cc) * 2) + buf * BUF;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ks_base + off), "l"(gp + c), "r"(ok ? 16 : 0) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vs_base + off), "l"(gp + 2 * c), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // per-lane ldmatrix.x4 offsets (bytes): K: row = key (lane & 7), the four matrices are the d columns (lane >> 3) * 8 of a 32-wide d block;
  // V (transposed): row = key (lane & 15), the matrices pair (keys 0-7 / 8-15) x (d tile j, j + 1)
  const uint32_t k_lane = (uint32_t)(((lane & 7) * FA_PITCH + (lane >> 3) * 8) * 2);
  const uint32_t v_lane = (uint32_t)(((lane & 15) * FA_PITCH + (lane >> 4) * 8) * 2);
  stage(0, 0);
  for (int k0 = 0, it = 0; k0 < Tn; k0 += FA_K, it++) {
    const int buf = it & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // tile `it` has landed; everybody is done with tile it - 1, whose buffer is refilled now
    if (k0 + FA_K < Tn) stage(k0 + FA_K, buf ^ 1);
    const uint32_t ks_addr = ks_base + buf * BUF + k_lane, vs_addr = vs_base + buf * BUF + v_lane;
    // S = Q K^T for both query tiles: 8 key tiles of 8, every K fragment pair (one ldmatrix.x4) feeds 4 MMAs
    float s[2][8][4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int i = 0; i < 4; i++) s[mt][j][i] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; kp++) {  // d block of 32 = k steps 2 kp, 2 kp + 1
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, ks_addr + (uint32_t)((j * 8 * FA_PITCH + kp * 32) * 2));
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          mma_bf16_16816(s[mt][j], qa[mt][2 * kp], b0, b1);
          mma_bf16_16816(s[mt][j], qa[mt][2 * kp + 1], b2, b3);
        }
      }
    }
    if (k0 + FA_K > Tn) {  // key tail (last tile only, warp-uniform)
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int key = k0 + j * 8 + 2 * q4;
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          if (key >= Tn) { s[mt][j][0] = -INFINITY; s[mt][j][2] = -INFINITY; }
          if (key + 1 >= Tn) { s[mt][j][1] = -INFINITY; s[mt][j][3] = -INFINITY; }
        }
      }
    }
    uint32_t pa[2][4][4];  // P as A fragments: key tile pairs (2 kt, 2 kt + 1)
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        mx0 = fmaxf(mx0, fmaxf(s[mt][j][0], s[mt][j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[mt][j][2], s[mt][j][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(mrun[mt][0], mx0), mn1 = fmaxf(mrun[mt][1], mx1);
      const float c0 = exp2f((mrun[mt][0] - mn0) * scale_log2e), c1 = exp2f((mrun[mt][1] - mn1) * scale_log2e);  // m = -inf on the first tile -> 0
      float l0 = lrun[mt][0] * c0, l1 = lrun[mt][1] * c1;
#pragma unroll
      for (int j = 0; j < 8; j++) { o[mt][j][0] *= c0; o[mt][j][1] *= c0; o[mt][j][2] *= c1; o[mt][j][3] *= c1; }
      const float mb0 = mn0 * scale_log2e, mb1 = mn1 * scale_log2e;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float p0 = exp2f(fmaf(s[mt][j][0], scale_log2e, -mb0)), p1 = exp2f(fmaf(s[mt][j][1], scale_log2e, -mb0));
        const float p2 = exp2f(fmaf(s[mt][j][2], scale_log2e, -mb1)), p3 = exp2f(fmaf(s[mt][j][3], scale_log2e, -mb1));
        l0 += p0 + p1; l1 += p2 + p3;
        pa[mt][j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
        pa[mt][j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      mrun[mt][0] = mn0; mrun[mt][1] = mn1; lrun[mt][0] = l0; lrun[mt][1] = l1;
    }
    // O += P V : k = keys (4 steps of 16), n = d (8 tiles of 8); one transposed ldmatrix.x4 = the V fragments of two d tiles, used by both query tiles
#pragma unroll
    for (int kt = 0; kt < 4; kt++) {
#pragma unroll
      for (int jp = 0; jp < 4; jp++) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(b0, b1, b2, b3, vs_addr + (uint32_t)((kt * 16 * FA_PITCH + jp * 16) * 2));
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          mma_bf16_16816(o[mt][2 * jp], pa[mt][kt], b0, b1);
          mma_bf16_16816(o[mt][2 * jp + 1], pa[mt][kt], b2, b3);
        }
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; mt++) {
    float l0 = lrun[mt][0], l1 = lrun[mt][1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int col = h * HD + j * 8 + 2 * q4;
      if (r0 < Tn) *reinterpret_cast<uint32_t*>(out + ((int64_t)n * Tn + r0) * out_ld + col) = pack_bf16(o[mt][j][0] * i0, o[mt][j][1] * i0);
      if (r1 < Tn) *reinterpret_cast<uint32_t*>(out + ((int64_t)n * Tn + r1) * out_ld + col) = pack_bf16(o[mt][j][2] * i1, o[mt][j][3] * i1);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// AttentionTSSA (nn/modules/block.py:1646-1683, the ToST token-statistics attention of the C2TSSA_DYT_Mona_EDFFN sibling blocks; SURVEY.md
// section 8f rank 3) on the projected tokens w = qkv(x), one CTA per image, fp32 statistics in shared memory:
//   pass 1  norm2[ch]  = sum_t w[t][ch]^2                                  (F.normalize over the TOKEN axis, :1670)
//   pass 2  s[t][h]    = temp[h] * sum_{ch in head h} w^2 / max(norm, 1e-12)^2
//           Pi[t][h]   = softmax over the HEADS (nn.Softmax(dim=1), :1653, 1674);  pisum[h] = sum_t Pi
//   pass 3  dots[ch]   = sum_t Pi[t][h(ch)] * w^2 / (pisum[h] + 1e-8)       (:1676)
//   pass 4  out[t][ch] = -w * Pi[t][h] / (1 + dots[ch])                     (:1677-1680)
// The token matrix of one image (T x c bf16, 100 KB at 20x20x128) stays in L2 between the passes.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AT_THREADS = 512;

template <typename T>
__global__ void __launch_bounds__(AT_THREADS) attention_tssa_kernel(const T* __restrict__ w, int64_t ld, int Tn, int c, int heads,
                                                                     const float* __restrict__ temp, T* __restrict__ out, int64_t out_ld) {
  pdl_sync();
  extern __shared__ float at_sm[];
  float* norm2 = at_sm;            // [c]   -> later 1 / max(norm, eps)^2
  float* dots = norm2 + c;         // [c]
  float* pisum = dots + c;         // [heads]
  float* pi = pisum + heads;       // [Tn * heads]
  const int tid = threadIdx.x, oct = c >> 3, o = (tid % oct) * 8, tl = tid / oct, step = AT_THREADS / oct;
  const int dh = c / heads, hd = o / dh;  // the 8 channels of a thread lie inside one head (dh % 8 == 0)
  const T* wp = w + (int64_t)blockIdx.x * Tn * ld;
  T* op = out + (int64_t)blockIdx.x * Tn * out_ld;
  for (int i = tid; i < 2 * c + heads + Tn * heads; i += AT_THREADS) at_sm[i] = 0.f;
  __syncthreads();
  float acc[8];
  // pass 1
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0.f;
  for (int t = tl; t < Tn; t += step) {
    float v[8];
    load8(wp + (int64_t)t * ld + o, v);
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = fmaf(v[i], v[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) atomicAdd(&norm2[o + i], acc[i]);
  __syncthreads();
  for (int i = tid; i < c; i += AT_THREADS) {
    const float nrm = fmaxf(sqrtf(norm2[i]), 1e-12f);
    norm2[i] = 1.0f / (nrm * nrm);
  }
  __syncthreads();
  // pass 2: per-token, per-head statistic
  {
    float inv[8];
#pragma unroll
    for (int i = 0; i < 8; i++) inv[i] = norm2[o + i];
    const float tmp = temp[hd];
    for (int t = tl; t < Tn; t += step) {
      float v[8], sx = 0.f;
      load8(wp + (int64_t)t * ld + o, v);
#pragma unroll
      for (int i = 0; i < 8; i++) sx = fmaf(v[i] * v[i], inv[i], sx);
      atomicAdd(&pi[t * heads + hd], sx * tmp);
    }
  }
  __syncthreads();
  for (int t = tid; t < Tn; t += AT_THREADS) {  // softmax across the heads of one token
    float mx = -INFINITY, se = 0.f;
    for (int h = 0; h < heads; h++) mx = fmaxf(mx, pi[t * heads + h]);
    for (int h = 0; h < heads; h++) se += expf(pi[t * heads + h] - mx);
    for (int h = 0; h < heads; h++) {
      const float pv = expf(pi[t * heads + h] - mx) / se;
      pi[t * heads + h] = pv;
      atomicAdd(&pisum[h], pv);
    }
  }
  __syncthreads();
  // pass 3
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0.f;
  for (int t = tl; t < Tn; t += step) {
    float v[8];
    load8(wp + (int64_t)t * ld + o, v);
    const float pv = pi[t * heads + hd];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = fmaf(pv * v[i], v[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) atomicAdd(&dots[o + i], acc[i]);
  __syncthreads();
  // pass 4
  {
    const float ps = 1.0f / (pisum[hd] + 1e-8f);
    float att[8];
#pragma unroll
    for (int i = 0; i < 8; i++) att[i] = -1.0f / (1.0f + dots[o + i] * ps);
    for (int t = tl; t < Tn; t += step) {
      float v[8];
      load8(wp + (int64_t)t * ld + o, v);
      const float pv = pi[t * heads + hd];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = v[i] * pv * att[i];
      store8(op + (int64_t)t * out_ld + o, v);
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
    YAD_LAUNCH(tssa_kernel<T>, qkv->n * heads, 256, smem, st, (const T*)qkv->ptr, qkv->ld, Tn, c, heads, temps, (T*)out->ptr, out->ld, Tout,
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
  const float scale = 1.0f / sqrtf((float)HD);
  if (dtype == YAD_BF16) {  // tensor-core path
    static int v2 = -1;  // YAD_MHA_V2=0 keeps the 64-query CTA / 16 rows per warp kernel
    if (v2 < 0) { const char* ev = getenv("YAD_MHA_V2"); v2 = (ev && ev[0] == '0') ? 0 : 1; }
    if (v2 && Tn > FA_Q) {  // (a single 64-query CTA per (image, head) covers short sequences with less padding)
      dim3 g3((Tn + FA2_Q - 1) / FA2_Q, qkv->n * heads);
      YAD_LAUNCH(mha_mma2_kernel, g3, 128, 0, st, (const bf16*)qkv->ptr, qkv->ld, Tn, c, heads, (bf16*)out->ptr, out->ld, scale * 1.44269504088896340736f);
      YAD_LAUNCH_CHECK("mha");
      return 0;
    }
    dim3 g2((Tn + FA_Q - 1) / FA_Q, qkv->n * heads);
    YAD_LAUNCH(mha_mma_kernel, g2, 128, 0, st, (const bf16*)qkv->ptr, qkv->ld, Tn, c, heads, (bf16*)out->ptr, out->ld, scale * 1.44269504088896340736f);
    YAD_LAUNCH_CHECK("mha");
    return 0;
  }
  dim3 grid((Tn + QB - 1) / QB, qkv->n * heads);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(mha_kernel<T>, grid, QB, 0, st, (const T*)qkv->ptr, qkv->ld, Tn, c, heads, (T*)out->ptr, out->ld, scale);)
  YAD_LAUNCH_CHECK("mha");
  return 0;
}

int yad_attention_tssa(const yad_tensor* w, const float* temp, int heads, const yad_tensor* out, int dtype, void* stream) {
  YAD_CHECK(w && out && temp && w->ptr && out->ptr, "attention_tssa: null argument");
  const int c = w->c, Tn = w->h * w->w;
  YAD_CHECK(out->n == w->n && out->h == w->h && out->w == w->w && out->c == c, "attention_tssa: shape mismatch");
  YAD_CHECK(c % 8 == 0 && w->ld % 8 == 0 && out->ld % 8 == 0, "attention_tssa: channels / strides must be multiples of 8");
  YAD_CHECK(heads >= 1 && c % heads == 0 && (c / heads) % 8 == 0, "attention_tssa: head_dim = %d / %d must be a multiple of 8", c, heads);
  YAD_CHECK(AT_THREADS % (c / 8) == 0, "attention_tssa: c / 8 = %d must divide %d", c / 8, AT_THREADS);
  if (w->n == 0 || Tn == 0) return 0;
  const size_t smem = sizeof(float) * ((size_t)2 * c + heads + (size_t)Tn * heads);
  YAD_CHECK(smem <= 200 * 1024, "attention_tssa: %d tokens x %d heads need %zu bytes of shared memory (limit 200 KB)", Tn, heads, smem);
  cudaStream_t st = (cudaStream_t)stream;
  YAD_DISPATCH_DTYPE(dtype, {
    if (smem > 48 * 1024) cudaFuncSetAttribute(attention_tssa_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    YAD_LAUNCH(attention_tssa_kernel<T>, w->n, AT_THREADS, smem, st, (const T*)w->ptr, (int64_t)w->ld, Tn, c, heads, temp, (T*)out->ptr,
               (int64_t)out->ld);
  })
  YAD_LAUNCH_CHECK("attention_tssa");
  return 0;
}

}  // extern "C"
