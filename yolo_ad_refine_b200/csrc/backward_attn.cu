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
  pdl_sync();
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
  pdl_sync();
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
  pdl_sync();
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


// ---- tensor-core MHA backward (bf16): flash-attention style, mma.sync.m16n8k16 (bf16 -> fp32), nothing of size T x T is materialised -----------
//   kernel 1 (dq):  CTA = 64 query rows (4 warps x 16).  Pass 0 streams the keys once for the row statistics lse2 = log2(sum exp(s)) (in the
//                   exp2 domain) and D = dO . O; pass 1 streams K, V again: S = Q K^T, dP = dO V^T, dS = P o (dP - D), dQ += dS K.
//   kernel 2 (dkv): CTA = 64 key rows.  Streams Q, dO tiles: S^T = K Q^T, P^T, dV += P^T dO, dP^T = V dO^T, dS^T, dK += dS^T Q.
//   Operand fragments follow the forward kernel (attention.cu): row-major tiles in shared memory with a conflict-free pitch, B fragments
//   through ldmatrix (x2 for [n][k] tiles, x2.trans for [k][n] tiles), P / dS re-packed from accumulators to A fragments in registers.
constexpr int BA_T = 64, BA_PITCH = 72;
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// A fragments (16 rows x 64 cols) of a row-major global matrix: rows r0 = row0 + g, r1 = r0 + 8 (zero when >= Tn)
__device__ __forceinline__ void load_afrag(const bf16* base, int64_t ld, int row0, int Tn, int g, int q4, uint32_t (&a)[4][4]) {
  const int r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
  for (int kt = 0; kt < 4; kt++) {
    const int col = kt * 16 + 2 * q4;
    a[kt][0] = r0 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r0 * ld + col) : 0u;
    a[kt][1] = r1 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r1 * ld + col) : 0u;
    a[kt][2] = r0 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r0 * ld + col + 8) : 0u;
    a[kt][3] = r1 < Tn ? *reinterpret_cast<const uint32_t*>(base + (int64_t)r1 * ld + col + 8) : 0u;
  }
}
// stage a 64 x 64 tile (rows row0.., zero beyond Tn) of a row-major global matrix into shared memory [64][BA_PITCH]
__device__ __forceinline__ void stage_tile(const bf16* src, int64_t ld, int row0, int Tn, bf16* dst, int tid) {
  for (int ch = tid; ch < BA_T * 8; ch += 128) {
    const int row = ch >> 3, cc = (ch & 7) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + row < Tn) v = *reinterpret_cast<const uint4*>(src + (int64_t)(row0 + row) * ld + cc);
    *reinterpret_cast<uint4*>(dst + row * BA_PITCH + cc) = v;
  }
}
// asynchronous variant (cp.async, 16 B per request, zero fill beyond Tn): the next tile streams in while the tensor cores work on the current one
__device__ __forceinline__ void stage_tile_async(const bf16* src, int64_t ld, int row0, int Tn, uint32_t dst_addr, int tid) {
  for (int ch = tid; ch < BA_T * 8; ch += 128) {
    const int row = ch >> 3, cc = (ch & 7) * 8;
    const bool ok = row0 + row < Tn;
    const bf16* g = src + (int64_t)(ok ? row0 + row : Tn - 1) * ld + cc;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_addr + (uint32_t)((row * BA_PITCH + cc) * 2)), "l"(g), "r"(ok ? 16 : 0)
                 : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// acc[j] (16 x 8 tile j of a 16 x 64 product) = A(16 x 64 frags) * tile^T, tile = [64 n][64 k] row-major in shared memory
__device__ __forceinline__ void mm_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile_addr, int lane) {
#pragma unroll
  for (int j = 0; j < 8; j++) {
#pragma unroll
    for (int i = 0; i < 4; i++) acc[j][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 4; kt++) {
      uint32_t b0, b1;
      ldsm2(b0, b1, tile_addr + (uint32_t)(((j * 8 + (lane & 7)) * BA_PITCH + kt * 16 + ((lane >> 3) & 1) * 8) * 2));
      mma16816(acc[j], a[kt], b0, b1);
    }
  }
}
// acc[j] += A(16 x 64 frags over the tile's rows) * tile, tile = [64 k][64 n] row-major in shared memory
__device__ __forceinline__ void mm_nn_acc(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile_addr, int lane) {
#pragma unroll
  for (int kt = 0; kt < 4; kt++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint32_t b0, b1;
      ldsm2t(b0, b1, tile_addr + (uint32_t)(((kt * 16 + (lane & 15)) * BA_PITCH + j * 8) * 2));
      mma16816(acc[j], a[kt], b0, b1);
    }
  }
}

__global__ void __launch_bounds__(128) mha_bwd_dq_mma_kernel(const bf16* __restrict__ qkv, int64_t ld, int Tn, int c, int heads,
                                                             const bf16* __restrict__ out, int64_t out_ld, const bf16* __restrict__ dout,
                                                             int64_t dout_ld, float scale, float scale_log2e, bf16* __restrict__ dqkv, int64_t dld,
                                                             float* __restrict__ lse_d) {
  pdl_sync();
  __shared__ __align__(16) bf16 Ks[2][BA_T * BA_PITCH];
  __shared__ __align__(16) bf16 Vs[2][BA_T * BA_PITCH];
  const int n = blockIdx.y / heads, h = blockIdx.y % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
  const bf16* base = qkv + (int64_t)n * Tn * ld + h * HD;
  const bf16* obase = out + (int64_t)n * Tn * out_ld + h * HD;
  const bf16* gbase = dout + (int64_t)n * Tn * dout_ld + h * HD;
  const int q0 = blockIdx.x * BA_T + warp * 16, r0 = q0 + g, r1 = r0 + 8;
  uint32_t qa[4][4], ga[4][4];
  load_afrag(base, ld, q0, Tn, g, q4, qa);
  load_afrag(gbase, dout_ld, q0, Tn, g, q4, ga);
  // D = dO . O for rows r0, r1 (each lane holds 16 of the 64 columns of both rows)
  float D0 = 0.f, D1 = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; kt++) {
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      const int col = kt * 16 + 2 * q4 + hh * 8;
      if (r0 < Tn) {
        const __nv_bfloat162 o2 = *reinterpret_cast<const __nv_bfloat162*>(obase + (int64_t)r0 * out_ld + col);
        const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(&ga[kt][hh * 2]);
        D0 += __bfloat162float(o2.x) * __bfloat162float(g2.x) + __bfloat162float(o2.y) * __bfloat162float(g2.y);
      }
      if (r1 < Tn) {
        const __nv_bfloat162 o2 = *reinterpret_cast<const __nv_bfloat162*>(obase + (int64_t)r1 * out_ld + col);
        const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(&ga[kt][hh * 2 + 1]);
        D1 += __bfloat162float(o2.x) * __bfloat162float(g2.x) + __bfloat162float(o2.y) * __bfloat162float(g2.y);
      }
    }
  }
  D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
  D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
  const uint32_t ks_base = (uint32_t)__cvta_generic_to_shared(&Ks[0][0]), vs_base = (uint32_t)__cvta_generic_to_shared(&Vs[0][0]);
  constexpr uint32_t BUF = BA_T * BA_PITCH * 2;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f, lse0 = INFINITY, lse1 = INFINITY;
  float dq[8][4];
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) dq[j][i] = 0.f;
  const int nt = (Tn + BA_T - 1) / BA_T;  // the two passes form one sequence of 2 * nt tiles through a double buffer
  stage_tile_async(base + c, ld, 0, Tn, ks_base, tid);
  cp_async_commit();
  for (int it = 0; it < 2 * nt; it++) {
    const int pass = it >= nt, k0 = (pass ? it - nt : it) * BA_T, buf = it & 1;
    cp_async_wait_all();
    __syncthreads();  // tile `it` has landed; everybody is done with tile it - 1, whose buffer is refilled below
    if (it + 1 < 2 * nt) {
      const int np = it + 1 >= nt, nk0 = (np ? it + 1 - nt : it + 1) * BA_T;
      stage_tile_async(base + c, ld, nk0, Tn, ks_base + (buf ^ 1) * BUF, tid);
      if (np) stage_tile_async(base + 2 * c, ld, nk0, Tn, vs_base + (buf ^ 1) * BUF, tid);
      cp_async_commit();
    }
    const uint32_t ks_addr = ks_base + buf * BUF, vs_addr = vs_base + buf * BUF;
    {
      float s[8][4];
      mm_nt(s, qa, ks_addr, lane);
      if (pass == 0) {
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
        l0 *= exp2f((m0 - mn0) * scale_log2e);
        l1 *= exp2f((m1 - mn1) * scale_log2e);
        const float mb0 = mn0 * scale_log2e, mb1 = mn1 * scale_log2e;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          l0 += exp2f(fmaf(s[j][0], scale_log2e, -mb0)) + exp2f(fmaf(s[j][1], scale_log2e, -mb0));
          l1 += exp2f(fmaf(s[j][2], scale_log2e, -mb1)) + exp2f(fmaf(s[j][3], scale_log2e, -mb1));
        }
        m0 = mn0; m1 = mn1;
      } else {
        float dp[8][4];
        mm_nt(dp, ga, vs_addr, lane);
        uint32_t dsa[4][4];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int key = k0 + j * 8 + 2 * q4;
          const bool v0 = key < Tn, v1 = key + 1 < Tn;
          const float p0 = v0 ? exp2f(fmaf(s[j][0], scale_log2e, -lse0)) : 0.f, p1 = v1 ? exp2f(fmaf(s[j][1], scale_log2e, -lse0)) : 0.f;
          const float p2 = v0 ? exp2f(fmaf(s[j][2], scale_log2e, -lse1)) : 0.f, p3 = v1 ? exp2f(fmaf(s[j][3], scale_log2e, -lse1)) : 0.f;
          dsa[j >> 1][(j & 1) * 2 + 0] = pack2(p0 * (dp[j][0] - D0), p1 * (dp[j][1] - D0));
          dsa[j >> 1][(j & 1) * 2 + 1] = pack2(p2 * (dp[j][2] - D1), p3 * (dp[j][3] - D1));
        }
        mm_nn_acc(dq, dsa, ks_addr, lane);
      }
    }
    if (it == nt - 1) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      lse0 = r0 < Tn ? m0 * scale_log2e + log2f(l0) : INFINITY;
      lse1 = r1 < Tn ? m1 * scale_log2e + log2f(l1) : INFINITY;
      if (q4 == 0) {
        if (r0 < Tn) { lse_d[((int64_t)blockIdx.y * Tn + r0) * 2] = lse0; lse_d[((int64_t)blockIdx.y * Tn + r0) * 2 + 1] = D0; }
        if (r1 < Tn) { lse_d[((int64_t)blockIdx.y * Tn + r1) * 2] = lse1; lse_d[((int64_t)blockIdx.y * Tn + r1) * 2 + 1] = D1; }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int col = h * HD + j * 8 + 2 * q4;
    if (r0 < Tn) *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)n * Tn + r0) * dld + col) = pack2(dq[j][0] * scale, dq[j][1] * scale);
    if (r1 < Tn) *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)n * Tn + r1) * dld + col) = pack2(dq[j][2] * scale, dq[j][3] * scale);
  }
}

__global__ void __launch_bounds__(128) mha_bwd_dkv_mma_kernel(const bf16* __restrict__ qkv, int64_t ld, int Tn, int c, int heads,
                                                              const bf16* __restrict__ dout, int64_t dout_ld, float scale, float scale_log2e,
                                                              const float* __restrict__ lse_d, bf16* __restrict__ dqkv, int64_t dld) {
  pdl_sync();
  __shared__ __align__(16) bf16 Qs[2][BA_T * BA_PITCH];
  __shared__ __align__(16) bf16 Gs[2][BA_T * BA_PITCH];
  __shared__ float Ls[2][BA_T], Ds[2][BA_T];
  const int n = blockIdx.y / heads, h = blockIdx.y % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
  const bf16* base = qkv + (int64_t)n * Tn * ld + h * HD;
  const bf16* gbase = dout + (int64_t)n * Tn * dout_ld + h * HD;
  const int k0 = blockIdx.x * BA_T + warp * 16, r0 = k0 + g, r1 = r0 + 8;
  uint32_t ka[4][4], va[4][4];
  load_afrag(base + c, ld, k0, Tn, g, q4, ka);
  load_afrag(base + 2 * c, ld, k0, Tn, g, q4, va);
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) { dk[j][i] = 0.f; dv[j][i] = 0.f; }
  const uint32_t qs_base = (uint32_t)__cvta_generic_to_shared(&Qs[0][0]), gs_base = (uint32_t)__cvta_generic_to_shared(&Gs[0][0]);
  constexpr uint32_t BUF = BA_T * BA_PITCH * 2;
  auto stage = [&](int q0, int buf) {
    stage_tile_async(base, ld, q0, Tn, qs_base + buf * BUF, tid);
    stage_tile_async(gbase, dout_ld, q0, Tn, gs_base + buf * BUF, tid);
    cp_async_commit();
    if (tid < BA_T) {
      const int t = q0 + tid;
      Ls[buf][tid] = t < Tn ? lse_d[((int64_t)blockIdx.y * Tn + t) * 2] : INFINITY;  // invalid query column -> P = 0
      Ds[buf][tid] = t < Tn ? lse_d[((int64_t)blockIdx.y * Tn + t) * 2 + 1] : 0.f;
    }
  };
  stage(0, 0);
  for (int q0 = 0, it = 0; q0 < Tn; q0 += BA_T, it++) {
    const int buf = it & 1;
    cp_async_wait_all();
    __syncthreads();  // tile `it` has landed; the other buffer is free
    if (q0 + BA_T < Tn) stage(q0 + BA_T, buf ^ 1);
    const uint32_t qs_addr = qs_base + buf * BUF, gs_addr = gs_base + buf * BUF;
    const float* Lb = Ls[buf];
    const float* Db = Ds[buf];
    float st[8][4], dpt[8][4];
    mm_nt(st, ka, qs_addr, lane);   // S^T = K Q^T
    mm_nt(dpt, va, gs_addr, lane);  // dP^T = V dO^T
    uint32_t pta[4][4], dsta[4][4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int qc = j * 8 + 2 * q4;
      const float L0 = Lb[qc], L1 = Lb[qc + 1], E0 = Db[qc], E1 = Db[qc + 1];
      const float p0 = exp2f(fmaf(st[j][0], scale_log2e, -L0)), p1 = exp2f(fmaf(st[j][1], scale_log2e, -L1));
      const float p2 = exp2f(fmaf(st[j][2], scale_log2e, -L0)), p3 = exp2f(fmaf(st[j][3], scale_log2e, -L1));
      pta[j >> 1][(j & 1) * 2 + 0] = pack2(p0, p1);
      pta[j >> 1][(j & 1) * 2 + 1] = pack2(p2, p3);
      dsta[j >> 1][(j & 1) * 2 + 0] = pack2(p0 * (dpt[j][0] - E0), p1 * (dpt[j][1] - E1));
      dsta[j >> 1][(j & 1) * 2 + 1] = pack2(p2 * (dpt[j][2] - E0), p3 * (dpt[j][3] - E1));
    }
    mm_nn_acc(dv, pta, gs_addr, lane);   // dV += P^T dO
    mm_nn_acc(dk, dsta, qs_addr, lane);  // dK += dS^T Q
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int col = h * HD + j * 8 + 2 * q4;
    if (r0 < Tn) {
      *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)n * Tn + r0) * dld + c + col) = pack2(dk[j][0] * scale, dk[j][1] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)n * Tn + r0) * dld + 2 * c + col) = pack2(dv[j][0], dv[j][1]);
    }
    if (r1 < Tn) {
      *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)n * Tn + r1) * dld + c + col) = pack2(dk[j][2] * scale, dk[j][3] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)n * Tn + r1) * dld + 2 * c + col) = pack2(dv[j][2], dv[j][3]);
    }
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
    YAD_LAUNCH(tssa_bwd_kernel<T>, qkv->n * heads, 256, smem, st, (const T*)qkv->ptr, qkv->ld, Tn, c, heads, temps, (const T*)dout->ptr, dout->ld, Tout,
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
  if (dtype == YAD_BF16) {  // tensor-core path
    dim3 g2((Tn + BA_T - 1) / BA_T, qkv->n * heads);
    const float sl2 = scale * 1.44269504088896340736f;
    YAD_LAUNCH(mha_bwd_dq_mma_kernel, g2, 128, 0, st, (const bf16*)qkv->ptr, qkv->ld, Tn, c, heads, (const bf16*)out->ptr, out->ld, (const bf16*)dout->ptr,
                                              dout->ld, scale, sl2, (bf16*)dqkv->ptr, dqkv->ld, lse_d);
    YAD_LAUNCH(mha_bwd_dkv_mma_kernel, g2, 128, 0, st, (const bf16*)qkv->ptr, qkv->ld, Tn, c, heads, (const bf16*)dout->ptr, dout->ld, scale, sl2, lse_d,
                                               (bf16*)dqkv->ptr, dqkv->ld);
    YAD_LAUNCH_CHECK("mha_bwd");
    return 0;
  }
  dim3 grid((Tn + RB - 1) / RB, qkv->n * heads);
  YAD_DISPATCH_DTYPE(dtype, {
    YAD_LAUNCH(mha_bwd_dq_kernel<T>, grid, 128, 0, st, (const T*)qkv->ptr, qkv->ld, Tn, c, heads, (const T*)out->ptr, out->ld, (const T*)dout->ptr, dout->ld,
                                              scale, (T*)dqkv->ptr, dqkv->ld, lse_d);
    YAD_LAUNCH(mha_bwd_dkv_kernel<T>, grid, 128, 0, st, (const T*)qkv->ptr, qkv->ld, Tn, c, heads, (const T*)dout->ptr, dout->ld, scale, lse_d,
                                               (T*)dqkv->ptr, dqkv->ld);
  })
  YAD_LAUNCH_CHECK("mha_bwd");
  return 0;
}

}  // extern "C"
