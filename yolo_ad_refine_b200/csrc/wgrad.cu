// Weight-gradient kernels of the training path and the deformable-convolution column helpers.
//   conv wgrad : dW[co][tap][ci] += sum over output pixels m of dy[m][co] * x[pix(m, tap)][ci]   (GEMM with K = n*ho*wo, split-K + fp32 atomics)
//                bf16: mma.sync.m16n8k16 with both operands read through transposed ldmatrix (the reduction axis is the pixel axis,
//                which is the strided one in NHWC); fp32: SIMT reference twin.
//   dwconv wgrad, deform_col (modulated bilinear sampling -> column tensor) and its backward.
#include "common.cuh"

namespace {

struct WgGeom {
  int n, hi, wi, cin, ho, wo, cout;
  int kh, kw, stride, pad_h, pad_w;
  int x_ld, dy_ld;
  int64_t M;        // n*ho*wo
  int64_t per;      // pixels per split
};

__device__ __forceinline__ int64_t in_pixel(const WgGeom& g, int64_t m, int ky, int kx) {
  const int ox = (int)(m % g.wo);
  const int oy = (int)((m / g.wo) % g.ho);
  const int img = (int)(m / ((int64_t)g.wo * g.ho));
  const int iy = oy * g.stride - g.pad_h + ky, ix = ox * g.stride - g.pad_w + kx;
  if (iy < 0 || iy >= g.hi || ix < 0 || ix >= g.wi) return -1;
  return ((int64_t)img * g.hi + iy) * g.wi + ix;
}

// ---- SIMT twin (fp32 or bf16 storage, fp32 FMA): 64 co x 64 ci tile, 32 pixels per step, 4x4 outputs per thread ----------------
constexpr int SB_P = 32, SB_C = 64;
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy, WgGeom g, float* __restrict__ dw) {
  pdl_sync();
  __shared__ __align__(16) float sA[SB_P][SB_C + 4];  // dy[pixel][co]
  __shared__ __align__(16) float sB[SB_P][SB_C + 4];  // x[pixel][ci]
  const int ntaps = g.kh * g.kw;
  const int tap = blockIdx.x % ntaps, ci0 = (blockIdx.x / ntaps) * SB_C, co0 = blockIdx.y * SB_C;
  const int ky = tap / g.kw, kx = tap % g.kw;
  const int64_t m0 = (int64_t)blockIdx.z * g.per, m1 = min(g.M, m0 + g.per);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int lp = tid >> 3, lo = (tid & 7) * 8;  // loader: pixel lp (0..31), channel octet lo
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  for (int64_t mb = m0; mb < m1; mb += SB_P) {
    const int64_t m = mb + lp;
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = 0.f; b[i] = 0.f; }
    if (m < m1) {
      if (co0 + lo < g.cout) load8(dy + m * g.dy_ld + co0 + lo, a);
      const int64_t ip = in_pixel(g, m, ky, kx);
      if (ip >= 0 && ci0 + lo < g.cin) load8(x + ip * g.x_ld + ci0 + lo, b);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; i++) { sA[lp][lo + i] = a[i]; sB[lp][lo + i] = b[i]; }
    __syncthreads();
#pragma unroll 8
    for (int p = 0; p < SB_P; p++) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[p][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sB[p][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int co = co0 + ty * 4 + i;
    if (co >= g.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < g.cin) atomicAdd(&dw[((int64_t)co * ntaps + tap) * g.cin + ci], acc[i][j]);
    }
  }
}

// ---- tensor-core kernel (bf16): 64 co x 64 ci tile per CTA, 64 pixels per step; 4 warps x (16 co x 64 ci) ----------------------------
constexpr int WG_P = 64, WG_PITCH = 72;
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

__global__ void __launch_bounds__(128) conv_wgrad_mma_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, WgGeom g, float* __restrict__ dw) {
  pdl_sync();
  __shared__ __align__(16) bf16 sA[2][WG_P * WG_PITCH];  // dy[pixel][co]
  __shared__ __align__(16) bf16 sB[2][WG_P * WG_PITCH];  // x[pixel][ci]
  const int ntaps = g.kh * g.kw;
  const int tap = blockIdx.x % ntaps, ci0 = (blockIdx.x / ntaps) * 64, co0 = blockIdx.y * 64;
  const int ky = tap / g.kw, kx = tap % g.kw;
  const int64_t m0 = (int64_t)blockIdx.z * g.per, m1 = min(g.M, m0 + g.per);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) acc[j][i] = 0.f;
  const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(&sA[0][0]), b_base = (uint32_t)__cvta_generic_to_shared(&sB[0][0]);
  constexpr uint32_t BUF_BYTES = WG_P * WG_PITCH * 2;
  uint4 ra[4], rb[4];  // register-staged next tile: 64 pixels x 8 chunks / 128 threads = 4 chunks per thread per operand
  auto fetch = [&](int64_t mb) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int ch = tid + q * 128, row = ch >> 3, cc = (ch & 7) * 8;
      const int64_t m = mb + row;
      ra[q] = make_uint4(0u, 0u, 0u, 0u);
      rb[q] = ra[q];
      if (m < m1) {
        if (co0 + cc < g.cout) ra[q] = *reinterpret_cast<const uint4*>(dy + m * g.dy_ld + co0 + cc);
        const int64_t ip = in_pixel(g, m, ky, kx);
        if (ip >= 0 && ci0 + cc < g.cin) rb[q] = *reinterpret_cast<const uint4*>(x + ip * g.x_ld + ci0 + cc);
      }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int ch = tid + q * 128, row = ch >> 3, cc = (ch & 7) * 8;
      *reinterpret_cast<uint4*>(&sA[buf][row * WG_PITCH + cc]) = ra[q];
      *reinterpret_cast<uint4*>(&sB[buf][row * WG_PITCH + cc]) = rb[q];
    }
  };
  int buf = 0;
  if (m0 < m1) {
    fetch(m0);
    stash(0);
  }
  __syncthreads();
  for (int64_t mb = m0; mb < m1; mb += WG_P) {
    const bool more = mb + WG_P < m1;
    if (more) fetch(mb + WG_P);  // global loads in flight while the tensor cores work on the current tile
    const uint32_t a_addr = a_base + buf * BUF_BYTES, b_addr = b_base + buf * BUF_BYTES;
#pragma unroll
    for (int ks = 0; ks < WG_P / 16; ks++) {
      uint32_t a[4];
      {
        const int mat = lane >> 3, r = lane & 7, mi = mat & 1, kj = mat >> 1;
        ldsm_x4_trans(a, a_addr + (uint32_t)(((ks * 16 + kj * 8 + r) * WG_PITCH + warp * 16 + mi * 8) * 2));
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, b_addr + (uint32_t)(((ks * 16 + (lane & 15)) * WG_PITCH + j * 8) * 2));
        mma_bf16_16816(acc[j], a, b0, b1);
      }
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  const int gq = lane >> 2, q4 = lane & 3;
#pragma unroll
  for (int j = 0; j < 8; j++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int co = co0 + warp * 16 + gq + (i >> 1) * 8;
      const int ci = ci0 + j * 8 + 2 * q4 + (i & 1);
      if (co < g.cout && ci < g.cin) atomicAdd(&dw[((int64_t)co * ntaps + tap) * g.cin + ci], acc[j][i]);
    }
  }
}


// ---- small-channel 3x3 wgrad (cin, cout <= 32; the first layers and the head's narrow convs): HBM-bound, so every operand byte is read ONCE.
//   A persistent CTA walks tiles of 8 x 32 output pixels of one image: the input patch (with its 1-pixel halo; stride 1 or 2) and the dy tile
//   are staged in shared memory with 128-bit loads (zero-filled padding / ragged edges), then all 9 taps are accumulated from the same staged
//   data: warp (kg, ng) owns the 8-channel input group ng for all taps (9 n-tiles x MT m-tiles of mma.sync.m16n8k16 accumulators in
//   registers) and every KW-th 16-pixel k-step.  One shared-memory reduction + one global atomic per weight per CTA at the very end.
constexpr int WS_TH = 8, WS_TW = 32, WS_THREADS = 256;
__host__ __device__ constexpr int ws_pitch(int c) { return ((c >> 3) & 1) ? c + 16 : c + 8; }  // odd number of 16-byte slots per row: conflict-free ldmatrix

template <int MT>
__global__ void __launch_bounds__(WS_THREADS) conv_wgrad_small_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, WgGeom g,
                                                                          float* __restrict__ dw, int tiles_x, int tiles_y, int total_tiles) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t ws_smem[];
  const int s = g.stride;
  const int PH = (WS_TH - 1) * s + 3, PW = (WS_TW - 1) * s + 3;
  const int cpitch = ws_pitch(g.cin), dpitch = ws_pitch(MT * 16);
  // two (patch, dy tile) buffer pairs: tile i + 1 streams in through cp.async while tile i feeds the tensor cores
  const uint32_t patch_elems = (uint32_t)PH * PW * cpitch, dy_elems = (uint32_t)WS_TH * WS_TW * dpitch, pair_elems = patch_elems + dy_elems;
  bf16* patch = reinterpret_cast<bf16*>(ws_smem);
  bf16* dys = patch + patch_elems;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NW = g.cin >> 3, KW = 8 / NW;  // NW in {1, 2, 4}
  const int ng = warp % NW, kg = warp / NW;
  float acc[9][MT][4];
#pragma unroll
  for (int t = 0; t < 9; t++)
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int i = 0; i < 4; i++) acc[t][m][i] = 0.f;
  // both dy staging areas are zeroed once: channel columns >= cout (padding up to MT*16) are never written afterwards
  for (int b = 0; b < 2; b++)
    for (int i = tid; i < WS_TH * WS_TW * dpitch / 8; i += WS_THREADS) reinterpret_cast<uint4*>(dys + (size_t)b * pair_elems)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const uint32_t patch_s0 = (uint32_t)__cvta_generic_to_shared(patch), dys_s0 = (uint32_t)__cvta_generic_to_shared(dys);
  const int coct = g.cin >> 3, doct = g.cout >> 3;
  // tile-independent chunk geometry lives in registers (see conv_small.cu): per tile a slot costs two adds, the bounds test and one cp.async
  constexpr int MAXS = 9, MAXD = 4;  // 17 x 65 x 2 patch chunks / 256 threads; 256 x 4 dy chunks / 256 threads
  int sl_py[MAXS], sl_px[MAXS], sl_g[MAXS], dl_py[MAXD], dl_px[MAXD], dl_g[MAXD];
  uint32_t sl_s[MAXS], dl_s[MAXD];
  const int nchunks = PH * PW * coct, dchunks = WS_TH * WS_TW * doct;
#pragma unroll
  for (int j = 0; j < MAXS; j++) {
    const int ch = tid + j * WS_THREADS;
    const int oc = ch % coct, pp = ch / coct, px = pp % PW, py = pp / PW;
    sl_py[j] = ch < nchunks ? py : -100000;  // fails every bounds test
    sl_px[j] = px;
    sl_g[j] = (py * g.wi + px) * g.x_ld + oc * 8;
    sl_s[j] = (uint32_t)((pp * cpitch + oc * 8) * 2);
  }
#pragma unroll
  for (int j = 0; j < MAXD; j++) {
    const int ch = tid + j * WS_THREADS;
    const int oc = ch % doct, pp = ch / doct, px = pp % WS_TW, py = pp / WS_TW;
    dl_py[j] = ch < dchunks ? py : 100000;
    dl_px[j] = px;
    dl_g[j] = (py * g.wo + px) * g.dy_ld + oc * 8;
    dl_s[j] = (uint32_t)((pp * dpitch + oc * 8) * 2);
  }
  auto stage = [&](int tile, int buf) {  // zero fill (src-size 0) for the conv padding and ragged tile edges
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, img = tile / (tiles_x * tiles_y);
    const int oy0 = ty * WS_TH, ox0 = tx * WS_TW;
    const int iy0 = oy0 * s - g.pad_h, ix0 = ox0 * s - g.pad_w;
    const uint32_t pdst = patch_s0 + (uint32_t)buf * pair_elems * 2u, ddst = dys_s0 + (uint32_t)buf * pair_elems * 2u;
    const bf16* xb = x + (((int64_t)img * g.hi + iy0) * g.wi + ix0) * g.x_ld;
    const bf16* db = dy + (((int64_t)img * g.ho + oy0) * g.wo + ox0) * g.dy_ld;
#pragma unroll
    for (int j = 0; j < MAXS; j++) {
      if (j * WS_THREADS < nchunks && sl_py[j] > -100000) {
        const int iy = iy0 + sl_py[j], ix = ix0 + sl_px[j];
        const bool ok = (unsigned)iy < (unsigned)g.hi && (unsigned)ix < (unsigned)g.wi;
        const bf16* src = ok ? xb + sl_g[j] : x;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(pdst + sl_s[j]), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
    }
#pragma unroll
    for (int j = 0; j < MAXD; j++) {
      if (j * WS_THREADS < dchunks && dl_py[j] < 100000) {
        const bool ok = oy0 + dl_py[j] < g.ho && ox0 + dl_px[j] < g.wo;
        const bf16* src = ok ? db + dl_g[j] : dy;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ddst + dl_s[j]), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < total_tiles) stage(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y;
    const int oy0 = ty * WS_TH, ox0 = tx * WS_TW;
    const int buf = it & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // this tile has landed; everybody is done with the previous one (the other buffer pair)
    if (tile + (int)gridDim.x < total_tiles) stage(tile + gridDim.x, buf ^ 1);
    const uint32_t patch_s = patch_s0 + (uint32_t)buf * pair_elems * 2u, dys_s = dys_s0 + (uint32_t)buf * pair_elems * 2u;
    for (int kk = kg; kk < WS_TH * WS_TW / 16; kk += KW) {
      const int py = kk / (WS_TW / 16), px0 = (kk % (WS_TW / 16)) * 16;
      if (oy0 + py >= g.ho || ox0 + px0 >= g.wo) continue;  // warp-uniform: this k-step holds only zero dy
      uint32_t a[MT][4];
      {
        const int mat = lane >> 3, r = lane & 7, mi = mat & 1, kj = mat >> 1;
#pragma unroll
        for (int m = 0; m < MT; m++) ldsm_x4_trans(a[m], dys_s + (uint32_t)(((py * WS_TW + px0 + kj * 8 + r) * dpitch + m * 16 + mi * 8) * 2));
      }
      const int r16 = lane & 15;
#pragma unroll
      for (int t = 0; t < 9; t++) {
        const int ky = t / 3, kx = t % 3;
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, patch_s + (uint32_t)((((py * s + ky) * PW + (px0 + r16) * s + kx) * cpitch + ng * 8) * 2));
#pragma unroll
        for (int m = 0; m < MT; m++) mma_bf16_16816(acc[t][m], a[m], b0, b1);
      }
    }
  }
  // cross-warp reduction in shared memory (reusing the patch area), then one global atomic per weight
  __syncthreads();
  float* red = reinterpret_cast<float*>(ws_smem);
  const int nw = MT * 16 * 9 * g.cin;
  for (int i = tid; i < nw; i += WS_THREADS) red[i] = 0.f;
  __syncthreads();
  const int gq = lane >> 2, q4 = lane & 3;
#pragma unroll
  for (int t = 0; t < 9; t++)
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int co = m * 16 + gq + (i >> 1) * 8, ci = ng * 8 + 2 * q4 + (i & 1);
        atomicAdd(&red[(co * 9 + t) * g.cin + ci], acc[t][m][i]);
      }
  __syncthreads();
  for (int i = tid; i < nw; i += WS_THREADS) {
    const int co = i / (9 * g.cin);
    if (co < g.cout && red[i] != 0.f) atomicAdd(&dw[i], red[i]);
  }
}

// ---- depthwise wgrad: dW[tap][c] += sum_p dy[p][c] * x[p + tap][c].  A thread owns (8-channel octet, kernel row ky, pixel lane) and keeps the k
//      taps of its row in registers while it strides over the CTA's pixel chunk: dy is read once per kernel row, the k x k neighbourhood of x
//      comes from L1/L2; one global atomic per (tap, channel) per thread at the end (warp-shuffle pre-reduction when lanes share an octet).
template <typename T, int K>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(yad_tensor x, yad_tensor dy, float* __restrict__ dw) {
  pdl_sync();
  const int c = x.c, oct = c >> 3, r = K >> 1;
  const int per_pix = oct * K, lanes = blockDim.x / per_pix;
  const int tid = threadIdx.x;
  const int o = (tid % oct) * 8, ky = (tid / oct) % K, pl = tid / per_pix;
  const int64_t npix = (int64_t)x.n * x.h * x.w;
  const int64_t per = (npix + gridDim.x - 1) / gridDim.x, p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  float acc[K][8];
#pragma unroll
  for (int t = 0; t < K; t++)
#pragma unroll
    for (int i = 0; i < 8; i++) acc[t][i] = 0.f;
  if (pl < lanes) {
    for (int64_t p = p0 + pl; p < p1; p += lanes) {
      const int px = (int)(p % x.w), py = (int)((p / x.w) % x.h);
      const int iy = py + ky - r;
      if (iy < 0 || iy >= x.h) continue;
      float gg[8];
      load8(reinterpret_cast<const T*>(dy.ptr) + p * dy.ld + o, gg);
      const T* xrow = reinterpret_cast<const T*>(x.ptr) + (p + (int64_t)(ky - r) * x.w - r) * x.ld + o;
#pragma unroll
      for (int t = 0; t < K; t++) {
        const int ix = px + t - r;
        if (ix < 0 || ix >= x.w) continue;
        float v[8];
        load8(xrow + (int64_t)t * x.ld, v);
#pragma unroll
        for (int i = 0; i < 8; i++) acc[t][i] = fmaf(v[i], gg[i], acc[t][i]);
      }
    }
  }
  // pixel lanes of the CTA combine in shared memory, then one global atomic per (tap, channel) per CTA
  extern __shared__ float dws[];  // [K*K][c]
  for (int i = tid; i < K * K * c; i += blockDim.x) dws[i] = 0.f;
  __syncthreads();
  if (pl < lanes) {
#pragma unroll
    for (int t = 0; t < K; t++)
#pragma unroll
      for (int i = 0; i < 8; i++) atomicAdd(&dws[(ky * K + t) * c + o + i], acc[t][i]);
  }
  __syncthreads();
  for (int i = tid; i < K * K * c; i += blockDim.x) atomicAdd(&dw[i], dws[i]);
}

// ---- modulated deformable sampling (DCNv2, 3x3, stride 1, pad 1, one offset group) ---------------------------------------------------------
struct Sample {
  bool valid;
  int y0, x0;
  float ly, lx, mk;
};
template <typename T>
__device__ __forceinline__ Sample dcn_sample(const T* __restrict__ om, int tap, int oy, int ox, int h, int w) {
  Sample s;
  const int ky = tap / 3, kx = tap % 3;
  const float dy = ld1(om + 2 * tap), dx = ld1(om + 2 * tap + 1);
  s.mk = sigmoidf_(ld1(om + 18 + tap));
  const float py = (float)(oy - 1 + ky) + dy, px = (float)(ox - 1 + kx) + dx;
  s.valid = py > -1.f && px > -1.f && py < (float)h && px < (float)w;
  const float fy = floorf(py), fx = floorf(px);
  s.y0 = (int)fy; s.x0 = (int)fx;
  s.ly = py - fy; s.lx = px - fx;
  return s;
}

// pixel q of the tile-major order (8 x 8 tiles, so that the gathers of a CTA share their neighbourhood in L1) -> (n, oy, ox); false = padding
__device__ __forceinline__ bool tile_pixel(uint32_t q, int h, int w, int tiles_x, int tiles_per_img, int* n, int* oy, int* ox) {
  const uint32_t tile = q >> 6, r = q & 63;
  const uint32_t img = tile / (uint32_t)tiles_per_img, t = tile - img * (uint32_t)tiles_per_img;
  const uint32_t ty = t / (uint32_t)tiles_x, tx = t - ty * (uint32_t)tiles_x;
  *n = (int)img;
  *oy = (int)(ty * 8 + (r >> 3));
  *ox = (int)(tx * 8 + (r & 7));
  return *oy < h && *ox < w;
}

// col[n, y, x, tap*C + c] = mask * bilinear(x)   (same arithmetic as the deformable mode of yad_conv2d).
// thread = (octet, tap, pixel lane): the block strides over pixels in tile-major order; all index arithmetic is 32-bit.
template <typename T>
__global__ void deform_col_kernel(yad_tensor x, const T* __restrict__ om, int om_ld, yad_tensor col, int tiles_x, int tiles_per_img, uint32_t npix_t) {
  pdl_sync();
  const int C = x.c, oct = C >> 3, per_pix = 9 * oct;
  const int lanes = blockDim.x / per_pix;
  const int o = (threadIdx.x % oct) * 8, tap = (threadIdx.x / oct) % 9, pl = threadIdx.x / per_pix;
  if (pl >= lanes) return;
  for (uint32_t q = blockIdx.x * lanes + pl; q < npix_t; q += gridDim.x * lanes) {
    int n, oy, ox;
    if (!tile_pixel(q, x.h, x.w, tiles_x, tiles_per_img, &n, &oy, &ox)) continue;
    const int64_t p = ((int64_t)n * x.h + oy) * x.w + ox;
    const Sample s = dcn_sample(om + p * om_ld, tap, oy, ox, x.h, x.w);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 0.f;
    if (s.valid) {
      const float wgt[4] = {(1.f - s.ly) * (1.f - s.lx), (1.f - s.ly) * s.lx, s.ly * (1.f - s.lx), s.ly * s.lx};
#pragma unroll
      for (int c4 = 0; c4 < 4; c4++) {
        const int yy = s.y0 + (c4 >> 1), xx = s.x0 + (c4 & 1);
        if (yy >= 0 && yy < x.h && xx >= 0 && xx < x.w) {
          float t[8];
          load8(reinterpret_cast<const T*>(x.ptr) + (((int64_t)n * x.h + yy) * x.w + xx) * x.ld + o, t);
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] += wgt[c4] * t[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] *= s.mk;
    }
    store8(reinterpret_cast<T*>(col.ptr) + p * col.ld + tap * C + o, v);
  }
}

// backward of deform_col: `oct` lanes (C/8, a power of two <= 32) cooperate on one (pixel, tap).
// dx_f: fp32 dense (n,h,w,C) accumulated with atomics; dom: NHWC view (>= 27 channels, 32 allocated): [2t] d(dy), [2t+1] d(dx), [18+t] d(mask logit)
template <typename T>
__global__ void deform_col_bwd_kernel(yad_tensor x, const T* __restrict__ om, int om_ld, yad_tensor dcol, float* __restrict__ dx_f, yad_tensor dom) {
  pdl_sync();
  const int C = x.c, oct = C >> 3;
  const int64_t groups = (int64_t)x.n * x.h * x.w * 9;
  const int gpb = blockDim.x / oct;  // (pixel, tap) groups per block
  const int sub = threadIdx.x % oct, gl = threadIdx.x / oct;
  for (int64_t base = (int64_t)blockIdx.x * gpb; base < groups; base += (int64_t)gridDim.x * gpb) {  // block-uniform trip count (shuffles below)
    const bool active = base + gl < groups;
    const int64_t gi = active ? base + gl : groups - 1;
    const int tap = (int)(gi % 9);
    const int64_t p = gi / 9;
    const int ox = (int)(p % x.w), oy = (int)((p / x.w) % x.h), n = (int)(p / ((int64_t)x.w * x.h));
    const Sample s = dcn_sample(om + p * om_ld, tap, oy, ox, x.h, x.w);
    const int o = sub * 8;
    float dmk = 0.f, dpy = 0.f, dpx = 0.f;
    if (s.valid && active) {
      float g[8];
      load8(reinterpret_cast<const T*>(dcol.ptr) + p * dcol.ld + tap * C + o, g);
      const float wgt[4] = {(1.f - s.ly) * (1.f - s.lx), (1.f - s.ly) * s.lx, s.ly * (1.f - s.lx), s.ly * s.lx};
      const float wy[4] = {-(1.f - s.lx), -s.lx, (1.f - s.lx), s.lx};  // d wgt / d ly
      const float wx[4] = {-(1.f - s.ly), (1.f - s.ly), -s.ly, s.ly};  // d wgt / d lx
#pragma unroll
      for (int c4 = 0; c4 < 4; c4++) {
        const int yy = s.y0 + (c4 >> 1), xx = s.x0 + (c4 & 1);
        if (yy >= 0 && yy < x.h && xx >= 0 && xx < x.w) {
          const int64_t q = ((int64_t)n * x.h + yy) * x.w + xx;
          float t[8];
          load8(reinterpret_cast<const T*>(x.ptr) + q * x.ld + o, t);
          const float wm = s.mk * wgt[c4];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            dmk = fmaf(g[i] * wgt[c4], t[i], dmk);
            dpy = fmaf(g[i] * wy[c4], t[i], dpy);
            dpx = fmaf(g[i] * wx[c4], t[i], dpx);
          }
          // 128-bit vector reductions (red.global.add.v4.f32, sm_90+): 2 instead of 8 atomics per corner
          float4* dst = reinterpret_cast<float4*>(dx_f + q * C + o);
          atomicAdd(dst, make_float4(g[0] * wm, g[1] * wm, g[2] * wm, g[3] * wm));
          atomicAdd(dst + 1, make_float4(g[4] * wm, g[5] * wm, g[6] * wm, g[7] * wm));
        }
      }
    }
    for (int d = oct >> 1; d > 0; d >>= 1) {
      dmk += __shfl_xor_sync(0xffffffffu, dmk, d);
      dpy += __shfl_xor_sync(0xffffffffu, dpy, d);
      dpx += __shfl_xor_sync(0xffffffffu, dpx, d);
    }
    if (sub == 0 && active) {
      T* d = reinterpret_cast<T*>(dom.ptr) + p * dom.ld;
      st1(d + 2 * tap, dpy * s.mk);
      st1(d + 2 * tap + 1, dpx * s.mk);
      st1(d + 18 + tap, dmk * s.mk * (1.f - s.mk));
      if (tap == 0)
        for (int i = 27; i < dom.c; i++) st1(d + i, 0.f);
    }
  }
}

int grid_for(int64_t items, int tpb) {
  int64_t g = (items + tpb - 1) / tpb;
  const int64_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int yad_conv_wgrad_tc(const yad_tensor* x, const yad_tensor* dy, const yad_conv_desc* d, float* dw, void* stream);  // wgrad_tc.cu

extern "C" {

/* Weight gradient of yad_conv2d's NORMAL mode: dw fp32 [cout][kh*kw][cin] += ...; the caller zeroes dw once per step (gradients of a
 * weight shared by several calls accumulate).  impl: 0 = auto (tensor cores for bf16), 1 = SIMT. */
int yad_conv_wgrad(const yad_tensor* x, const yad_tensor* dy, const yad_conv_desc* d, float* dw, int dtype, void* stream) {
  YAD_CHECK(d->mode == YAD_CONV_NORMAL, "conv_wgrad: only the normal mode (transposed / deformable convolutions are expressed through it)");
  YAD_CHECK(x->n == dy->n && x->c % 8 == 0 && dy->c % 8 == 0 && x->ld % 8 == 0 && dy->ld % 8 == 0, "conv_wgrad: bad views");
  YAD_CHECK(dy->h == (x->h + 2 * d->pad_h - d->kh) / d->stride + 1 && dy->w == (x->w + 2 * d->pad_w - d->kw) / d->stride + 1,
            "conv_wgrad: dy %dx%d does not match x %dx%d k%dx%d s%d", dy->h, dy->w, x->h, x->w, d->kh, d->kw, d->stride);
  WgGeom g;
  g.n = x->n; g.hi = x->h; g.wi = x->w; g.cin = x->c; g.ho = dy->h; g.wo = dy->w; g.cout = dy->c;
  g.kh = d->kh; g.kw = d->kw; g.stride = d->stride; g.pad_h = d->pad_h; g.pad_w = d->pad_w;
  g.x_ld = x->ld; g.dy_ld = dy->ld;
  g.M = (int64_t)g.n * g.ho * g.wo;
  const int ntaps = g.kh * g.kw;
  const int tiles = cdiv(g.cin, 64) * ntaps * cdiv(g.cout, 64);
  const bool mma = dtype == YAD_BF16 && d->impl != 1;
  if (mma && d->kh == 3 && d->kw == 3 && d->pad_h == 1 && d->pad_w == 1 && (d->stride == 1 || d->stride == 2) && x->c <= 32 && dy->c <= 32 &&
      (x->c == 8 || x->c == 16 || x->c == 32)) {
    // small-channel path: every byte of x and dy is read once, all 9 taps from one staged tile
    const int s = d->stride;
    const int tiles_x = cdiv(g.wo, WS_TW), tiles_y = cdiv(g.ho, WS_TH), total = g.n * tiles_x * tiles_y;
    const int MT = dy->c > 16 ? 2 : 1;
    const int PH = (WS_TH - 1) * s + 3, PW = (WS_TW - 1) * s + 3;
    size_t smem = 2 * ((size_t)PH * PW * ws_pitch(x->c) * 2 + (size_t)WS_TH * WS_TW * ws_pitch(MT * 16) * 2);
    const size_t red = (size_t)MT * 16 * 9 * x->c * 4;
    if (smem < red) smem = red;
    int grid = 148 * 2;
    if (grid > total) grid = total;
    cudaStream_t st = (cudaStream_t)stream;
    if (MT == 2) {
      static bool attr2 = false;
      if (!attr2) { cudaFuncSetAttribute(conv_wgrad_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr2 = true; }
      YAD_LAUNCH(conv_wgrad_small_kernel<2>, grid, WS_THREADS, smem, st, (const bf16*)x->ptr, (const bf16*)dy->ptr, g, dw, tiles_x, tiles_y, total);
    } else {
      static bool attr1 = false;
      if (!attr1) { cudaFuncSetAttribute(conv_wgrad_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr1 = true; }
      YAD_LAUNCH(conv_wgrad_small_kernel<1>, grid, WS_THREADS, smem, st, (const bf16*)x->ptr, (const bf16*)dy->ptr, g, dw, tiles_x, tiles_y, total);
    }
    YAD_LAUNCH_CHECK("conv_wgrad (small-channel)");
    return 0;
  }
  if (mma && d->impl == 0) {  // tcgen05 / TMEM path (stride 1, cin % 64 == 0); -1 = geometry not covered
    const int r = yad_conv_wgrad_tc(x, dy, d, dw, stream);
    if (r >= 0) return r;
  }
  const int step = mma ? WG_P : SB_P;
  int splits = cdiv(148 * 6, tiles);
  const int64_t max_splits = (g.M + step * 4 - 1) / (step * 4);
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  g.per = ((g.M + splits - 1) / splits + step - 1) / step * step;
  splits = (int)((g.M + g.per - 1) / g.per);
  dim3 grid(cdiv(g.cin, 64) * ntaps, cdiv(g.cout, 64), splits);
  cudaStream_t st = (cudaStream_t)stream;
  if (mma) {
    YAD_LAUNCH(conv_wgrad_mma_kernel, grid, 128, 0, st, (const bf16*)x->ptr, (const bf16*)dy->ptr, g, dw);
  } else {
    YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(conv_wgrad_simt_kernel<T>, grid, 256, 0, st, (const T*)x->ptr, (const T*)dy->ptr, g, dw);)
  }
  YAD_LAUNCH_CHECK("conv_wgrad");
  return 0;
}

/* dw fp32 [k*k][c] += ... (accumulating) */
int yad_dwconv_wgrad(const yad_tensor* x, const yad_tensor* dy, int k, float* dw, int dtype, void* stream) {
  YAD_CHECK(x->n == dy->n && x->h == dy->h && x->w == dy->w && x->c == dy->c && x->c % 8 == 0, "dwconv_wgrad: shape mismatch");
  YAD_CHECK((k == 3 || k == 7) && (x->c / 8) * k <= 256, "dwconv_wgrad: k = %d with %d channels is not built (k in {3, 7}, c/8*k <= 256)", k, x->c);
  const int64_t npix = (int64_t)x->n * x->h * x->w;
  const int lanes = 256 / ((x->c / 8) * k);
  int64_t want = (npix + (int64_t)lanes * 16 - 1) / ((int64_t)lanes * 16);  // >= 16 pixels per thread before the final atomics
  int grid = (int)(want < 1 ? 1 : (want > 148 * 2 ? 148 * 2 : want));
  const size_t smem = (size_t)k * k * x->c * sizeof(float);
  YAD_CHECK(smem <= 48 * 1024, "dwconv_wgrad: k = %d with %d channels needs %zu B of shared memory", k, x->c, smem);
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 3) {
    YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH((dwconv_wgrad_kernel<T, 3>), grid, 256, smem, st, *x, *dy, dw);)
  } else {
    YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH((dwconv_wgrad_kernel<T, 7>), grid, 256, smem, st, *x, *dy, dw);)
  }
  YAD_LAUNCH_CHECK("dwconv_wgrad");
  return 0;
}

int yad_deform_col(const yad_tensor* x, const yad_tensor* offmask, const yad_tensor* col, int dtype, void* stream) {
  YAD_CHECK(col->c == 9 * x->c && col->n == x->n && col->h == x->h && col->w == x->w && offmask->c >= 27 && x->c % 8 == 0,
            "deform_col: col must be (n,h,w,9*c) and offmask have >= 27 channels");
  const int per_pix = 9 * (x->c / 8);
  YAD_CHECK(per_pix <= 1024, "deform_col: %d channels are too many", x->c);
  const int tiles_x = (x->w + 7) / 8, tiles_per_img = tiles_x * ((x->h + 7) / 8);
  const int64_t npix_t = (int64_t)x->n * tiles_per_img * 64;
  YAD_CHECK(npix_t * 9 < (1ll << 32), "deform_col: tensor too large for 32-bit indexing");
  const int lanes = per_pix <= 288 ? 288 / per_pix : 1, tpb = lanes * per_pix;
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(deform_col_kernel<T>, grid_for(npix_t, lanes), tpb, 0, (cudaStream_t)stream, *x, (const T*)offmask->ptr, offmask->ld, *col,
                                                                                                          tiles_x, tiles_per_img, (uint32_t)npix_t);)
  YAD_LAUNCH_CHECK("deform_col");
  return 0;
}

/* dx_f: fp32 dense (n,h,w,c), zeroed by this call */
int yad_deform_col_bwd(const yad_tensor* x, const yad_tensor* offmask, const yad_tensor* dcol, float* dx_f, const yad_tensor* doffmask, int dtype,
                       void* stream) {
  const int oct = x->c / 8;
  YAD_CHECK(dcol->c == 9 * x->c && offmask->c >= 27 && doffmask->c >= 27 && x->c % 8 == 0 && (oct & (oct - 1)) == 0 && oct <= 32,
            "deform_col_bwd: channel count %d must be 8 * a power of two <= 256", x->c);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dx_f, 0, sizeof(float) * (int64_t)x->n * x->h * x->w * x->c, st);
  const int64_t groups = (int64_t)x->n * x->h * x->w * 9;
  const int gpb = 256 / oct;
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(deform_col_bwd_kernel<T>, grid_for(groups, gpb), 256, 0, st, *x, (const T*)offmask->ptr, offmask->ld, *dcol, dx_f, *doffmask);)
  YAD_LAUNCH_CHECK("deform_col_bwd");
  return 0;
}

}  // extern "C"
