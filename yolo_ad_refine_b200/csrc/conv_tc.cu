// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
// GEMM view per launch:  D[128 pixels][N_TILE couts] += A[128][64] * B[N_TILE][64]^T  per 64-wide K chunk, K = taps * cin.
//   * A (im2col rows) and B (weights) are staged in shared memory in the canonical K-major SWIZZLE_128B layout (row = 128 B, 16-byte
//     chunk c of row r stored at chunk position c ^ (r & 7); tile base 1024-byte aligned) -- the layout a SWIZZLE_128B TMA box would
//     produce -- by 128 producer threads (8 lanes sweep one 128-byte row: coalesced 128-bit loads, zero-fill for padding / ragged edges / K tail),
//     made visible to the tensor core with fence.proxy.async, and handed over through a ring of mbarriers.
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x N_TILE x 16) and commits to the "slot empty" /
//     "accumulator full" mbarriers (tcgen05.commit).
//   * the 128 producer threads then become the epilogue: tcgen05.ld (32 lanes x 32 bit x 16 columns per warp), fused
//     bias / per-image / per-pixel scale / activation / alpha / mul / add, 128-bit bf16 stores into the (possibly channel-sliced) output.
// Modes: normal (1x1, 3x3, 7x1; stride 1 or 2), transposed 3x3 s2 (4 sub-pixel phases, each a dense GEMM over the input grid), and
// modulated deformable 3x3 (bilinear gather in the producer).
#include <cuda.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;  // UMMA M (pixels per CTA)
constexpr int BK = 64;   // K elements per stage (= 128 bytes of bf16 = one swizzle row)
constexpr int MAX_TAPS = 9;
constexpr int NPROD = 128;         // producer / epilogue threads (warps 0-3)
constexpr int NTHREADS = NPROD + 32;  // + the MMA warp

struct TcParams {
  const bf16* x;
  const bf16* w;
  const bf16* om;
  void* y;
  int n, hi, wi, cin, x_ld;
  int hm, wm;  // logical pixel grid per image that the M index enumerates
  int ho, wo, cout, y_ld;
  int stride;      // source pixel = stride * m + dy[tap]
  int os, py, px;  // destination pixel = os * m + py
  int ntaps;
  int dy[MAX_TAPS], dx[MAX_TAPS], wtap[MAX_TAPS];
  int w_row;  // elements per weight row (all taps * cin)
  int deform, om_ld;
  int n_tile, stages, tmem_cols, pipe_bytes;
  int out_f32;
  int cout_real;  // channel count GroupNorm statistics are defined over (= cout of the view)
  int bn_stats;   // 1: e.gn_stats is double [cout][2] = per-channel (sum, sum of squares) over the WHOLE batch (train-mode BatchNorm)
  yad_epilogue e;
};

// (PTX wrappers: tc_ptx.cuh)
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO(16 B) << 16 | SBO(1024 B) << 32 |
// version 1 << 46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// generic K-major descriptor: row_bytes in {32, 64, 128} <-> SWIZZLE_32B (6) / SWIZZLE_64B (4) / SWIZZLE_128B (2); SBO = 8 rows
__device__ __forceinline__ uint64_t make_sdesc_rb(uint32_t saddr, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)((8u * row_bytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format BF16 (1) @7/@10, K-major A and B, N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t make_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }

__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void sts16(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void bf8_to_f(const uint4& u, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 f_to_bf8(const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}

// bf16-grade activations for the tensor-core epilogue: one MUFU.TANH per value (sigmoid(x) = 0.5 + 0.5 tanh(x / 2))
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int N>
__device__ __forceinline__ void apply_act_bf16_n(float (&v)[N], int act) {
  if (act == YAD_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < N; i++) { const float h = 0.5f * v[i]; v[i] = fmaf(h, tanh_approx(h), h); }
  } else if (act == YAD_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = fmaf(0.5f, tanh_approx(0.5f * v[i]), 0.5f);
  } else if (act != YAD_ACT_NONE) {
    apply_act_n<N>(v, act);
  }
}

// Fused GroupNorm statistics: the 16 freshly computed output channels [co, co + 16) of this thread's row are folded into per-(image, group)
// sums.  cpg = channels per group in {4, 8, 16} (host-checked).  Rows of a warp normally belong to one image: warp-shuffle reduction and
// one double atomicAdd pair per group; a warp that straddles two images falls back to per-thread atomics.
template <int CPG>
__device__ __forceinline__ void gn_accumulate_t(const TcParams& p, const float (&v)[16], int co, int dp, int img) {
  constexpr int NG = 16 / CPG;
  const int groups = p.e.gn_groups;
  const bool valid = dp >= 0;
  const int img0 = __shfl_sync(0xffffffffu, valid ? img : -1, 0);
  const bool uniform = __all_sync(0xffffffffu, !valid || img == img0) && img0 >= 0;
  float s[NG], q[NG];
#pragma unroll
  for (int g = 0; g < NG; g++) {
    s[g] = 0.f; q[g] = 0.f;
#pragma unroll
    for (int i = 0; i < CPG; i++) { const float x = valid ? v[g * CPG + i] : 0.f; s[g] += x; q[g] = fmaf(x, x, q[g]); }
  }
  const int g0 = co / CPG;
  if (uniform) {
#pragma unroll
    for (int g = 0; g < NG; g++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s[g] += __shfl_xor_sync(0xffffffffu, s[g], o); q[g] += __shfl_xor_sync(0xffffffffu, q[g], o); }
    }
    // lanes 0 .. 2*NG-1 each issue one atomic
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int g = 0; g < NG; g++) {
      if (g0 + g < groups) {
        if (lane == 2 * g) atomicAdd(&p.e.gn_stats[((int64_t)img0 * groups + g0 + g) * 2], (double)s[g]);
        if (lane == 2 * g + 1) atomicAdd(&p.e.gn_stats[((int64_t)img0 * groups + g0 + g) * 2 + 1], (double)q[g]);
      }
    }
  } else if (valid) {
#pragma unroll
    for (int g = 0; g < NG; g++) {
      if (g0 + g < groups) {
        atomicAdd(&p.e.gn_stats[((int64_t)img * groups + g0 + g) * 2], (double)s[g]);
        atomicAdd(&p.e.gn_stats[((int64_t)img * groups + g0 + g) * 2 + 1], (double)q[g]);
      }
    }
  }
}
__device__ __forceinline__ void gn_accumulate(const TcParams& p, const float (&v)[16], int co, int dp, int img) {
  const int cpg = p.cout_real / p.e.gn_groups;
  if (cpg == 4) gn_accumulate_t<4>(p, v, co, dp, img);
  else if (cpg == 8) gn_accumulate_t<8>(p, v, co, dp, img);
  else gn_accumulate_t<16>(p, v, co, dp, img);
}

// Per-channel batch statistics (train-mode BatchNorm) fused into phase 2 of the epilogue: in phase 2 a lane owns 8 fixed channels of a 64-column
// chunk and sweeps rows, so its partial sums accumulate in registers across rows AND across the tiles of a persistent CTA; they meet once,
// at the end (or when the CTA moves to another column tile): shuffle reduction over the lanes that share the channels, then one double
// atomicAdd pair per channel per warp.  The sums are taken over the bf16-rounded values that are stored (what the BatchNorm kernels read back).
struct BnAcc {
  float s[2][8], q[2][8];  // [64-column chunk of this warp's column range][channel]
  int n0;                  // column tile the sums belong to (-1: empty)
};
__device__ __forceinline__ void bn_acc_clear(BnAcc& a) {
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int i = 0; i < 8; i++) { a.s[c][i] = 0.f; a.q[c][i] = 0.f; }
  a.n0 = -1;
}
__device__ __forceinline__ void bn_acc_flush(const TcParams& p, BnAcc& a, int lane, int col_begin, int col_end) {
  if (a.n0 < 0) return;  // warp-uniform
#pragma unroll
  for (int ci = 0; ci < 2; ci++) {
    const int c0 = col_begin + ci * 64;
    if (c0 >= col_end) break;
    const int cw = min(64, col_end - c0), cpr = cw >> 3, rpp = 32 / cpr;
    const int rr = lane / cpr, ch = (lane - rr * cpr) * 8, co = a.n0 + c0 + ch;
    const bool pow2 = (cpr & (cpr - 1)) == 0;
    if (pow2) {
      for (int d = cpr; d < 32; d <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          a.s[ci][i] += __shfl_xor_sync(0xffffffffu, a.s[ci][i], d);
          a.q[ci][i] += __shfl_xor_sync(0xffffffffu, a.q[ci][i], d);
        }
      }
    }
    if ((pow2 ? rr == 0 : rr < rpp) && co < p.cout) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (co + i < p.cout_real) {
          atomicAdd(&p.e.gn_stats[(int64_t)(co + i) * 2], (double)a.s[ci][i]);
          atomicAdd(&p.e.gn_stats[(int64_t)(co + i) * 2 + 1], (double)a.q[ci][i]);
        }
      }
    }
  }
  bn_acc_clear(a);
}

constexpr int STG_ROW = 144;            // bytes per staged row: 64 bf16 + 16 (16-byte aligned rows, conflict-free 128-bit accesses)
constexpr int STG_WARP = 32 * STG_ROW;  // staging bytes per epilogue warp

// Epilogue of one warp over the 32 accumulator rows of its TMEM lane quarter and the column range [col_begin, col_end) of the tile.
//   phase 1 (thread = TMEM lane = tile row): tcgen05.ld 16 columns at a time, scale / bias / activation / alpha, round to bf16 and park
//   in a per-warp smem staging tile; phase 2 (lanes sweep each row contiguously): optional mul / add, 128-bit coalesced stores.
//   (The conv output is rounded to bf16 before mul / add -- exactly what the unfused bf16 reference path materialises.)
// dp = destination pixel index of this thread's row (or -1), img = its image (for img_scale).
template <bool BN>
__device__ __forceinline__ void epilogue_warp(const TcParams& p, uint32_t tmem_acc, int quarter, int lane, uint8_t* stg, int* drow, int dp, int img,
                                              int n0, int col_begin, int col_end, BnAcc& bn) {
  const yad_epilogue& e = p.e;
  float sc = 1.0f;
  if (dp >= 0) {
    if (e.img_scale) sc = e.img_scale[img];
    if (e.pix_scale) sc *= __bfloat162float(reinterpret_cast<const bf16*>(e.pix_scale)[(int64_t)dp * e.pix_scale_ld]);
  }
  drow[lane] = dp;
  const uint32_t lane_base = ((uint32_t)(quarter * 32)) << 16;
  for (int c0 = col_begin; c0 < col_end; c0 += 64) {
    const int cw = min(64, col_end - c0);
    __syncwarp();
    for (int q0 = 0; q0 < cw; q0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_acc + lane_base + (uint32_t)(c0 + q0), r);
      const int co = n0 + c0 + q0;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
      if (p.out_f32) {  // self-test path: raw accumulators straight to global
        if (dp >= 0) {
          float* o = reinterpret_cast<float*>(p.y) + (int64_t)dp * p.y_ld + co;
#pragma unroll
          for (int i = 0; i < 16; i++)
            if (co + i < p.cout) o[i] = v[i];
        }
        continue;
      }
      if (e.bias && co + 16 <= p.cout) {
#pragma unroll
        for (int i4 = 0; i4 < 4; i4++) {
          const float4 b4 = *reinterpret_cast<const float4*>(e.bias + co + 4 * i4);
          v[4 * i4] = fmaf(v[4 * i4], sc, b4.x); v[4 * i4 + 1] = fmaf(v[4 * i4 + 1], sc, b4.y);
          v[4 * i4 + 2] = fmaf(v[4 * i4 + 2], sc, b4.z); v[4 * i4 + 3] = fmaf(v[4 * i4 + 3], sc, b4.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = fmaf(v[i], sc, (e.bias && co + i < p.cout) ? e.bias[co + i] : 0.f);
      }
      apply_act_bf16_n<16>(v, e.act);
      if (e.alpha != 1.0f) {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] *= e.alpha;
      }
      if (!BN && e.gn_stats) gn_accumulate(p, v, co, dp, img);
      float lo[8], hi[8];
#pragma unroll
      for (int i = 0; i < 8; i++) { lo[i] = v[i]; hi[i] = v[8 + i]; }
      uint4* dst = reinterpret_cast<uint4*>(stg + lane * STG_ROW + q0 * 2);
      dst[0] = f_to_bf8(lo);
      dst[1] = f_to_bf8(hi);
    }
    if (p.out_f32) continue;
    __syncwarp();
    // phase 2: cpr lanes per row, rpp rows per pass (cw in {16, 32, 48, 64} -> cpr in {2, 4, 6, 8})
    const int cpr = cw >> 3, rpp = 32 / cpr;
    const int rr = lane / cpr, ch = (lane - rr * cpr) * 8;
    const int co = n0 + c0 + ch;
    const int bci = (c0 - col_begin) >> 6;
    if (BN && bn.n0 != n0) {  // warp-uniform: this CTA moved to another column tile
      bn_acc_flush(p, bn, lane, col_begin, col_end);
      bn.n0 = n0;
    }
    if (rr < rpp && co < p.cout) {
      for (int row = rr; row < 32; row += rpp) {
        const int d = drow[row];
        if (d < 0) continue;
        uint4 u = *reinterpret_cast<const uint4*>(stg + row * STG_ROW + ch * 2);
        if (BN) {  // static indices only: the accumulators must stay in registers
          float bv[8];
          bf8_to_f(u, bv);
          if (bci == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) { bn.s[0][i] += bv[i]; bn.q[0][i] = fmaf(bv[i], bv[i], bn.q[0][i]); }
          } else if (bci == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) { bn.s[1][i] += bv[i]; bn.q[1][i] = fmaf(bv[i], bv[i], bn.q[1][i]); }
          }
        }
        if (e.mul || e.add) {
          float v[8];
          bf8_to_f(u, v);
          if (e.mul) {
            float mv[8];
            bf8_to_f(ldg16(reinterpret_cast<const bf16*>(e.mul) + (int64_t)d * e.mul_ld + co), mv);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] *= mv[i];
          }
          if (e.add) {
            float adv[8];
            bf8_to_f(ldg16(reinterpret_cast<const bf16*>(e.add) + (int64_t)d * e.add_ld + co), adv);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] += adv[i];
          }
          u = f_to_bf8(v);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.y) + (int64_t)d * p.y_ld + co) = u;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
template <bool DEFORM>
__global__ void __launch_bounds__(NTHREADS) conv_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // carve: [1024-aligned] stages x (A 16 KB + B n_tile*128 B), then barriers
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = BM * 128, b_bytes = (uint32_t)p.n_tile * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t bars = base + (uint32_t)p.pipe_bytes;  // full[stages], empty[stages], tmem_full, tmem_ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * p.stages);
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = p.ntaps * p.cin;
  const int nk = (K + BK - 1) / BK;
  const int n0 = blockIdx.y * p.n_tile;

  if (tid == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), NPROD); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_sync();  // on-chip prologue done; from here on the kernel reads what its predecessors wrote

  if (warp < 4) {
    // ================= producer =================
    // Coalesced gather: lane l of warp w handles the 16-byte chunk c = l & 7 of rows 32w + 4i + (l >> 3), i = 0..7, so one warp
    // instruction reads 4 rows x 128 contiguous bytes (4 L1 wavefronts instead of 32) and writes 4 conflict-free swizzled smem rows.
    const int64_t M = (int64_t)p.n * p.hm * p.wm;
    const int lane = tid & 31, sr = lane >> 3, c = lane & 7;
    int rimg[8], ryx[8];  // per owned row: image index (-1 = beyond M) and (stride*my << 16 | stride*mx) (DEFORM: stride = 1)
    {
      // decode the first owned row with 32-bit divisions, then step by 4 pixels with carries (M < 2^31 is checked by the launcher)
      const uint32_t m0 = (uint32_t)blockIdx.x * BM + warp * 32 + sr, hw = (uint32_t)(p.hm * p.wm);
      int img = (int)(m0 / hw);
      const uint32_t r = m0 - (uint32_t)img * hw;
      int my = (int)(r / (uint32_t)p.wm), mx = (int)(r - (uint32_t)my * (uint32_t)p.wm);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const bool ok = (int64_t)m0 + 4 * i < M;
        rimg[i] = ok ? img * p.hi : -1;  // pre-multiplied row base of the image
        ryx[i] = ((p.stride * my) << 16) | (p.stride * mx);
        mx += 4;
        while (mx >= p.wm) { mx -= p.wm; my++; }
        while (my >= p.hm) { my -= p.hm; img++; }
      }
    }
    // DEFORM: sampling parameters of (row, tap) are computed once per CTA by the row's own thread (offsets, sigmoid(mask), bilinear
    // weights with invalid corners zeroed, clamped corner position) and parked in smem; the gather below only does 4 loads + 32 FMAs.
    float4* dpar_w = reinterpret_cast<float4*>(smem_raw + (((tmem_ptr_addr + 16u + 15u) & ~15u) - raw));
    int* dpar_pos = reinterpret_cast<int*>(dpar_w + BM * MAX_TAPS);
    if (DEFORM) {
      const uint32_t m = (uint32_t)blockIdx.x * BM + tid, hw = (uint32_t)(p.hm * p.wm);
      if ((int64_t)m < M) {
        const int img = (int)(m / hw);
        const int r = (int)(m - (uint32_t)img * hw);
        const int my = r / p.wm, mx = r - my * p.wm;
        const bf16* o = p.om + ((int64_t)(img * p.ho + my) * p.wo + mx) * p.om_ld;
        for (int t = 0; t < p.ntaps; t++) {
          const float ody = __bfloat162float(o[2 * t]), odx = __bfloat162float(o[2 * t + 1]);
          const float mk = sigmoidf_(__bfloat162float(o[18 + t]));
          const float fy_ = (float)(my + p.dy[t]) + ody, fx_ = (float)(mx + p.dx[t]) + odx;
          float4 wq = make_float4(0.f, 0.f, 0.f, 0.f);
          int y0 = 0, x0 = 0;
          if (fy_ > -1.f && fx_ > -1.f && fy_ < (float)p.hi && fx_ < (float)p.wi) {
            const float fy = floorf(fy_), fx = floorf(fx_);
            y0 = (int)fy; x0 = (int)fx;
            const float ly = fy_ - fy, lx = fx_ - fx;
            const bool vy0 = y0 >= 0, vy1 = y0 + 1 < p.hi, vx0 = x0 >= 0, vx1 = x0 + 1 < p.wi;
            wq.x = (vy0 && vx0) ? (1.f - ly) * (1.f - lx) * mk : 0.f;
            wq.y = (vy0 && vx1) ? (1.f - ly) * lx * mk : 0.f;
            wq.z = (vy1 && vx0) ? ly * (1.f - lx) * mk : 0.f;
            wq.w = (vy1 && vx1) ? ly * lx * mk : 0.f;
          }
          dpar_w[tid * MAX_TAPS + t] = wq;
          dpar_pos[tid * MAX_TAPS + t] = (y0 << 16) | (x0 & 0xFFFF);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // producer warps only
    }
    RingPos ring;
    for (int kc = 0; kc < nk; kc++, ring.next(p.stages)) {
      const int s = ring.idx;
      const uint32_t ph = ring.ph;
      const uint32_t a_s = base + s * stage_bytes, b_s = a_s + a_bytes;
      const int kk = kc * BK + c * 8;  // this thread's K offset inside the chunk: fixed tap / channel for all its rows
      const bool kvalid = kk < K;
      const int t = kvalid ? kk / p.cin : 0;
      const int ci = kk - t * p.cin;
      const int tdy = p.dy[t], tdx = p.dx[t];
      const bf16* xk = p.x + ci;
      // ---- A: gather into registers first (global latency overlaps the wait for the slot)
      uint4 av[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        av[i] = make_uint4(0u, 0u, 0u, 0u);
        if (!kvalid || rimg[i] < 0) continue;
        const int my = ryx[i] >> 16, mx = ryx[i] & 0xFFFF;
        if (!DEFORM) {
          const int sy = my + tdy, sx = mx + tdx;
          if ((unsigned)sy < (unsigned)p.hi && (unsigned)sx < (unsigned)p.wi)
            av[i] = ldg16(xk + (int64_t)((rimg[i] + sy) * p.wi + sx) * p.x_ld);
        } else {
          // modulated deformable 3x3: parameters from smem, corners clamped into the map (their weights are already zero when invalid)
          const int rrow = warp * 32 + 4 * i + sr;
          const float4 wq = dpar_w[rrow * MAX_TAPS + t];
          const int pos = dpar_pos[rrow * MAX_TAPS + t];
          const int y0 = pos >> 16, x0 = (int)(short)(pos & 0xFFFF);
          const int ya = max(y0, 0), yb = min(y0 + 1, p.hi - 1), xa = max(x0, 0), xb = min(x0 + 1, p.wi - 1);
          const int ib = rimg[i];
          const uint4 u00 = ldg16(xk + (int64_t)((ib + ya) * p.wi + xa) * p.x_ld), u01 = ldg16(xk + (int64_t)((ib + ya) * p.wi + xb) * p.x_ld);
          const uint4 u10 = ldg16(xk + (int64_t)((ib + yb) * p.wi + xa) * p.x_ld), u11 = ldg16(xk + (int64_t)((ib + yb) * p.wi + xb) * p.x_ld);
          // blend in packed bf16x2 (the result is rounded to bf16 for the MMA anyway): 16 HFMA2 instead of 32 cvt + 32 FMA + 4 pack
          const __nv_bfloat162 w00 = __float2bfloat162_rn(wq.x), w01 = __float2bfloat162_rn(wq.y), w10 = __float2bfloat162_rn(wq.z),
                               w11 = __float2bfloat162_rn(wq.w);
          const __nv_bfloat162* h00 = reinterpret_cast<const __nv_bfloat162*>(&u00);
          const __nv_bfloat162* h01 = reinterpret_cast<const __nv_bfloat162*>(&u01);
          const __nv_bfloat162* h10 = reinterpret_cast<const __nv_bfloat162*>(&u10);
          const __nv_bfloat162* h11 = reinterpret_cast<const __nv_bfloat162*>(&u11);
          uint4 res;
          __nv_bfloat162* hr = reinterpret_cast<__nv_bfloat162*>(&res);
#pragma unroll
          for (int q = 0; q < 4; q++) hr[q] = __hfma2(w11, h11[q], __hfma2(w10, h10[q], __hfma2(w01, h01[q], __hmul2(w00, h00[q]))));
          av[i] = res;
        }
      }
      mbar_wait(empty_bar(s), ph ^ 1u);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int r = warp * 32 + 4 * i + sr;
        sts16(a_s + r * 128 + ((c ^ (r & 7)) << 4), av[i]);
      }
      // ---- B: weight rows (L2 resident), 16 rows per pass over the 4 warps
      const int64_t woff = (int64_t)p.wtap[t] * p.cin + ci;
      for (int r0 = 0; r0 < p.n_tile; r0 += 64) {
        uint4 bv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int row = r0 + 16 * j + 4 * warp + sr, co = n0 + row;
          bv[j] = make_uint4(0u, 0u, 0u, 0u);
          if (kvalid && row < p.n_tile && co < p.cout) bv[j] = ldg16(p.w + (int64_t)co * p.w_row + woff);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int row = r0 + 16 * j + 4 * warp + sr;
          if (row < p.n_tile) sts16(b_s + row * 128 + ((c ^ (row & 7)) << 4), bv[j]);
        }
      }
      fence_proxy_async();
      mbar_arrive(full_bar(s));
    }

    // ================= epilogue =================
    mbar_wait(tmem_full_bar, 0u);
    tc_fence_after();
    // the pipeline buffers are dead now (every MMA that read them has completed): alias the staging tile onto them
    uint8_t* stg = smem_raw + (base - raw) + warp * STG_WARP;
    int* drow = reinterpret_cast<int*>(smem_raw + (base - raw) + 4 * STG_WARP) + warp * 32;
    {
      const uint32_t m = (uint32_t)blockIdx.x * BM + tid, hw = (uint32_t)(p.hm * p.wm);
      int dp = -1, img = 0;
      if ((int64_t)m < M) {
        img = (int)(m / hw);
        const int r = (int)(m - (uint32_t)img * hw);
        const int my = r / p.wm, mx = r - my * p.wm;
        dp = (img * p.ho + p.os * my + p.py) * p.wo + (p.os * mx + p.px);
      }
      BnAcc bn;  // unused: batch statistics are fused on the TMA-fed kernel only (the host runs yad_gn_stats after this kernel otherwise)
      epilogue_warp<false>(p, tmem_base, warp, lane, stg, drow, dp, img, n0, 0, p.n_tile, bn);
    }
    tc_fence_before();
  } else {
    // ================= MMA issuer: warp-uniform loop, the elected lane issues (tc_ptx.cuh) =================
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc(p.n_tile);
      const uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      RingPos sr;
      for (int kc = 0; kc < nk; kc++, sr.next(p.stages)) {
        const int s = sr.idx;
        mbar_wait(full_bar(s), sr.ph);
        tc_fence_after();
        const uint32_t a_s = base + s * stage_bytes, b_s = a_s + a_bytes;
        const uint32_t a_lo = ((a_s & 0x3FFFFu) >> 4) | (1u << 16), b_lo = ((b_s & 0x3FFFFu) >> 4) | (1u << 16);
        if (leader) {
          if (kc == 0) umma_bf16<false>(tmem_base, pack64(a_lo, hi), pack64(b_lo, hi), idesc);
          else umma_bf16<true>(tmem_base, pack64(a_lo, hi), pack64(b_lo, hi), idesc);
          umma_bf16<true>(tmem_base, pack64(a_lo + 2u, hi), pack64(b_lo + 2u, hi), idesc);  // 32 bytes along K inside the 128-byte swizzle row
          umma_bf16<true>(tmem_base, pack64(a_lo + 4u, hi), pack64(b_lo + 4u, hi), idesc);
          umma_bf16<true>(tmem_base, pack64(a_lo + 6u, hi), pack64(b_lo + 6u, hi), idesc);
          umma_commit(empty_bar(s));  // implies tcgen05.fence::before_thread_sync
        }
      }
      if (leader) umma_commit(tmem_full_bar);
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}


// =====================================================================================================================
// TMA-fed variant: stride-1 convolutions whose 64-wide K chunks lie inside one tap (1x1 with any cin; kxk with cin % 64 == 0).
//   warp 0 / lane 0 : producer -- per K chunk one cp.async.bulk.tensor load of the A box (64 channels x BW x BH pixels of one image, or
//                     64 channels x 128 consecutive pixels for 1x1) and one of the B box (64 K x N_TILE couts), SWIZZLE_128B, arriving on
//                     the stage's mbarrier with complete_tx; out-of-bounds coordinates (conv padding, ragged tiles, K tail) are zero-filled
//                     by the TMA unit, so there is no address arithmetic in the kernel at all.
//   warp 1 / lane 0 : tcgen05.mma issuer (same as above); warp 1 also owns the TMEM allocation.
//   warps 2-9       : epilogue (TMEM lane quarter = warp % 4; the two warps of a quarter split the tile's columns).
// =====================================================================================================================
// Epilogue of one warp through a TMA store (1x1 convolutions, rows = consecutive pixels): phase 1 as in epilogue_warp, but the bf16 tile is staged
// in the shared-memory image of a (store_cols channels x 32 rows) box of the output tensor map -- rows of 2 * store_cols bytes, 16-byte chunks XOR-ed
// with address bits [7, 10) as SWIZZLE_32B / 64B / 128B prescribe (conflict-free 128-bit stores) -- and lane 0 hands the box to the TMA unit.
// Rows beyond the last pixel and channels beyond cout are clipped by the unit.  The staging tile is reused only after the previous store has
// finished reading it (cp.async.bulk.wait_group.read), which in steady state happened a whole tile ago.
__device__ __forceinline__ void epilogue_warp_tma(const TcParams& p, const CUtensorMap* tmY, uint32_t tmem_acc, int quarter, int lane, uint32_t stg_addr,
                                                  uint8_t* stg, int row0, int dp, int img, int n0, int col_begin, int col_end, int store_cols) {
  const yad_epilogue& e = p.e;
  float sc = 1.0f;
  if (dp >= 0) {
    if (e.img_scale) sc = e.img_scale[img];
    if (e.pix_scale) sc *= __bfloat162float(reinterpret_cast<const bf16*>(e.pix_scale)[(int64_t)dp * e.pix_scale_ld]);
  }
  const uint32_t lane_base = ((uint32_t)(quarter * 32)) << 16;
  const uint32_t row_bytes = 2u * (uint32_t)store_cols, swz_mask = (row_bytes >> 4) - 1u;  // 1 / 3 / 7
  const uint32_t row_off = (uint32_t)lane * row_bytes;
  for (int c0 = col_begin; c0 < col_end; c0 += store_cols) {
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
    for (int q0 = 0; q0 < store_cols; q0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_acc + lane_base + (uint32_t)(c0 + q0), r);
      const int co = n0 + c0 + q0;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
      if (e.bias && co + 16 <= p.cout) {
#pragma unroll
        for (int i4 = 0; i4 < 4; i4++) {
          const float4 b4 = *reinterpret_cast<const float4*>(e.bias + co + 4 * i4);
          v[4 * i4] = fmaf(v[4 * i4], sc, b4.x); v[4 * i4 + 1] = fmaf(v[4 * i4 + 1], sc, b4.y);
          v[4 * i4 + 2] = fmaf(v[4 * i4 + 2], sc, b4.z); v[4 * i4 + 3] = fmaf(v[4 * i4 + 3], sc, b4.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = fmaf(v[i], sc, (e.bias && co + i < p.cout) ? e.bias[co + i] : 0.f);
      }
      apply_act_bf16_n<16>(v, e.act);
      if (e.alpha != 1.0f) {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] *= e.alpha;
      }
      if (e.gn_stats) gn_accumulate(p, v, co, dp, img);
      float lo[8], hi[8];
#pragma unroll
      for (int i = 0; i < 8; i++) { lo[i] = v[i]; hi[i] = v[8 + i]; }
      const uint32_t a0 = stg_addr + row_off + (uint32_t)q0 * 2u, a1 = a0 + 16u;
      const uint32_t o0 = (a0 ^ (((a0 >> 7) & swz_mask) << 4)) - stg_addr, o1 = (a1 ^ (((a1 >> 7) & swz_mask) << 4)) - stg_addr;
      *reinterpret_cast<uint4*>(stg + o0) = f_to_bf8(lo);
      *reinterpret_cast<uint4*>(stg + o1) = f_to_bf8(hi);
    }
    if (e.mul || e.add) {  // phase 2 on the staged tile: lanes sweep each row contiguously, the operands are read with 128-bit coalesced loads
      __syncwarp();
      const int cpr = store_cols >> 3, rpp = 32 / cpr;  // 2 / 4 / 8 lanes per row
      const int rr = lane / cpr, ch = (lane - rr * cpr) * 8, co = n0 + c0 + ch;
      const int64_t M = (int64_t)p.n * p.ho * p.wo;
      if (co < p.cout) {
        for (int row = rr; row < 32; row += rpp) {
          const int64_t d = (int64_t)row0 + row;
          if (d >= M) break;
          const uint32_t a = stg_addr + (uint32_t)row * row_bytes + (uint32_t)ch * 2u;
          uint4* cell = reinterpret_cast<uint4*>(stg + ((a ^ (((a >> 7) & swz_mask) << 4)) - stg_addr));
          float v[8];
          bf8_to_f(*cell, v);
          if (e.mul) {
            float mv[8];
            bf8_to_f(ldg16(reinterpret_cast<const bf16*>(e.mul) + d * e.mul_ld + co), mv);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] *= mv[i];
          }
          if (e.add) {
            float adv[8];
            bf8_to_f(ldg16(reinterpret_cast<const bf16*>(e.add) + d * e.add_ld + co), adv);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] += adv[i];
          }
          *cell = f_to_bf8(v);
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmY, stg_addr, n0 + c0, row0);
      tma_store_commit();
    }
  }
}

constexpr int TMA_EPI_WARPS = 8;                       // two warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int TMA_THREADS = 64 + 32 * TMA_EPI_WARPS;

struct TmaParams {
  TcParams p;
  int patch;            // 0: rows = 128 consecutive pixels (1x1, 2-D map); 1: rows = BH x BW patch of one image (4-D map)
  int bw, bh, tiles_x, tiles_y;
  int a_bytes;          // bytes TMA writes per A box
  int tiles_n, total_tiles, acc_stages;
  int bk;               // K elements per chunk: 16 / 32 / 64 (smem row = 2*bk bytes, SWIZZLE_32B / 64B / 128B)
  int store_cols;       // > 0: the epilogue leaves through TMA (1x1 convolutions without fused batch statistics): every epilogue warp stages its
                        //      32 rows x store_cols channels (16 / 32 / 64) in the swizzle of the output map and one lane issues a bulk tensor store
  int s2;               // stride-2 mode: 5-D parity-split map, coordinates (cpx[t] + ci, tx0 + dx[t], tpy[t], ty0 + dy[t], img)
  int cpx[MAX_TAPS], tpy[MAX_TAPS];
};

template <bool BN>  // BN: fused per-channel batch statistics (train-mode BatchNorm); a separate instantiation keeps the inference kernel lean
__global__ void __launch_bounds__(TMA_THREADS) conv_tma_kernel(const __grid_constant__ TmaParams tp, const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY) {
  // Persistent: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA producer runs ahead across tile boundaries and the
  // accumulator is double-buffered in TMEM (when 2 * N_TILE <= 512 columns), so loads, MMAs and the epilogue of consecutive tiles overlap.
  extern __shared__ uint8_t smem_raw[];
  const TcParams& p = tp.p;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t row_bytes = 2u * (uint32_t)tp.bk;
  const uint32_t a_bytes = BM * row_bytes, b_bytes = (uint32_t)p.n_tile * row_bytes;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
  const uint32_t stg_off = (uint32_t)p.stages * stage_bytes;             // dedicated epilogue staging (not aliased: the pipeline stays busy)
  const uint32_t bars = base + (uint32_t)p.pipe_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * p.stages + 2 + a); };
  const uint32_t tmem_ptr_addr = bars + 8u * (2 * p.stages + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kpt = (p.cin + tp.bk - 1) / tp.bk;  // K chunks per tap
  const int nk = p.ntaps * kpt;
  const int acc_stages = tp.acc_stages;
  const int rx_ = tp.patch ? tp.tiles_x : 1, ry_ = tp.patch ? tp.tiles_y : 1;  // tile-grid radices (flat launches: the M tile index is the image digit)

  if (tid == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TMA_EPI_WARPS); }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (tp.store_cols) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_sync();  // on-chip prologue done; from here on the kernel reads what its predecessors wrote

  if (warp == 0) {
    {  // warp-uniform producer loop, the elected lane issues
      const bool leader = elect_one();
      const uint32_t tx_bytes = (uint32_t)tp.a_bytes + b_bytes;
      TileDigits ti;
      ti.init((int)blockIdx.x, (int)gridDim.x, tp.tiles_n, rx_, ry_);
      RingPos sr;  // pipeline slot of the running K chunk (across tiles)
      for (int tile = blockIdx.x; tile < tp.total_tiles; tile += gridDim.x, ti.next(tp.tiles_n, rx_, ry_)) {
        const int mt = ti.img, n0 = ti.nt * p.n_tile;  // flat launches: rx_ = ry_ = 1, the image digit is the M tile
        const int img = ti.img, ty0 = ti.ty * tp.bh, tx0 = ti.tx * tp.bw;
        int t = 0, ci = 0;
        for (int kc = 0; kc < nk; kc++, sr.next(p.stages)) {
          const int s = sr.idx;
          const uint32_t ph = sr.ph;
          const uint32_t a_s = base + s * stage_bytes, b_s = a_s + a_bytes;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (leader) {
            mbar_expect_tx(full_bar(s), tx_bytes);
            if (tp.s2)
              tma_load_5d(a_s, &tmA, full_bar(s), tp.cpx[t] + ci, tx0 + p.dx[t], tp.tpy[t], ty0 + p.dy[t], img);
            else if (tp.patch)
              tma_load_4d(a_s, &tmA, full_bar(s), ci, tx0 + p.dx[t], ty0 + p.dy[t], img);
            else
              tma_load_2d(a_s, &tmA, full_bar(s), ci, mt * BM);
            tma_load_2d(b_s, &tmB, full_bar(s), p.wtap[t] * p.cin + ci, n0);
          }
          ci += tp.bk;
          if (ci >= kpt * tp.bk) { ci = 0; t++; }
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop (see tc_ptx.cuh): the whole warp walks it, the elected lane issues; two 32-bit adds per tcgen05.mma
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc(p.n_tile);
      const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
      const uint32_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
      RingPos sr, ar;
      for (int tile = blockIdx.x; tile < tp.total_tiles; tile += gridDim.x, ar.next(acc_stages)) {
        const int acc = ar.idx;
        mbar_wait(tempty_bar(acc), ar.ph ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_tile);
        const int ks = tp.bk / 16;
        for (int kc = 0; kc < nk; kc++, sr.next(p.stages)) {
          const int s = sr.idx;
          mbar_wait(full_bar(s), sr.ph);
          tc_fence_after();
          const uint32_t a_s = base + s * stage_bytes, b_s = a_s + a_bytes;
          const uint32_t a_lo = ((a_s & 0x3FFFFu) >> 4) | (1u << 16), b_lo = ((b_s & 0x3FFFFu) >> 4) | (1u << 16);
          if (leader) {
            if (kc == 0) umma_bf16<false>(d_tmem, pack64(a_lo, hi), pack64(b_lo, hi), idesc);
            else umma_bf16<true>(d_tmem, pack64(a_lo, hi), pack64(b_lo, hi), idesc);
            for (int k = 1; k < ks; k++) umma_bf16<true>(d_tmem, pack64(a_lo + 2u * k, hi), pack64(b_lo + 2u * k, hi), idesc);
            umma_commit(empty_bar(s));
          }
        }
        if (leader) umma_commit(tfull_bar(acc));
      }
    }
    __syncwarp();
    tc_fence_before();
  } else {
    const int quarter = warp & 3, ew = warp - 2, half = ew >> 2;  // warps 2..9 -> quarters 2,3,0,1,2,3,0,1; half 0 / 1
    const int row = quarter * 32 + lane;
    uint8_t* stg = smem_raw + (base - raw) + stg_off + ew * STG_WARP;
    int* drow = reinterpret_cast<int*>(smem_raw + (base - raw) + stg_off + TMA_EPI_WARPS * STG_WARP) + ew * 32;
    // column split between the two warps of a quarter: multiples of 16, first half rounded up
    const int csplit = ((p.n_tile / 16 + 1) / 2) * 16;
    const int col_begin = half ? csplit : 0, col_end = half ? p.n_tile : csplit;
    BnAcc bn;
    bn_acc_clear(bn);
    TileDigits ti;
    ti.init((int)blockIdx.x, (int)gridDim.x, tp.tiles_n, rx_, ry_);
    RingPos ar;
    const int ry = tp.patch ? row / tp.bw : 0, rx = tp.patch ? row - ry * tp.bw : 0;  // this thread's pixel inside the patch tile
    for (int tile = blockIdx.x; tile < tp.total_tiles; tile += gridDim.x, ti.next(tp.tiles_n, rx_, ry_), ar.next(acc_stages)) {
      const int mt = ti.img, n0 = ti.nt * p.n_tile;
      int dp = -1, img = 0;
      if (tp.patch) {
        img = ti.img;
        const int oy = ti.ty * tp.bh + ry, ox = ti.tx * tp.bw + rx;  // position in the logical (hm x wm) grid
        if (ry < tp.bh && oy < p.hm && ox < p.wm) dp = (img * p.ho + p.os * oy + p.py) * p.wo + (p.os * ox + p.px);
      } else {
        const uint32_t m = (uint32_t)mt * BM + row;
        if (m < (uint32_t)(p.n * p.ho * p.wo)) { dp = (int)m; img = (int)(m / (uint32_t)(p.ho * p.wo)); }
      }
      const int acc = ar.idx;
      mbar_wait(tfull_bar(acc), ar.ph);
      tc_fence_after();
      if (!BN && tp.store_cols)
        epilogue_warp_tma(p, &tmY, tmem_base + (uint32_t)(acc * p.n_tile), quarter, lane, base + stg_off + (uint32_t)(ew * STG_WARP), stg,
                          mt * BM + quarter * 32, dp, img, n0, col_begin, col_end, tp.store_cols);
      else
        epilogue_warp<BN>(p, tmem_base + (uint32_t)(acc * p.n_tile), quarter, lane, stg, drow, dp, img, n0, col_begin, col_end, bn);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));  // all epilogue warps -> accumulator free
    }
    if (BN) bn_acc_flush(p, bn, lane, col_begin, col_end);
    if (!BN && tp.store_cols && lane == 0) tma_store_wait_read();  // shared memory must outlive the last bulk store's read
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims / strides innermost first; strides[i] (bytes) for dims 1..rank-1.
int make_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int bk = BK) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { yad_set_error("conv2d_tma: cuTensorMapEncodeTiled is unavailable"); return 1; }
  cuuint64_t d[5], st[5];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) st[i] = strides_bytes[i];
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { yad_set_error("conv2d_tma: cuTensorMapEncodeTiled failed with %d", (int)r); return 1; }
  return 0;
}

int pick_n_tile(int cout);
}  // namespace
int yad_conv2d_v2_supported(const yad_tensor* x, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y);
int yad_conv2d_v2(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream);
int yad_conv2d_dcn_v2(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream);
int yad_conv2d_v2_transposed(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream);
namespace {

// best (bw, bh) with bw * bh <= 128 for an (h, w) map: maximise useful pixels per 128-row MMA tile
void pick_patch(int h, int w, int* bw_out, int* bh_out) {
  double best = -1;
  for (int bw = 1; bw <= (w < 128 ? w : 128); bw++) {
    int bh = 128 / bw;
    if (bh > h) bh = h;
    if (bh < 1 || bh > 256) continue;
    int tx = (w + bw - 1) / bw, ty = (h + bh - 1) / bh;
    double eff = (double)h * w / ((double)tx * ty * 128.0);
    if (eff > best + 1e-9) { best = eff; *bw_out = bw; *bh_out = bh; }
  }
}

int pick_bk(int cin) { return cin <= 16 ? 16 : (cin <= 32 ? 32 : 64); }

int tma_supported(const yad_tensor* x, const yad_conv_desc* d, const yad_tensor* y) {
  if (d->mode != YAD_CONV_NORMAL) return 0;
  const int taps = d->kh * d->kw;
  if (taps > MAX_TAPS) return 0;
  if (((uintptr_t)x->ptr & 15) || (x->ld % 8) || (y->ld % 8)) return 0;
  if (x->c < 16) return 0;  // 8-channel inputs: half of every 16-wide chunk would be padding; the gather kernel packs 8 taps per chunk instead
  // kxk: every K chunk must lie inside one tap: cin is a multiple of the chunk, or a single (zero-filled) chunk covers it
  if (taps > 1 && (x->c % pick_bk(x->c)) != 0 && x->c > pick_bk(x->c)) return 0;
  if (d->stride == 1) {
    if (taps > 1 && (y->h != x->h || y->w != x->w)) return 0;
  } else if (d->stride == 2) {  // 3x3 s2 p1 on even maps stored without channel slicing: parity-split 5-D map
    if (!(d->kh == 3 && d->kw == 3 && d->pad_h == 1 && d->pad_w == 1)) return 0;
    if ((x->h & 1) || (x->w & 1) || x->ld != x->c) return 0;
    if (x->c % pick_bk(x->c)) return 0;  // the zero-filled K tail needs dim 0 == cin, which the parity-split view does not have
  } else {
    return 0;
  }
  return get_encode() != nullptr;
}

int launch_tma(TcParams& p, const yad_conv_desc* d, cudaStream_t st) {
  TmaParams tp;
  memset(&tp, 0, sizeof(tp));
  p.n_tile = pick_n_tile(p.cout);
  tp.acc_stages = (2 * p.n_tile <= 512) ? 2 : 1;
  p.tmem_cols = 32;
  while (p.tmem_cols < tp.acc_stages * p.n_tile) p.tmem_cols <<= 1;
  tp.bk = pick_bk(p.cin);
  tp.s2 = p.stride == 2;
  const int row_bytes = 2 * tp.bk;
  const int stage_bytes = ((BM + p.n_tile) * row_bytes + 1023) / 1024 * 1024;  // stages stay 1024-byte aligned for every swizzle mode
  const size_t stg = TMA_EPI_WARPS * STG_WARP + TMA_EPI_WARPS * 32 * 4;  // dedicated epilogue staging + row tables
  int stages = (int)((110 * 1024 - stg) / stage_bytes);  // aim at two resident CTAs per SM
  stages = stages > 4 ? 4 : (stages < 2 ? 2 : stages);
  p.stages = stages;
  const size_t pipe = ((size_t)stages * stage_bytes + stg + 127) / 128 * 128;
  p.pipe_bytes = (int)pipe;
  const size_t smem = 1024 + pipe + 8 * (2 * stages + 4) + 16;

  CUtensorMap tmA, tmB;
  int tiles_m;
  if (tp.s2) {
    // input pixel (2*oy + ky - 1, 2*ox + kx - 1) = parity (py, px) of half-resolution cell (oy + ay, ox + ax)
    tp.patch = 1;
    pick_patch(p.hm, p.wm, &tp.bw, &tp.bh);
    tp.tiles_x = (p.wm + tp.bw - 1) / tp.bw;
    tp.tiles_y = (p.hm + tp.bh - 1) / tp.bh;
    for (int t = 0; t < p.ntaps; t++) {
      const int qy = t / 3 - 1, qx = t % 3 - 1;
      tp.tpy[t] = qy & 1; tp.cpx[t] = (qx & 1) * p.cin;
      p.dy[t] = qy < 0 ? -1 : 0; p.dx[t] = qx < 0 ? -1 : 0;
    }
    uint64_t dims[5] = {(uint64_t)2 * p.cin, (uint64_t)p.wi / 2, 2, (uint64_t)p.hi / 2, (uint64_t)p.n};
    uint64_t strides[4] = {(uint64_t)2 * p.x_ld * 2, (uint64_t)p.wi * p.x_ld * 2, (uint64_t)2 * p.wi * p.x_ld * 2, (uint64_t)p.hi * p.wi * p.x_ld * 2};
    uint32_t box[5] = {(uint32_t)tp.bk, (uint32_t)tp.bw, 1, (uint32_t)tp.bh, 1};
    if (make_map(&tmA, p.x, 5, dims, strides, box, tp.bk)) return 1;
    tp.a_bytes = tp.bw * tp.bh * row_bytes;
    tiles_m = p.n * tp.tiles_x * tp.tiles_y;
  } else if (p.ntaps == 1 && p.os == 1 && p.dy[0] == 0 && p.dx[0] == 0) {  // 1x1: rows are consecutive pixels
    tp.patch = 0;
    const int64_t M = (int64_t)p.n * p.hi * p.wi;
    uint64_t dims[2] = {(uint64_t)p.cin, (uint64_t)M}, strides[1] = {(uint64_t)p.x_ld * 2};
    uint32_t box[2] = {(uint32_t)tp.bk, BM};
    if (make_map(&tmA, p.x, 2, dims, strides, box, tp.bk)) return 1;
    tp.a_bytes = BM * row_bytes;
    tp.tiles_x = tp.tiles_y = 1;
    tiles_m = (int)((M + BM - 1) / BM);
  } else {
    tp.patch = 1;
    pick_patch(p.hm, p.wm, &tp.bw, &tp.bh);
    tp.tiles_x = (p.wm + tp.bw - 1) / tp.bw;
    tp.tiles_y = (p.hm + tp.bh - 1) / tp.bh;
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)p.wi, (uint64_t)p.hi, (uint64_t)p.n};
    uint64_t strides[3] = {(uint64_t)p.x_ld * 2, (uint64_t)p.wi * p.x_ld * 2, (uint64_t)p.hi * p.wi * p.x_ld * 2};
    uint32_t box[4] = {(uint32_t)tp.bk, (uint32_t)tp.bw, (uint32_t)tp.bh, 1};
    if (make_map(&tmA, p.x, 4, dims, strides, box, tp.bk)) return 1;
    tp.a_bytes = tp.bw * tp.bh * row_bytes;
    tiles_m = p.n * tp.tiles_x * tp.tiles_y;
  }
  {
    const int cout_rows = (p.cout + 7) / 8 * 8;
    uint64_t dims[2] = {(uint64_t)p.w_row, (uint64_t)cout_rows}, strides[1] = {(uint64_t)p.w_row * 2};
    uint32_t box[2] = {(uint32_t)tp.bk, (uint32_t)p.n_tile};
    if (make_map(&tmB, p.w, 2, dims, strides, box, tp.bk)) return 1;
  }
  tp.tiles_n = (p.cout + p.n_tile - 1) / p.n_tile;
  tp.total_tiles = tiles_m * tp.tiles_n;
  CUtensorMap tmY;
  memset(&tmY, 0, sizeof(tmY));
  {
    static int store_env = -1;
    if (store_env < 0) { const char* ev = getenv("YAD_CONV_TMA_STORE"); store_env = (ev && ev[0] == '0') ? 0 : 1; }
    const int csplit = ((p.n_tile / 16 + 1) / 2) * 16;  // the column split of the two warps of a TMEM lane quarter (conv_tma_kernel)
    const int sc = csplit > 64 ? 64 : csplit;
    if (store_env && !tp.patch && !p.bn_stats && !p.out_f32 && (sc == 16 || sc == 32 || sc == 64) && csplit % sc == 0 &&
        p.n_tile - csplit == csplit && ((uintptr_t)p.y & 15) == 0) {
      const int64_t M = (int64_t)p.n * p.ho * p.wo;
      uint64_t dims[2] = {(uint64_t)p.cout, (uint64_t)M}, strides[1] = {(uint64_t)p.y_ld * 2};
      uint32_t box[2] = {(uint32_t)sc, 32};
      if (make_map(&tmY, p.y, 2, dims, strides, box, sc)) return 1;
      tp.store_cols = sc;
    }
  }
  tp.p = p;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaFuncSetAttribute(conv_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      yad_set_error("conv2d_tma: cannot raise the dynamic shared memory limit");
      num_sms = 0;
      return 2;
    }
  }
  // resident CTAs per SM: shared memory and TMEM columns (512 per SM)
  int per_sm = (int)((227 * 1024) / smem);
  const int by_tmem = 512 / p.tmem_cols;
  if (per_sm > by_tmem) per_sm = by_tmem;
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sms * per_sm;
  if (grid > tp.total_tiles) grid = tp.total_tiles;
  if (p.bn_stats)
    YAD_LAUNCH(conv_tma_kernel<true>, grid, TMA_THREADS, smem, st, tp, tmA, tmB, tmY);
  else
    YAD_LAUNCH(conv_tma_kernel<false>, grid, TMA_THREADS, smem, st, tp, tmA, tmB, tmY);
  YAD_LAUNCH_CHECK("conv2d_tma");
  (void)d;
  return 0;
}

int pick_n_tile(int cout) {
  int c16 = (cout + 15) / 16 * 16;
  if (c16 <= 256) return c16;
  // split evenly into tiles of at most 256 channels (multiples of 16)
  int tiles = (c16 + 255) / 256;
  int nt = ((c16 + tiles - 1) / tiles + 15) / 16 * 16;
  return nt;
}

int launch(TcParams& p, int64_t M, cudaStream_t st) {
  if (M + BM >= (int64_t)1 << 31 || (int64_t)p.n * p.ho * p.wo >= (int64_t)1 << 31) {
    yad_set_error("conv2d_tc: more than 2^31 pixels");
    return 1;
  }
  p.n_tile = pick_n_tile(p.cout);
  p.tmem_cols = 32;
  while (p.tmem_cols < p.n_tile) p.tmem_cols <<= 1;
  const int stage_bytes = BM * 128 + p.n_tile * 128;
  const int K = p.ntaps * p.cin, nk = (K + BK - 1) / BK;
  int stages = (96 * 1024) / stage_bytes;
  stages = stages > 4 ? 4 : (stages < 2 ? 2 : stages);
  if (p.deform && stages > 2) stages = 2;  // keep three CTAs per SM resident next to the sampling-parameter table
  if (stages > nk) stages = nk < 1 ? 1 : nk;
  p.stages = stages;
  size_t pipe = (size_t)stages * stage_bytes;
  const size_t stg = 4 * STG_WARP + 4 * 32 * 4;  // epilogue staging (aliased onto the pipeline buffers) + row table
  if (pipe < stg) pipe = (stg + 127) / 128 * 128;
  p.pipe_bytes = (int)pipe;
  const size_t smem = 1024 + pipe + 8 * (2 * stages + 1) + 16 + (p.deform ? 32 + BM * MAX_TAPS * 20 : 0);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      yad_set_error("conv2d_tc: cannot raise the dynamic shared memory limit");
      return 2;
    }
    attr_set = true;
  }
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((p.cout + p.n_tile - 1) / p.n_tile));
  if (p.deform)
    YAD_LAUNCH(conv_tc_kernel<true>, grid, NTHREADS, smem, st, p);
  else
    YAD_LAUNCH(conv_tc_kernel<false>, grid, NTHREADS, smem, st, p);
  YAD_LAUNCH_CHECK("conv2d_tc");
  return 0;
}

}  // namespace

int yad_conv2d_tc_supported(const yad_tensor* x, const yad_conv_desc* d, const yad_tensor* y) {
  if (x->c % 8 || x->ld % 8 || y->c % 8 || y->ld % 8) return 0;
  if (((uintptr_t)x->ptr & 15) || ((uintptr_t)y->ptr & 15)) return 0;
  if (d->kh * d->kw > MAX_TAPS) return 0;
  if (d->mode == YAD_CONV_DEFORM && (x->c % BK) != 0) return 0;
  return 1;
}

int yad_conv2d_tc(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const bf16*)x->ptr; p.w = (const bf16*)w; p.om = (const bf16*)d->offmask; p.y = y->ptr;
  p.n = x->n; p.hi = x->h; p.wi = x->w; p.cin = x->c; p.x_ld = x->ld;
  p.ho = y->h; p.wo = y->w; p.cout = y->c; p.y_ld = y->ld;
  p.w_row = d->kh * d->kw * x->c;
  p.om_ld = d->offmask_ld;
  p.e = *e;
  p.cout_real = y->c;
  YAD_CHECK(x->n == y->n, "conv2d: batch mismatch %d vs %d", x->n, y->n);
  cudaStream_t st = (cudaStream_t)stream;
  if (e->gn_stats && e->gn_groups == 0) {  // train-mode BatchNorm: per-channel statistics over the whole batch, double [cout][2]
    YAD_CHECK(d->mode != YAD_CONV_TRANSPOSED && !e->mul && !e->add, "conv2d: fused BatchNorm statistics: plain / deformable convolutions without mul / add");
    YAD_CHECK(y->c <= 512, "conv2d: fused BatchNorm statistics support up to 512 output channels");
    p.bn_stats = 1;
    cudaMemsetAsync(e->gn_stats, 0, sizeof(double) * 2 * y->c, st);
  } else if (e->gn_stats) {
    const int cpg = e->gn_groups > 0 ? y->c / e->gn_groups : 0;
    YAD_CHECK(e->gn_groups > 0 && y->c % e->gn_groups == 0 && (cpg == 4 || cpg == 8 || cpg == 16),
              "conv2d: fused GroupNorm statistics need 4, 8 or 16 channels per group (got %d channels / %d groups)", y->c, e->gn_groups);
    YAD_CHECK(d->mode != YAD_CONV_TRANSPOSED, "conv2d: fused GroupNorm statistics are not built for the transposed mode");
    cudaMemsetAsync(e->gn_stats, 0, sizeof(double) * 2 * e->gn_groups * y->n, st);
  }
  if (d->mode == YAD_CONV_NORMAL || d->mode == YAD_CONV_DEFORM) {
    if (d->mode == YAD_CONV_NORMAL) {
      YAD_CHECK(y->h == (x->h + 2 * d->pad_h - d->kh) / d->stride + 1 && y->w == (x->w + 2 * d->pad_w - d->kw) / d->stride + 1,
                "conv2d: output shape %dx%d does not match input %dx%d k%dx%d s%d p%d,%d", y->h, y->w, x->h, x->w, d->kh, d->kw, d->stride,
                d->pad_h, d->pad_w);
    } else {
      YAD_CHECK(d->kh == 3 && d->kw == 3 && d->stride == 1 && d->pad_h == 1 && d->pad_w == 1 && y->h == x->h && y->w == x->w && d->offmask,
                "conv2d: deformable mode is 3x3 s1 p1 with an offset/mask view");
      p.deform = 1;
    }
    p.hm = y->h; p.wm = y->w; p.stride = d->stride; p.os = 1; p.py = 0; p.px = 0;
    p.ntaps = d->kh * d->kw;
    for (int t = 0; t < p.ntaps; t++) { p.dy[t] = t / d->kw - d->pad_h; p.dx[t] = t % d->kw - d->pad_w; p.wtap[t] = t; }
    // impl 4: the resident-weight kernel of conv_v2.cu (error when the shape is not eligible); 5: conv_tma_kernel (the round-1 TMA kernel, A/B runs)
    if (d->mode == YAD_CONV_NORMAL && d->impl != 3 && d->impl != 5 && !p.bn_stats && yad_conv2d_v2_supported(x, d, e, y))
      return yad_conv2d_v2(x, w, d, e, y, stream);
    if (d->mode == YAD_CONV_DEFORM && d->impl != 3 && d->impl != 5 && !p.bn_stats) {  // fused shared-memory gather kernel (conv_v2.cu)
      const int r = yad_conv2d_dcn_v2(x, w, d, e, y, stream);
      if (r >= 0) return r;
    }
    YAD_CHECK(d->impl != 4, "conv2d: impl 4 (resident-weight tcgen05 kernel) does not support this shape / epilogue");
    YAD_CHECK(!e->gate_h, "conv2d: the separable gate epilogue runs on conv2_kernel (bf16: channels and strides multiples of 8, cin >= 16) or the SIMT kernel only");
    if (d->impl != 3 && tma_supported(x, d, y)) return launch_tma(p, d, st);
    if (p.bn_stats) {  // batch statistics are fused on the TMA-fed kernel only: thread-gathered kernel first, stand-alone statistics after it
      p.bn_stats = 0;
      p.e.gn_stats = nullptr;
      const int r = launch(p, (int64_t)x->n * p.hm * p.wm, st);
      if (r) return r;
      yad_tensor flat = *y;
      flat.n = 1;
      flat.h = y->n * y->h;
      return yad_gn_stats(&flat, y->c, e->gn_stats, YAD_BF16, stream);
    }
    return launch(p, (int64_t)x->n * p.hm * p.wm, st);
  }
  if (d->mode == YAD_CONV_TRANSPOSED) {
    YAD_CHECK(d->kh == 3 && d->kw == 3 && d->stride == 2 && d->pad_h == 1 && d->pad_w == 1 && y->h == 2 * x->h && y->w == 2 * x->w,
              "conv2d: transposed mode is k3 s2 p1 op1 only");
    if (d->impl != 3 && d->impl != 5 && !p.bn_stats) {  // the four phases on conv2_kernel (haloed patch, resident taps, strided TMA store)
      const int r = yad_conv2d_v2_transposed(x, w, d, e, y, stream);
      if (r >= 0) return r;
      YAD_CHECK(d->impl != 4, "conv2d: impl 4 (resident-weight tcgen05 kernel) does not support this transposed shape / epilogue");
    }
    // oy = 2*iy - 1 + ky: even rows take ky = 1 (iy = m); odd rows take ky = 0 (iy = m + 1) and ky = 2 (iy = m)
    p.hm = x->h; p.wm = x->w; p.stride = 1; p.os = 2;
    for (int py = 0; py < 2; py++)
      for (int px = 0; px < 2; px++) {
        p.py = py; p.px = px; p.ntaps = 0;
        for (int ky = 0; ky < 3; ky++) {
          if (((py + 1 - ky) & 1) != 0) continue;
          for (int kx = 0; kx < 3; kx++) {
            if (((px + 1 - kx) & 1) != 0) continue;
            p.dy[p.ntaps] = (py + 1 - ky) / 2; p.dx[p.ntaps] = (px + 1 - kx) / 2; p.wtap[p.ntaps] = ky * 3 + kx;
            p.ntaps++;
          }
        }
        // each phase is a stride-1 convolution over the input grid with 1, 2 or 4 taps: TMA-eligible when cin % 64 == 0
        const bool tma_ok = d->impl != 3 && (x->c % BK) == 0 && !((uintptr_t)x->ptr & 15) && get_encode() != nullptr;
        int r = tma_ok ? launch_tma(p, d, st) : launch(p, (int64_t)x->n * p.hm * p.wm, st);
        if (r) return r;
      }
    return 0;
  }
  YAD_CHECK(false, "conv2d: bad mode %d", d->mode);
}

// C[M][N] (fp32) = A[M][K] (bf16 row-major) x B[N][K]^T (bf16) through the UMMA/TMEM path: validates the descriptor encodings, the
// swizzled staging and the TMEM read-back independently of the convolution address logic.
extern "C" int yad_tc_gemm_selftest(const void* a, const void* b, float* c, int m, int n, int k, void* stream) {
  YAD_CHECK(k % 8 == 0 && n % 8 == 0, "tc_gemm_selftest: k and n must be multiples of 8");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const bf16*)a; p.w = (const bf16*)b; p.y = c;
  p.n = 1; p.hi = m; p.wi = 1; p.cin = k; p.x_ld = k;
  p.hm = m; p.wm = 1; p.ho = m; p.wo = 1; p.cout = n; p.y_ld = n;
  p.stride = 1; p.os = 1; p.ntaps = 1; p.w_row = k; p.out_f32 = 1;
  p.e.alpha = 1.0f;
  return launch(p, m, (cudaStream_t)stream);
}
