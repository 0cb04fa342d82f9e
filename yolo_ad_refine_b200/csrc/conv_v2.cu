// Second-generation tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation): stride-1 1x1 and 3x3 (pad 1)
// convolutions of yad_conv2d's NORMAL mode, the shapes that carry the inference step (nn/modules/conv.py:36-54 Conv.forward_fuse, the Conv_GN / 1x1
// convolutions of nn/modules/head.py:1131-1175, the neck's lateral 1x1 of the yaml).  What changed against conv_tma_kernel (conv_tc.cu), and why
// (profiles/r1_ncu_conv_s6.json: L2->SM traffic 13.6x the unique input on 3x3, 60 % integer / control instructions in the epilogue on 1x1):
//
//   * WEIGHTS ARE RESIDENT.  A persistent CTA loads the whole packed weight matrix [n tiles][taps][K chunks] once (one TMA box and one mbarrier per
//     64-wide K chunk, so the first MMAs start as soon as the first chunk lands) and keeps it in shared memory for all of its tiles; the pipeline
//     ring carries activations only.
//   * 3x3: ONE HALOED PATCH PER K CHUNK.  The M tile is an 8 (x) by 16 (y) pixel rectangle of one image; the producer fetches the (8+2) x (16+2)
//     pixel patch x 64 channels with a single 4-D box (conv padding = TMA zero fill) and the nine taps are nine UMMA A-descriptors into that one
//     patch: start address shifted by ((dy+1) * pitch + (dx+1)) pixel rows of 128 bytes, stride-byte-offset = one patch row (every 8 consecutive
//     M rows are the 8 pixels of one tile row), matrix-base-offset = (start >> 7) & 7 as the PTX ISA prescribes for SWIZZLE_128B operands that do
//     not start on a 1024-byte boundary.  L2->SM traffic per tile drops from 9 A boxes + 9 B boxes to 1.4x the tile's own input.
//   * LEAN EPILOGUE.  Activation, mul/add, GroupNorm statistics and per-row scales are template parameters; the bias sits in shared memory; one
//     thread = one accumulator row: tcgen05.ld 32 columns, bias + activation, pack to bf16, 128-bit conflict-free stores into the swizzled image
//     of a TMA store box, one bulk tensor store per (warp, piece) -- 2-D box for 1x1 (rows = consecutive pixels), 4-D box for 3x3 (4 tile rows of
//     8 pixels; the unit clips ragged tiles).  The TMEM accumulator is released right after the last tcgen05.ld of a tile, before the stores.
//     mul / add operands are applied on the staged tile in packed bf16x2 (row-contiguous 128-bit loads).  GroupNorm partial sums leave through a
//     butterfly reduce-scatter (16 shuffles per 32 columns instead of 80) and one fp64 atomic per (group, statistic) and warp.
//
// Always SWIZZLE_128B with 64-channel (128-byte) rows: narrower inputs (cin 16 .. 48, K tails) are zero-filled by the TMA unit and only the
// ceil(cin / 16) K steps that carry data are issued.
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int V2_BM = 128;
constexpr int V2_MAX_EPI_WARPS = 16;  // 8 (two per TMEM lane quarter) or 16 (four per quarter: twice the warps to hide the epilogue's latencies)
constexpr int V2_MAX_THREADS = 64 + 32 * V2_MAX_EPI_WARPS;
constexpr int V2_BH = 16, V2_BW = 8;   // 3x3 M tile: 16 rows of 8 pixels
constexpr int V2_MAX_BCHUNKS = 32;     // resident weight chunks (one mbarrier each)
constexpr int V2_MAX_ACC = 8;          // TMEM accumulator stages (tile-split epilogue of narrow tiles: 8 x 64 columns)
constexpr int ACT_GENERIC = -1;        // activation chosen at run time (relu / gelu / hardswish)

struct V2Params {
  int n, hm, wm, hw;        // output grid per image (stride 1: = input grid)
  int cin, cout;
  int m_total;              // n * hm * wm
  int n_tile, tiles_n, total_tiles;
  int tiles_x, tiles_y;     // patch mode: tiles per image
  int kpt, ntaps, pw;       // 64-wide K chunks per tap; taps; patch pitch in pixels
  int tap_row[9];           // patch row (= pixel index inside the patch) the A descriptor of tap t starts at
  int custom;               // 1: tap list given at run time (a phase of the stride-2 transposed convolution: 1, 2 or 4 taps with offsets 0 / +1);
                            // the patch then starts AT the tile's first pixel (org = 0) and weight tap t sits at K offset wtap[t] * cin
  int org;                  // the haloed patch starts org pixels left of / above the tile (1 for the 3x3 pad-1 convolution)
  int wtap[9];
  int stages, acc_stages, tmem_cols;
  int dbg;                  // bring-up / profiling switch (YAD_CONV2_DBG): 1 = the MMA lane issues nothing, 2 = the epilogue skips its TMEM reads and math (results are garbage)
  int tsplit;               // narrow tiles (n_tile <= 64): number of epilogue warp groups (4 warps = the 4 TMEM lane quarters) that take whole tiles in
                            // turn instead of sharing the columns of one tile (0: column split); several tiles' epilogue latencies then overlap
  int ksplit;               // 3x3: taps are dealt round-robin to `ksplit` accumulators (independent tcgen05.mma dependency chains), summed in the epilogue
  uint32_t off_b, b_chunk_bytes, off_a, a_stage_bytes, a_tx_bytes, off_stg, stg_warp_bytes, off_bias, off_bars;
  uint32_t off_pf, pf_warp_bytes;  // operand-row staging (cp.async): per epilogue warp 2 KB per operand (4 rows x 32 lanes x 16 bytes); 0 = not allocated
  int sc, ew, stg_bufs;     // columns per TMA store box (16 / 32 / 64); epilogue warps (8 / 16); staging tiles per warp (2: the bulk store of piece i
                            // is still reading its tile while piece i + 1 is staged)
  int bnd[5];               // column ranges [bnd[w], bnd[w + 1]) of the ew / 4 warps that share a TMEM lane quarter
  const float* bias;
  const float* img_scale;
  const bf16* pix_scale;
  int pix_scale_ld;
  int act;
  float alpha;
  const bf16* mul;
  int mul_ld;
  const bf16* add;
  int add_ld;
  const bf16* gate_h;  // separable gate rows (n, hm, cout) / (n, wm, cout): y *= round_bf16(gate_h * gate_w), flat (1x1) tiles only
  const bf16* gate_w;
  int gate_ld;
  double* gn_stats;
  int gn_groups, cpg;
};

// K-major SWIZZLE_128B matrix descriptor with an explicit stride-byte-offset (distance between 8-row groups) and the matrix base offset for
// start addresses that are 128-byte but not 1024-byte aligned (tap-shifted views of the haloed patch)
// The matrix-base-offset field (bits 49-51) stays 0: measured on B200 (tools/v2_tap_debug.py), the tensor core applies the 128-byte swizzle to the
// ABSOLUTE shared-memory address bits [7, 10) -> [4, 7), exactly as the TMA unit does when it writes the box, so a descriptor may start on any
// 128-byte row of a TMA-written tile and use any multiple of 128 bytes as its stride-byte-offset (1280 = a 10-pixel patch row works); setting the
// field to (start >> 7) & 7 breaks every start that is not 1024-byte aligned.
// The descriptor is kept as two 32-bit words so that the per-MMA arithmetic (K advance, tap shift, weight chunk) is one 32-bit add on the low word.
__device__ __forceinline__ uint32_t v2_desc_hi(uint32_t sbo) { return (sbo >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t v2_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t v2_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(V2_BM >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2),
               "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(tm),
               "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ float v2_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint4 lds16(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts16v(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}

// Butterfly reduce-scatter of NV per-thread partial sums over the 32 lanes of a warp: afterwards a[0] of lane l holds the warp total of value
// index l >> (5 - log2 NV); NV in {2, 4, 8, 16}.  NV - 1 + (5 - log2 NV) shuffles instead of 5 * NV.
template <int W>
__device__ __forceinline__ void rs_step(float* a, int lane, int off) {
  if constexpr (W > 1) {
    constexpr int H = W / 2;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
      const float send = up ? a[i] : a[i + H];
      const float keep = up ? a[i + H] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  } else {
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], off);
  }
}
template <int NV>
__device__ __forceinline__ void reduce_scatter(float (&a)[NV], int lane) {
  rs_step<NV>(a, lane, 16);
  rs_step<(NV > 1 ? NV / 2 : 1)>(a, lane, 8);
  rs_step<(NV > 2 ? NV / 4 : 1)>(a, lane, 4);
  rs_step<(NV > 4 ? NV / 8 : 1)>(a, lane, 2);
  rs_step<(NV > 8 ? NV / 16 : 1)>(a, lane, 1);
}
template <int U, int CPG>
__device__ __forceinline__ void gn_unit(const V2Params& p, const float (&v)[U], bool valid, int lane, int img, bool uniform, int co) {
  constexpr int NG = U / CPG, NV = 2 * NG;
  float a[NV];
#pragma unroll
  for (int g = 0; g < NG; g++) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < CPG; i++) { const float x = valid ? v[g * CPG + i] : 0.f; s += x; q = fmaf(x, x, q); }
    a[2 * g] = s; a[2 * g + 1] = q;
  }
  const int g0 = co / CPG;
  if (uniform) {  // warp-uniform: all valid rows of this warp belong to image `img`
    reduce_scatter<NV>(a, lane);
    constexpr int SH = NV == 16 ? 1 : (NV == 8 ? 2 : (NV == 4 ? 3 : 4));
    const int idx = lane >> SH;
    if ((lane & ((1 << SH) - 1)) == 0 && g0 + (idx >> 1) < p.gn_groups) atomicAdd(&p.gn_stats[((int64_t)img * p.gn_groups + g0) * 2 + idx], (double)a[0]);
  } else if (valid) {
#pragma unroll
    for (int i = 0; i < NV; i++)
      if (g0 + (i >> 1) < p.gn_groups) atomicAdd(&p.gn_stats[((int64_t)img * p.gn_groups + g0) * 2 + i], (double)a[i]);
  }
}

// Per-tile bookkeeping without divisions: every role (producer, MMA issuer, epilogue warps) walks tile, tile + gridDim.x, ... and used to
// decompose each tile index with 4-6 integer divisions (~140 instructions and ~1000 cycles of dependent latency per tile in the MMA warp alone:
// the tensor pipe idled through it, profiles/r2_ncu_conv_v2.md).  The decomposition is now carried as mixed-radix digits
// (column tile, tile x, tile y, image) advanced by the digits of the grid stride; flat (1x1) launches have tiles_x = tiles_y = 1, so `img` is
// the M tile index there.  V2Ring is a ring-buffer slot with its mbarrier phase.
struct V2TileIter {
  int nt, tx, ty, img, d_nt, d_tx, d_ty, d_img;
  __device__ __forceinline__ void init(const V2Params& p, int tile0, int step) {
    nt = tile0 % p.tiles_n; int mt = tile0 / p.tiles_n;
    tx = mt % p.tiles_x; mt /= p.tiles_x;
    ty = mt % p.tiles_y; img = mt / p.tiles_y;
    d_nt = step % p.tiles_n; mt = step / p.tiles_n;
    d_tx = mt % p.tiles_x; mt /= p.tiles_x;
    d_ty = mt % p.tiles_y; d_img = mt / p.tiles_y;
  }
  template <bool PATCH = true>
  __device__ __forceinline__ void next(const V2Params& p) {
    nt += d_nt; int c = nt >= p.tiles_n ? 1 : 0; nt -= c ? p.tiles_n : 0;
    if constexpr (PATCH) {
      tx += d_tx + c; c = tx >= p.tiles_x ? 1 : 0; tx -= c ? p.tiles_x : 0;
      ty += d_ty + c; c = ty >= p.tiles_y ? 1 : 0; ty -= c ? p.tiles_y : 0;
    }  // flat tiles: tiles_x = tiles_y = 1, tx = ty = d_tx = d_ty = 0 and the carry goes straight to the M-tile index
    img += d_img + c;
  }
};
struct V2Ring {
  int idx = 0;
  uint32_t ph = 0;
  __device__ __forceinline__ void next(int n) { if (++idx == n) { idx = 0; ph ^= 1u; } }
  __device__ __forceinline__ void advance(int k, int n) {  // k steps at once (k >= 0)
    idx += k;
    while (idx >= n) { idx -= n; ph ^= 1u; }
  }
};

// One unit of U (16 / 32) accumulator columns of this thread's row: TMEM -> registers -> scale / bias / activation / alpha (-> GroupNorm partial
// sums) -> bf16 -> swizzled staging row.  `first` : the staging tile is about to be overwritten for the first time since the last bulk store.
template <int U, int ACT, bool GN, bool SCALE>
__device__ __forceinline__ void epi_unit(const V2Params& p, uint32_t taddr, uint32_t bias_addr, float rsc, uint32_t row_addr, uint32_t ph, int cell0,
                                         bool first, int lane, bool valid, int img, bool uniform, int co) {
  uint32_t r[U];
  tmem_ld16_nowait(taddr, r);
  if constexpr (U == 32) tmem_ld16_nowait(taddr + 16u, r + 16);
  tmem_ld_wait();
  for (int sp = 1; sp < p.ksplit; sp++) {  // partial sums of the other tap groups (warp-uniform trip count)
    uint32_t r2[U];
    tmem_ld16_nowait(taddr + (uint32_t)(sp * p.n_tile), r2);
    if constexpr (U == 32) tmem_ld16_nowait(taddr + (uint32_t)(sp * p.n_tile) + 16u, r2 + 16);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < U; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
  }
  float v[U];
#pragma unroll
  for (int i = 0; i < U; i += 4) {
    const float4 b = lds_f4(bias_addr + 4u * (uint32_t)i);
    if constexpr (SCALE) {
      v[i] = fmaf(__uint_as_float(r[i]), rsc, b.x); v[i + 1] = fmaf(__uint_as_float(r[i + 1]), rsc, b.y);
      v[i + 2] = fmaf(__uint_as_float(r[i + 2]), rsc, b.z); v[i + 3] = fmaf(__uint_as_float(r[i + 3]), rsc, b.w);
    } else {
      v[i] = __uint_as_float(r[i]) + b.x; v[i + 1] = __uint_as_float(r[i + 1]) + b.y;
      v[i + 2] = __uint_as_float(r[i + 2]) + b.z; v[i + 3] = __uint_as_float(r[i + 3]) + b.w;
    }
  }
  if constexpr (ACT == YAD_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < U; i++) { const float h = 0.5f * v[i]; v[i] = fmaf(h, v2_tanh(h), h); }
  } else if constexpr (ACT == YAD_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < U; i++) v[i] = fmaf(0.5f, v2_tanh(0.5f * v[i]), 0.5f);
  } else if constexpr (ACT == YAD_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < U; i++) v[i] = fmaxf(v[i], 0.0f);
  } else if constexpr (ACT == ACT_GENERIC) {
    if (p.act == YAD_ACT_SILU) {
#pragma unroll
      for (int i = 0; i < U; i++) { const float h = 0.5f * v[i]; v[i] = fmaf(h, v2_tanh(h), h); }
    } else if (p.act == YAD_ACT_SIGMOID) {
#pragma unroll
      for (int i = 0; i < U; i++) v[i] = fmaf(0.5f, v2_tanh(0.5f * v[i]), 0.5f);
    } else if (p.act != YAD_ACT_NONE) {
      apply_act_n<U>(v, p.act);
    }
  }
  if (p.alpha != 1.0f) {
#pragma unroll
    for (int i = 0; i < U; i++) v[i] *= p.alpha;
  }
  if constexpr (GN) {
    if (p.gn_stats) {
      if (p.cpg == 4) gn_unit<U, 4>(p, v, valid, lane, img, uniform, co);
      else if (p.cpg == 8) gn_unit<U, 8>(p, v, valid, lane, img, uniform, co);
      else gn_unit<U, 16>(p, v, valid, lane, img, uniform, co);
    }
  }
  if (first) {  // the staging tile about to be overwritten: its last bulk store (one or two pieces ago) must have finished reading it
    if (elect_one()) {  // the same lane issues, commits and waits
      if (p.stg_bufs == 2) tma_store_wait_read1();
      else tma_store_wait_read();
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < U / 8; j++) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
    sts16v(row_addr + ((((uint32_t)(cell0 + j)) ^ ph) << 4), u);
  }
}

// The epilogue of one warp over all tiles of its CTA (shared by conv2_kernel and dcn2_kernel): TMEM -> scale / bias / activation (-> GroupNorm
// partial sums) -> bf16 -> swizzled staging tile -> optional mul / add on the staged tile -> bulk tensor store.  tfull0 / tempty0: shared-memory
// addresses of the accumulator-full / accumulator-empty mbarrier pairs.
template <bool PATCH, int ACT, bool MULADD, bool GN, bool SCALE, bool PRE = false>
__device__ __forceinline__ void v2_epilogue(const V2Params& p, const CUtensorMap* tmY, uint32_t base, uint32_t tmem_base, int warp, int lane, uint32_t tfull0,
                                            uint32_t tempty0) {
  const int per_img = p.tiles_x * p.tiles_y;
  {
    // ================= epilogue: 8 or 16 warps, TMEM lane quarter = warp & 3, the 2 / 4 warps of a quarter split the columns =================
    const int q = warp & 3, ew = warp - 2, way = ew >> 2;
    const int G = p.tsplit;
    const int cb = G ? 0 : p.bnd[way], ce = G ? p.n_tile : p.bnd[way + 1];
    const uint32_t stg0 = base + p.off_stg + (uint32_t)(ew * p.stg_bufs) * p.stg_warp_bytes;
    const uint32_t RB = 2u * (uint32_t)p.sc, swz = (RB >> 4) - 1u;
    const uint32_t ph = (((uint32_t)lane * RB) >> 7) & swz;  // every staging tile starts on its swizzle period
    uint32_t piece = 0;
    const uint32_t bias_s = base + p.off_bias;
    const uint32_t lane_base = ((uint32_t)(q * 32)) << 16;
    // phase-2 geometry (mul / add on the staged piece): cpr lanes sweep one row
    const int cpr = p.sc >> 3, rpp = 32 / cpr, rr0 = lane / cpr, cj = lane - rr0 * cpr;
    // a lane's consecutive operand rows (phase 2, cpr <= 4): rpp pixels apart in a flat tile; rpp / 8 image rows apart (same column) in a patch tile
    const int rstep = PATCH ? (rpp >> 3) * p.wm : rpp;              // in pixels
    const int rsh = 31 - __clz(PATCH ? (rpp >> 3 ? rpp >> 3 : 1) : rpp);  // log2 of the row step in its own unit (image rows / pixels); powers of two
    const int step_a = rstep * p.add_ld, step_m = rstep * p.mul_ld;  // in elements
    const bool pf_smem = MULADD && p.pf_warp_bytes != 0u && cpr <= 4 && !(p.dbg & 8);
    const uint32_t pf_lane = base + p.off_pf + (uint32_t)ew * p.pf_warp_bytes + (uint32_t)lane * 16u, pf_mul = p.add ? 2048u : 0u;
    // tile-split: warp group `way` takes the CTA's tiles way, way + G, way + 2 G, ...: its iterator and its accumulator ring step G tiles at a time
    // (walking every tile and skipping G - 1 of G cost ~60 instructions per skipped tile -- tile iterator + ring with their spilled state --
    // which for G = 4 was a third of the instructions of this issue-bound epilogue; ncu of conv3_kernel: 15 % of all warp instructions for G = 2)
    const int g1 = G ? G : 1, first_tile = (int)blockIdx.x + (G ? way : 0) * (int)gridDim.x;
    V2TileIter ti;
    ti.init(p, first_tile, g1 * (int)gridDim.x);
    V2Ring ar;     // accumulator stage of the current tile
    ar.advance(G ? way : 0, p.acc_stages);
    for (int tile = first_tile; tile < p.total_tiles; tile += g1 * (int)gridDim.x, ti.template next<PATCH>(p), ar.advance(g1, p.acc_stages)) {
      const int n0 = ti.nt * p.n_tile;
      int img = 0, ty0 = 0, tx0 = 0, m_base = 0;
      bool valid, uniform = true;
      int dp = 0;
      if (PATCH) {
        img = ti.img;
        ty0 = ti.ty * V2_BH;
        tx0 = ti.tx * V2_BW;
        const int oy = ty0 + 4 * q + (lane >> 3), ox = tx0 + (lane & 7);
        valid = oy < p.hm && ox < p.wm;
        dp = (img * p.hm + oy) * p.wm + ox;
      } else {
        m_base = ti.img * V2_BM + 32 * q;
        dp = m_base + lane;
        valid = dp < p.m_total;
        if (GN || SCALE) {
          const int img0 = m_base / p.hw, rem = m_base - img0 * p.hw;
          uniform = rem + 32 <= p.hw;
          img = img0 + ((rem + lane >= p.hw) ? 1 : 0);
        }
      }
      float rsc = 1.0f;
      if constexpr (SCALE) {
        if (valid) {
          if (p.img_scale) rsc = p.img_scale[img];
          if (p.pix_scale) rsc *= __bfloat162float(p.pix_scale[(int64_t)dp * p.pix_scale_ld]);
        }
      }
      // The mul / add rows of a piece are pulled into L2 by prefetch hints BEFORE its accumulator is read (for the first piece: before the wait for
      // the accumulator), so their DRAM latency runs under the MMAs / the TMEM reads instead of inside the per-tile epilogue chain (the bottleneck
      // adds doubled conv3_kernel's time: 37 -> 65 us at 8 -> 16 @160^2; the lateral 1x1 with Multiply + Add ran at 115 us against 41 us without
      // them).  Phase 2 then reads all rows of the piece at once, in the registers the accumulator has just left.  Holding the operands in registers
      // across the TMEM reads does not fit the 96-register budget of the 576-thread CTA: ptxas spilled them right behind the loads, which stalls
      // exactly where the prefetch was meant to run ahead.  YAD_CONV2_PREFETCH=0 (p.dbg & 8) keeps the old two-rows-at-a-time loop.
      constexpr bool PREFETCH = MULADD && !GN && !SCALE;
      const bool pre = PREFETCH && (p.add || p.mul || p.gate_h) && cpr <= 4 && !(p.dbg & 8);
      // separable gate of a flat (1x1) tile: (image, y, x) of this warp's first row once per tile (two integer divisions), rows step from there
      int g_img = 0, g_y = 0, g_x = 0;
      if constexpr (MULADD && !PATCH) {
        if (p.gate_h) {
          g_img = m_base / p.hw;
          const int rem = m_base - g_img * p.hw;
          g_y = rem / p.wm;
          g_x = rem - g_y * p.wm;
        }
      }
      // addresses of the gate rows of row `row` (which must be inside the map), 8 channels from co
      auto gate_addr = [&](int row, int co, const uint4*& ah, const uint4*& aw) {
        int gi = g_img, gy = g_y, gx = g_x + row;
        while (gx >= p.wm) { gx -= p.wm; gy++; }
        while (gy >= p.hm) { gy -= p.hm; gi++; }
        ah = reinterpret_cast<const uint4*>(p.gate_h + (int64_t)(gi * p.hm + gy) * p.gate_ld + co);
        aw = reinterpret_cast<const uint4*>(p.gate_w + (int64_t)(gi * p.wm + gx) * p.gate_ld + co);
      };
      // round_bf16(gate_h * gate_w): exactly the values yad_rowcol_gate would have written into a gate map, so the tile below is bit-identical to
      // the mul-operand path
      auto gate_mul = [&](const uint4& a, const uint4& b) -> uint4 {
        uint4 r;
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
        __nv_bfloat162* hr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int e = 0; e < 4; e++) hr[e] = __hmul2(ha[e], hb[e]);
        return r;
      };
      auto gate_vec = [&](int row, int co) -> uint4 {
        const uint4 *ah, *aw;
        gate_addr(row, co, ah, aw);
        return gate_mul(__ldg(ah), __ldg(aw));
      };
      const bool has_mul = p.mul || (!PATCH && p.gate_h);
      // The epilogue is issue-bound (ncu on the lateral 1x1 + add: 51 % issue-active, 740 warp instructions per 32 x 32 piece, most of them
      // predicates and 64-bit address arithmetic of these operand rows), so everything that does not change inside a tile is computed once:
      // the lane's operand pointers at its first row (rows are rpp pixels = step_a / step_m elements apart), and how many of its rows are valid.
      const bf16* add_l = nullptr;
      const bf16* mul_l = nullptr;
      int nrows = 0;
      if constexpr (PREFETCH) {
        if (pre) {
          int64_t d0;
          int left;  // valid rows from the lane's first row on, in units of one row step
          if (PATCH) {  // rows rr0 + k * rpp of the 8-pixel-wide tile: the same column ox, image rows (rpp >> 3) apart
            const int oy0 = ty0 + 4 * q + (rr0 >> 3), ox = tx0 + (rr0 & 7);
            d0 = (int64_t)(img * p.hm + oy0) * p.wm + ox;
            left = ox < p.wm ? p.hm - oy0 : 0;
          } else {
            d0 = (int64_t)m_base + rr0;
            left = p.m_total - (m_base + rr0);
          }
          nrows = left <= 0 ? 0 : (left + (1 << rsh) - 1) >> rsh;
          nrows = nrows < cpr ? nrows : cpr;
          add_l = p.add + d0 * p.add_ld + n0 + cj * 8;
          mul_l = p.mul + d0 * p.mul_ld + n0 + cj * 8;
        }
      }
      auto prefetch = [&](int c0) {
        if (n0 + c0 + cj * 8 >= p.cout) return;
        if (pf_smem) {  // cp.async into this lane's own slots (row k at + 512 bytes, the mul block behind the add block): no cross-lane hazards
#pragma unroll
          for (int k = 0; k < 4; k++) {
            if (k >= nrows) continue;
            if (p.add) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pf_lane + 512u * k), "l"(add_l + c0 + k * step_a) : "memory");
            if (p.mul) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pf_lane + pf_mul + 512u * k), "l"(mul_l + c0 + k * step_m) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          return;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (k >= nrows) continue;
          if (p.add) asm volatile("prefetch.global.L2 [%0];" ::"l"(add_l + c0 + k * step_a));
          if (p.mul) asm volatile("prefetch.global.L2 [%0];" ::"l"(mul_l + c0 + k * step_m));
        }
      };
      if constexpr (PREFETCH) {
        if (pre && cb < ce) prefetch(cb);
      }
      const int acc = ar.idx;
      mbar_wait((tfull0 + 8u * (uint32_t)acc), ar.ph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(acc * p.ksplit * p.n_tile) + lane_base;
      const bool store_ok = !PATCH || (ty0 + 4 * q < p.hm);  // warp-uniform: a patch-mode box entirely below the image is not issued
      if (cb >= ce) {  // this warp owns no columns of so narrow a tile: hand the accumulator back at once
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive((tempty0 + 8u * (uint32_t)acc));
      }
      for (int c0 = cb; c0 < ce; c0 += p.sc, piece++) {
        const uint32_t stg = stg0 + (piece & (uint32_t)(p.stg_bufs - 1)) * p.stg_warp_bytes;
        const uint32_t row_addr = stg + (uint32_t)lane * RB;
        if (p.dbg & 2) {
        } else if (p.sc >= 32) {
          for (int u0 = 0; u0 < p.sc; u0 += 32)
            epi_unit<32, ACT, GN, SCALE>(p, tacc + (uint32_t)(c0 + u0), bias_s + 4u * (uint32_t)(n0 + c0 + u0), rsc, row_addr, ph, u0 >> 3, u0 == 0, lane,
                                         valid, img, uniform, n0 + c0 + u0);
        } else {
          epi_unit<16, ACT, GN, SCALE>(p, tacc + (uint32_t)c0, bias_s + 4u * (uint32_t)(n0 + c0), rsc, row_addr, ph, 0, true, lane, valid, img, uniform,
                                       n0 + c0);
        }
        if (c0 + p.sc >= ce) {  // last TMEM read of this tile by this warp: hand the accumulator back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive((tempty0 + 8u * (uint32_t)acc));
        }
        if constexpr (PREFETCH) {
          if (pre) {  // rows outside the map get zeros and are clipped by the bulk store
            __syncwarp();
            const int co = n0 + c0 + cj * 8;
            uint4 pav[4], pmv[4];
            if (pf_smem) asm volatile("cp.async.wait_group 0;" ::: "memory");  // this lane's own copies have landed
#pragma unroll
            for (int k = 0; k < 4; k++) {  // every row of the piece in flight at once (L2 hits after the hints above)
              pav[k] = make_uint4(0u, 0u, 0u, 0u);
              pmv[k] = make_uint4(0u, 0u, 0u, 0u);
              if (co >= p.cout || k >= nrows) continue;
              if (pf_smem) {
                if (p.add) pav[k] = lds16(pf_lane + 512u * k);
                if (p.mul) pmv[k] = lds16(pf_lane + pf_mul + 512u * k);
              } else {
                if (p.add) pav[k] = __ldg(reinterpret_cast<const uint4*>(add_l + c0 + k * step_a));
                if (p.mul) pmv[k] = __ldg(reinterpret_cast<const uint4*>(mul_l + c0 + k * step_m));
              }
            }
            if constexpr (!PATCH) {
              if (p.gate_h) {  // the gate rows of all four rows in flight together with the add rows (L1 / L2 residents); rows outside the map
                               // multiply by zero and are clipped by the bulk store.  The epilogue is issue-bound (ncu: 51 % issue-active with the
                               // add operand alone), so the position of the lane's first row is found once and the others step from it by rpp pixels
                uint4 pgw[4];
                int gi = g_img, gy = g_y, gx = g_x + rr0;
                while (gx >= p.wm) { gx -= p.wm; gy++; }
                while (gy >= p.hm) { gy -= p.hm; gi++; }
                const bf16* gh0 = p.gate_h + co;
                const bf16* gw0 = p.gate_w + co;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  pgw[k] = make_uint4(0u, 0u, 0u, 0u);
                  if (k < nrows && co < p.cout) {
                    pmv[k] = __ldg(reinterpret_cast<const uint4*>(gh0 + (gi * p.hm + gy) * p.gate_ld));
                    pgw[k] = __ldg(reinterpret_cast<const uint4*>(gw0 + (gi * p.wm + gx) * p.gate_ld));
                  }
                  gx += rpp;
                  while (gx >= p.wm) { gx -= p.wm; gy++; }
                  while (gy >= p.hm) { gy -= p.hm; gi++; }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) pmv[k] = gate_mul(pmv[k], pgw[k]);
              }
            }
            if (!pf_smem && c0 + p.sc < ce) prefetch(c0 + p.sc);  // the next piece's operands travel under this piece's store and the next TMEM reads
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if (k >= cpr) continue;
              const uint32_t ra = stg + (uint32_t)(rr0 + k * rpp) * RB;
              const uint32_t ca = ra + ((((uint32_t)cj) ^ ((ra >> 7) & swz)) << 4);
              uint4 u = lds16(ca);
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
              const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&pav[k]);
              const __nv_bfloat162* hm = reinterpret_cast<const __nv_bfloat162*>(&pmv[k]);
              if (has_mul && p.add) {
#pragma unroll
                for (int e = 0; e < 4; e++) h[e] = __hfma2(h[e], hm[e], ha[e]);
              } else if (has_mul) {
#pragma unroll
                for (int e = 0; e < 4; e++) h[e] = __hmul2(h[e], hm[e]);
              } else if (p.add) {
#pragma unroll
                for (int e = 0; e < 4; e++) h[e] = __hadd2(h[e], ha[e]);
              }
              sts16v(ca, u);
            }
            // staged operands: the slots are refilled only after their values have been consumed above
            if (pf_smem && c0 + p.sc < ce) prefetch(c0 + p.sc);
          }
        }
        if constexpr (MULADD) {
          if ((p.mul || p.add || p.gate_h) && !pre) {
            __syncwarp();
            const int co = n0 + c0 + cj * 8;
            if (co < p.cout) {
#pragma unroll 4
              for (int row = rr0; row < 32; row += rpp) {
                int64_t d;
                bool ok;
                if (PATCH) {
                  const int oy = ty0 + 4 * q + (row >> 3), ox = tx0 + (row & 7);
                  ok = oy < p.hm && ox < p.wm;
                  d = (int64_t)(img * p.hm + oy) * p.wm + ox;
                } else {
                  d = (int64_t)m_base + row;
                  ok = d < p.m_total;
                }
                if (!ok) continue;
                const uint32_t ra = stg + (uint32_t)row * RB;
                const uint32_t ca = ra + ((((uint32_t)cj) ^ ((ra >> 7) & swz)) << 4);
                uint4 u = lds16(ca);
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
                uint4 mv = make_uint4(0u, 0u, 0u, 0u);
                if (p.mul) mv = __ldg(reinterpret_cast<const uint4*>(p.mul + d * p.mul_ld + co));
                if constexpr (!PATCH) {
                  if (p.gate_h) mv = gate_vec(row, co);
                }
                const __nv_bfloat162* hm = reinterpret_cast<const __nv_bfloat162*>(&mv);
                if (has_mul && p.add) {
                  const uint4 av = __ldg(reinterpret_cast<const uint4*>(p.add + d * p.add_ld + co));
                  const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&av);
#pragma unroll
                  for (int e = 0; e < 4; e++) h[e] = __hfma2(h[e], hm[e], ha[e]);
                } else if (has_mul) {
#pragma unroll
                  for (int e = 0; e < 4; e++) h[e] = __hmul2(h[e], hm[e]);
                } else if (p.add) {
                  const uint4 av = __ldg(reinterpret_cast<const uint4*>(p.add + d * p.add_ld + co));
                  const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&av);
#pragma unroll
                  for (int e = 0; e < 4; e++) h[e] = __hadd2(h[e], ha[e]);
                }
                sts16v(ca, u);
              }
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (store_ok) {  // warp-uniform
          if (elect_one()) {
            if (PATCH) tma_store_4d(tmY, stg, n0 + c0, tx0, ty0 + 4 * q, img);
            else tma_store_2d(tmY, stg, n0 + c0, m_base);
            tma_store_commit();
          }
        }
      }
    }
    if (elect_one()) tma_store_wait_read();  // shared memory must outlive the last bulk store's read
  }
}

template <bool PATCH, int ACT, bool MULADD, bool GN, bool SCALE>
__global__ void __launch_bounds__(V2_MAX_THREADS, 1) conv2_kernel(const __grid_constant__ V2Params p, const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + p.off_bars;
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * p.stages + V2_MAX_ACC + a); };
  auto b_bar = [&](int i) { return bars + 8u * (uint32_t)(2 * p.stages + 2 * V2_MAX_ACC + i); };
  const uint32_t tmem_ptr_addr = bars + 8u * (uint32_t)(2 * p.stages + 2 * V2_MAX_ACC + V2_MAX_BCHUNKS);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbch = p.tiles_n * p.ntaps * p.kpt;
  const int per_img = p.tiles_x * p.tiles_y;

  if (tid == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < p.acc_stages; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), p.tsplit ? 4u : (uint32_t)p.ew); }
    for (int i = 0; i < nbch; i++) mbar_init(b_bar(i), 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  pdl_sync();  // on-chip prologue done; from here on the kernel reads what its predecessors wrote
  {            // bias table (zeros without a bias): tiles_n * n_tile floats
    float* bs = reinterpret_cast<float*>(smem_raw + (base + p.off_bias - raw));
    const int nb = p.tiles_n * p.n_tile;
    for (int i = tid; i < nb; i += (int)blockDim.x) bs[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ================= TMA producer (warp-uniform loop, the elected lane issues) =================
    {
      const bool leader = elect_one();
      // resident weights: chunk (nt, t, c) = 64 K columns [t * cin + 64 c, +64) of the n_tile rows of column tile nt
      for (int i = 0; i < nbch; i++) {
        const int c = i % p.kpt, t = (i / p.kpt) % p.ntaps, nt = i / (p.kpt * p.ntaps);
        if (leader) {
          mbar_expect_tx(b_bar(i), p.b_chunk_bytes);
          tma_load_2d(base + p.off_b + (uint32_t)i * p.b_chunk_bytes, &tmB, b_bar(i), (p.custom ? p.wtap[t] : t) * p.cin + c * 64, nt * p.n_tile);
        }
      }
      V2TileIter ti;
      ti.init(p, (int)blockIdx.x, (int)gridDim.x);
      V2Ring sr;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ti.template next<PATCH>(p)) {
        const int img = ti.img, ty0 = ti.ty * V2_BH, tx0 = ti.tx * V2_BW;
        for (int c = 0; c < p.kpt; c++, sr.next(p.stages)) {
          const int s = sr.idx;
          const uint32_t a_s = base + p.off_a + (uint32_t)s * p.a_stage_bytes;
          mbar_wait(empty_bar(s), sr.ph ^ 1u);
          if (leader) {
            mbar_expect_tx(full_bar(s), p.a_tx_bytes);
            if (PATCH) tma_load_4d(a_s, &tmA, full_bar(s), c * 64, tx0 - p.org, ty0 - p.org, img);
            else tma_load_2d(a_s, &tmA, full_bar(s), c * 64, img * V2_BM);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp walks the loops (warp-uniform control flow keeps descriptors and barrier addresses in uniform registers); the lane chosen by
    // elect.sync issues.  Profile history (profiles/r2_ncu_conv_v2.md): one divergent thread rebuilding 64-bit descriptors needed ~200 issue cycles
    // per tcgen05.mma and was THE limiter of the 3x3 convolutions; the warp-uniform loop still spent ~120 instructions per tap on index
    // arithmetic.  Now the tap / K-step loops are fully unrolled with compile-time offsets: two 32-bit adds per MMA.
    {
      const bool leader_lane = elect_one();
      const bool leader = leader_lane && !(p.dbg & 1);
      const uint32_t idesc = v2_idesc(p.n_tile);
      const uint32_t a_hi = v2_desc_hi(PATCH ? (uint32_t)p.pw * 128u : 1024u), b_hi = v2_desc_hi(1024u);
      const uint32_t b_lo0 = v2_desc_lo(base + p.off_b), b_step = p.b_chunk_bytes >> 4;
      const uint32_t pw8 = (uint32_t)p.pw * 8u;                   // one patch row in descriptor units (16 bytes)
      const uint32_t tstep = (uint32_t)p.kpt * b_step;            // weight chunks of consecutive taps
      constexpr int NT = PATCH ? 9 : 1;
      // every weight chunk must have landed before its first use: the first tile waits chunk by chunk (so its MMAs start under the weight
      // load), later tiles never look at those barriers again
      V2Ring sr, ar;
      int i = 0;
      int nt = (int)(blockIdx.x % (unsigned)p.tiles_n);
      const int nt_step = (int)(gridDim.x % (unsigned)p.tiles_n);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, i++, ar.next(p.acc_stages), nt = nt + nt_step >= p.tiles_n ? nt + nt_step - p.tiles_n : nt + nt_step) {
        const int acc = ar.idx;
        mbar_wait(tempty_bar(acc), ar.ph ^ 1u);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.ksplit * p.n_tile);
        const uint32_t nt_cols = (uint32_t)p.n_tile;
        const int ks = p.ksplit;  // tap t accumulates into split t % ks; its first visit (chunk 0, t < ks) overwrites
        const bool first_pass = i < p.tiles_n;  // the first visit of column tile nt (tiles of one CTA cycle through the column tiles)
        for (int c = 0; c < p.kpt; c++, sr.next(p.stages)) {
          const int s = sr.idx;
          mbar_wait(full_bar(s), sr.ph);
          tc_fence_after();
          const uint32_t a_lo0 = v2_desc_lo(base + p.off_a + (uint32_t)s * p.a_stage_bytes);
          const int bi0 = nt * (PATCH && p.custom ? p.ntaps : NT) * p.kpt + c;  // chunk of tap 0
          const uint32_t b_lo_c = b_lo0 + (uint32_t)bi0 * b_step;
          const int krem = p.cin - c * 64;
          const int ksteps = krem >= 64 ? 4 : (krem + 15) >> 4;
          const uint32_t acc0 = c > 0 ? 1u : 0u;                    // the first MMA of the tile overwrites the accumulator
          if (PATCH && p.custom) {  // run-time tap list (transposed-convolution phase): 1, 2 or 4 taps, one accumulator
            for (int t = 0; t < p.ntaps; t++) {
              if (first_pass) {
                mbar_wait(b_bar(bi0 + t * p.kpt), 0u);
                tc_fence_after();
              }
              const uint32_t a_lo = a_lo0 + (uint32_t)p.tap_row[t] * 8u, b_lo = b_lo_c + (uint32_t)t * tstep;
              if (leader) {
                umma_f16(d_tmem, pack64(a_lo, a_hi), pack64(b_lo, b_hi), idesc, t > 0 ? 1u : acc0);
                for (int k = 1; k < ksteps; k++) umma_bf16<true>(d_tmem, pack64(a_lo + 2u * k, a_hi), pack64(b_lo + 2u * k, b_hi), idesc);
              }
            }
          } else if (first_pass) {
            for (int t = 0; t < NT; t++) {
              mbar_wait(b_bar(bi0 + t * p.kpt), 0u);
              tc_fence_after();
              const uint32_t a_lo = a_lo0 + (PATCH ? (uint32_t)(t / 3) * pw8 + (uint32_t)(t % 3) * 8u : 0u), b_lo = b_lo_c + (uint32_t)t * tstep;
              if (leader) {
                const uint32_t d = d_tmem + (uint32_t)(t % ks) * nt_cols;
                umma_f16(d, pack64(a_lo, a_hi), pack64(b_lo, b_hi), idesc, t >= ks ? 1u : acc0);
                for (int k = 1; k < ksteps; k++) umma_bf16<true>(d, pack64(a_lo + 2u * k, a_hi), pack64(b_lo + 2u * k, b_hi), idesc);
              }
            }
          } else if (ksteps == 4) {
            if (leader) {
#pragma unroll
              for (int t = 0; t < NT; t++) {
                const uint32_t a_lo = a_lo0 + (PATCH ? (uint32_t)(t / 3) * pw8 + (uint32_t)(t % 3) * 8u : 0u), b_lo = b_lo_c + (uint32_t)t * tstep;
                const uint32_t d = d_tmem + (uint32_t)(t % ks) * nt_cols;
                if (t < 4) umma_f16(d, pack64(a_lo, a_hi), pack64(b_lo, b_hi), idesc, t >= ks ? 1u : acc0);
                else umma_bf16<true>(d, pack64(a_lo, a_hi), pack64(b_lo, b_hi), idesc);
                umma_bf16<true>(d, pack64(a_lo + 2u, a_hi), pack64(b_lo + 2u, b_hi), idesc);
                umma_bf16<true>(d, pack64(a_lo + 4u, a_hi), pack64(b_lo + 4u, b_hi), idesc);
                umma_bf16<true>(d, pack64(a_lo + 6u, a_hi), pack64(b_lo + 6u, b_hi), idesc);
              }
            }
          } else {
            if (leader) {
#pragma unroll
              for (int t = 0; t < NT; t++) {
                const uint32_t a_lo = a_lo0 + (PATCH ? (uint32_t)(t / 3) * pw8 + (uint32_t)(t % 3) * 8u : 0u), b_lo = b_lo_c + (uint32_t)t * tstep;
                const uint32_t d = d_tmem + (uint32_t)(t % ks) * nt_cols;
                if (t < 4) umma_f16(d, pack64(a_lo, a_hi), pack64(b_lo, b_hi), idesc, t >= ks ? 1u : acc0);
                else umma_bf16<true>(d, pack64(a_lo, a_hi), pack64(b_lo, b_hi), idesc);
                for (int k = 1; k < ksteps; k++) umma_bf16<true>(d, pack64(a_lo + 2u * k, a_hi), pack64(b_lo + 2u * k, b_hi), idesc);
              }
            }
          }
          if (leader_lane) umma_commit(empty_bar(s));
        }
        if (leader_lane) umma_commit(tfull_bar(acc));
      }
    }
    __syncwarp();
    tc_fence_before();
  } else {
    // ================= epilogue warps =================
    v2_epilogue<PATCH, ACT, MULADD, GN, SCALE>(p, &tmY, base, tmem_base, warp, lane, tfull_bar(0), tempty_bar(0));
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// =====================================================================================================================
// Small-channel 3x3 stride-1 convolution (cin, cout in {8, 16, 32}): the C3k2 bottlenecks of the first stages.
// conv2_kernel spends a 128-byte swizzled row per pixel on 16 - 64 bytes of data; conv_small_kernel (mma.sync) is instruction-bound: 33 M warp
// instructions for 1.0 M MMAs at 8 -> 16 @160^2 (profiles/r2_ncu_conv_v2.md).  Here the haloed patch (18 rows x 10 pixels) is staged by ONE TMA
// box per tile and the tensor core reads it in place:
//   cin = 8   no swizzle: in the K-major no-swizzle canonical layout a core matrix is 8 rows x 16 bytes stored contiguously, and the 8 pixels of an
//             output row ARE 8 contiguous 16-byte chunks of the patch.  An A descriptor starts anywhere in the patch (tap shift = pixel offset), its
//             stride-byte-offset (between 8-row groups) is one patch row = 160 bytes, and its leading-byte-offset (between the two 16-byte K chunks
//             of a K = 16 step) is whatever separates the two TAPS the step covers: +16 bytes for horizontally adjacent taps, +128 for the row
//             wrap -- five MMAs per 128 pixels, no im2col copy.  Dense inputs (ld == 8) merge the pixel and channel axes: a patch row is one
//             160-byte TMA line.
//   cin = 16 / 32   SWIZZLE_32B / SWIZZLE_64B rows of one pixel each (one / two K steps per tap), stride-byte-offset = one patch row.
//   stride 2 (cin = 16, dense input)   the input is viewed as (2 x 16 channels of a horizontal pixel PAIR, w / 2, row parity, h / 2, n): two TMA boxes
//             per tile, one per row parity, of 17 rows x 9 pairs x 64 bytes (SWIZZLE_64B).  Output pixels that are neighbours in x read neighbouring
//             PAIRS, so the 8 rows of a core matrix are again consecutive 64-byte rows; a tap selects the row-parity plane, the pair offset and the
//             32-byte half of the row (x parity) -- 9 MMAs per 128 output pixels from 306 TMA lines (conv_tma_kernel: one box per tap, 1152 lines of
//             32 bytes, ~3 cycles per line: 150 us for the 16 -> 32 layer at 320^2).
// The weights are laid out once per CTA in the no-swizzle canonical form ([step][K chunk][8-row group][8 rows][16 bytes]; every descriptor carries
// its own layout type).  Epilogue: v2_epilogue in tile-split mode.  Two CTAs per SM.
constexpr int C3_MAX_MMA = 18;   // 9 taps x 32 channels / 16
constexpr int C3_PLANE_BYTES = 18 * 10 * 16, C3_PLANE_STRIDE = 2944;  // 2880 rounded up to 128 (TMA destination alignment)
constexpr int C3_S2_PLANE = 10240;    // stride 2: one row-parity plane = 17 rows x 9 pixel pairs x 64 bytes = 9792, rounded up to 1024
struct C3Params {
  V2Params v;
  int nmma;
  int dense;                            // cin = 8 and x.ld == 8: one patch row (10 pixels) is ONE 160-byte TMA line of the merged (w * 8) axis
  int s2;                               // stride 2, cin = 16, dense input: two row-parity planes of [17 rows][9 pixel PAIRS][2 x 16 channels] (SWIZZLE_64B)
  uint32_t a_hi;                        // high descriptor word of the A operand (stride-byte-offset = one patch row, layout type)
  uint32_t tx_bytes;
  uint32_t stage_bytes, off_w;          // A stage; weights
  uint32_t a_off[C3_MAX_MMA], a_lbo[C3_MAX_MMA];
  uint32_t b_mma_bytes, b_lbo;
  const bf16* w;
  int w_row;                            // 9 * cin
};

template <int ACT, bool MULADD>
__global__ void __launch_bounds__(320, 2) conv3_kernel(const __grid_constant__ C3Params cp, const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmY) {
  extern __shared__ uint8_t smem_raw[];
  const V2Params& p = cp.v;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + p.off_bars;
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * p.stages + V2_MAX_ACC + a); };
  const uint32_t tmem_ptr_addr = bars + 8u * (uint32_t)(2 * p.stages + 2 * V2_MAX_ACC);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < p.acc_stages; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4u); }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  pdl_sync();
  {  // bias table and the weights in canonical no-swizzle K-major form: chunk jc = K columns [8 jc, 8 jc + 8) of the [cout][9 cin] matrix
    float* bs = reinterpret_cast<float*>(smem_raw + (base + p.off_bias - raw));
    for (int i = tid; i < p.n_tile; i += (int)blockDim.x) bs[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
    uint8_t* wsm = smem_raw + (base + cp.off_w - raw);
    const int nchunk = cp.w_row >> 3;
    for (int i = tid; i < cp.nmma * 2 * p.n_tile; i += (int)blockDim.x) {
      const int n = i % p.n_tile, jc = i / p.n_tile;  // jc = 2 * step + chunk-in-step
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (n < p.cout && jc < nchunk) v = *reinterpret_cast<const uint4*>(cp.w + (int64_t)n * cp.w_row + jc * 8);
      *reinterpret_cast<uint4*>(wsm + (size_t)(jc >> 1) * cp.b_mma_bytes + (size_t)(jc & 1) * cp.b_lbo + (size_t)(n >> 3) * 128 + (size_t)(n & 7) * 16) = v;
    }
  }
  fence_proxy_async();  // generic-proxy writes of the weights -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    const bool leader = elect_one();
    V2TileIter ti;
    ti.init(p, (int)blockIdx.x, (int)gridDim.x);
    V2Ring sr;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ti.next(p), sr.next(p.stages)) {
      const int s = sr.idx;
      const uint32_t a_s = base + p.off_a + (uint32_t)s * cp.stage_bytes;
      mbar_wait(empty_bar(s), sr.ph ^ 1u);
      if (leader) {
        mbar_expect_tx(full_bar(s), cp.tx_bytes);
        if (cp.s2) {
          tma_load_5d(a_s, &tmA, full_bar(s), 0, ti.tx * V2_BW - 1, 0, ti.ty * V2_BH - 1, ti.img);
          tma_load_5d(a_s + C3_S2_PLANE, &tmA, full_bar(s), 0, ti.tx * V2_BW - 1, 1, ti.ty * V2_BH - 1, ti.img);
        } else if (cp.dense) {
          tma_load_3d(a_s, &tmA, full_bar(s), (ti.tx * V2_BW - 1) * 8, ti.ty * V2_BH - 1, ti.img);
        } else {
          tma_load_4d(a_s, &tmA, full_bar(s), 0, ti.tx * V2_BW - 1, ti.ty * V2_BH - 1, ti.img);
        }
      }
    }
  } else if (warp == 1) {
    const bool leader_lane = elect_one();
    const bool leader = leader_lane && !(p.dbg & 1);
    const uint32_t idesc = v2_idesc(p.n_tile);
    const uint32_t a_hi = cp.a_hi, b_hi = (128u >> 4) | (1u << 14);  // weights: no swizzle, SBO = one 8-row group
    const uint32_t b_lo0 = ((base + cp.off_w) & 0x3FFFFu) >> 4, b_lbo = (cp.b_lbo >> 4) << 16, b_step = cp.b_mma_bytes >> 4;
    V2Ring sr, ar;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, sr.next(p.stages), ar.next(p.acc_stages)) {
      const int acc = ar.idx, s = sr.idx;
      mbar_wait(tempty_bar(acc), ar.ph ^ 1u);
      mbar_wait(full_bar(s), sr.ph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_tile);
      const uint32_t a_s = base + p.off_a + (uint32_t)s * cp.stage_bytes;
      if (leader) {
        for (int i = 0; i < cp.nmma; i++) {
          const uint32_t a_lo = (((a_s + cp.a_off[i]) & 0x3FFFFu) >> 4) | ((cp.a_lbo[i] >> 4) << 16);
          const uint32_t b_lo = (b_lo0 + (uint32_t)i * b_step) | b_lbo;
          umma_f16(d_tmem, pack64(a_lo, a_hi), pack64(b_lo, b_hi), idesc, i > 0 ? 1u : 0u);
        }
      }
      if (leader_lane) {
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(acc));
      }
    }
    __syncwarp();
    tc_fence_before();
  } else {
    v2_epilogue<true, ACT, MULADD, false, false, true>(p, &tmY, base, tmem_base, warp, lane, tfull_bar(0), tempty_bar(0));
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// =====================================================================================================================
// Modulated deformable 3x3 convolution (DCNv2; mmcv ModulatedDeformConv2d of DyDCNv2, nn/modules/head.py:751-782), 64 input channels.
// conv_tc_kernel<DEFORM> gathered the four bilinear corners of every (pixel, tap) from global memory (4 x 9 x 128 B per pixel): 370 us at
// 80x80, batch 64, bound by L1 / L2 gather traffic.  Here the M tile is the 8 x 16 pixel rectangle of the 3x3 path and
//   warp 0        : TMA -- the tile's input patch with a halo of R pixels (one 4-D box, zero fill = the operator's zero padding) and the tile's
//                   offset / mask channels (one box) per tile, double-buffered;
//   warps 10-17   : sampling parameters of the 128 x 9 (pixel, tap) pairs (floor, bilinear weights x sigmoid(mask), position inside the patch)
//                   once per tile, then per tap the blended A tile: 4 conflict-free 128-bit SHARED-memory reads per 8 channels, packed bf16x2
//                   blend (same arithmetic as before), swizzled store into a 2-slot A ring (a corner outside the staged halo falls back to
//                   global loads);
//   warp 1        : 4 tcgen05.mma per tap against the resident weights, accumulator double-buffered in TMEM;
//   warps 2-9     : the conv_v2 epilogue (GroupNorm statistics, bf16, bulk tensor store).
// Shared-memory traffic (gather reads 590 KB + A writes / MMA reads 300 KB per tile) is the bound: ~90 us at 80x80.
// =====================================================================================================================
constexpr int DCN_GW = 8;            // gather warps
constexpr int DCN_TAPS = 9;
struct DcnParams {
  V2Params v;
  int R, pwr, phr;                   // halo radius, raw patch width / height in pixels
  int aslots;
  uint32_t off_raw, raw_stage_bytes, raw_tx_bytes, om_off, off_aslot, off_par;
  const bf16* x;
  int x_ld;
};
struct DcnSample {                   // 16 bytes per (tap, pixel)
  int pos;                           // row index of the top-left corner inside the raw patch, or -1: outside the staged halo
  int yx;                            // (y0 << 16) | (x0 & 0xffff) in image coordinates (fallback path)
  uint32_t w01, w23;                 // bilinear weights x mask as bf16 pairs: (w00, w01), (w10, w11)
};

template <int ACT, bool GN>
__global__ void __launch_bounds__(V2_MAX_THREADS, 1) dcn2_kernel(const __grid_constant__ DcnParams dp, const __grid_constant__ CUtensorMap tmX,
                                                                 const __grid_constant__ CUtensorMap tmOM, const __grid_constant__ CUtensorMap tmB,
                                                                 const __grid_constant__ CUtensorMap tmY) {
  extern __shared__ uint8_t smem_raw[];
  const V2Params& p = dp.v;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + p.off_bars;
  // barrier slots: rawfull[2] rawempty[2] afull[4] aempty[4] tfull[2] tempty[2] b[9]
  auto rawfull = [&](int i) { return bars + 8u * (uint32_t)i; };
  auto rawempty = [&](int i) { return bars + 8u * (uint32_t)(2 + i); };
  auto afull = [&](int i) { return bars + 8u * (uint32_t)(4 + i); };
  auto aempty = [&](int i) { return bars + 8u * (uint32_t)(8 + i); };
  auto tfull_bar = [&](int i) { return bars + 8u * (uint32_t)(12 + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (uint32_t)(14 + i); };
  auto b_bar = [&](int i) { return bars + 8u * (uint32_t)(16 + i); };
  const uint32_t tmem_ptr_addr = bars + 8u * 25u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int per_img = p.tiles_x * p.tiles_y;

  if (tid == 0) {
    for (int i = 0; i < 2; i++) { mbar_init(rawfull(i), 1); mbar_init(rawempty(i), DCN_GW); mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), (uint32_t)p.ew); }
    for (int i = 0; i < 4; i++) { mbar_init(afull(i), DCN_GW); mbar_init(aempty(i), 1); }
    for (int i = 0; i < DCN_TAPS; i++) mbar_init(b_bar(i), 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOM) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  pdl_sync();
  {
    float* bs = reinterpret_cast<float*>(smem_raw + (base + p.off_bias - raw));
    for (int i = tid; i < p.n_tile; i += (int)blockDim.x) bs[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ================= TMA producer =================
    const bool leader = elect_one();
    for (int t = 0; t < DCN_TAPS; t++) {
      if (leader) {
        mbar_expect_tx(b_bar(t), p.b_chunk_bytes);
        tma_load_2d(base + p.off_b + (uint32_t)t * p.b_chunk_bytes, &tmB, b_bar(t), t * 64, 0);
      }
    }
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
      const int img = tile / per_img, r = tile - img * per_img;
      const int ty0 = (r / p.tiles_x) * V2_BH, tx0 = (r % p.tiles_x) * V2_BW;
      const int rs = it & 1;
      const uint32_t dst = base + dp.off_raw + (uint32_t)rs * dp.raw_stage_bytes;
      mbar_wait(rawempty(rs), ((it >> 1) & 1u) ^ 1u);
      if (leader) {
        mbar_expect_tx(rawfull(rs), dp.raw_tx_bytes);
        tma_load_4d(dst, &tmX, rawfull(rs), 0, tx0 - dp.R, ty0 - dp.R, img);
        tma_load_4d(dst + dp.om_off, &tmOM, rawfull(rs), 0, tx0, ty0, img);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const bool leader = elect_one();
    const uint32_t idesc = v2_idesc(p.n_tile);
    const uint32_t hi = v2_desc_hi(1024u);
    const uint32_t b_lo0 = v2_desc_lo(base + p.off_b), b_step = p.b_chunk_bytes >> 4;
    uint32_t git = 0;
    int i = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, i++) {
      const int acc = i & 1;
      mbar_wait(tempty_bar(acc), (((uint32_t)(i >> 1)) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_tile);
      for (int t = 0; t < DCN_TAPS; t++, git++) {
        const int s = (int)(git % (uint32_t)dp.aslots);
        if (i == 0) { mbar_wait(b_bar(t), 0u); }
        mbar_wait(afull(s), (git / (uint32_t)dp.aslots) & 1u);
        tc_fence_after();
        const uint32_t a_lo = v2_desc_lo(base + dp.off_aslot + (uint32_t)s * 16384u), b_lo = b_lo0 + (uint32_t)t * b_step;
        if (leader) {
          umma_f16(d_tmem, pack64(a_lo, hi), pack64(b_lo, hi), idesc, t > 0 ? 1u : 0u);
          umma_bf16<true>(d_tmem, pack64(a_lo + 2u, hi), pack64(b_lo + 2u, hi), idesc);
          umma_bf16<true>(d_tmem, pack64(a_lo + 4u, hi), pack64(b_lo + 4u, hi), idesc);
          umma_bf16<true>(d_tmem, pack64(a_lo + 6u, hi), pack64(b_lo + 6u, hi), idesc);
          umma_commit(aempty(s));
        }
      }
      if (leader) umma_commit(tfull_bar(acc));
    }
    __syncwarp();
    tc_fence_before();
  } else if (warp < 2 + p.ew) {
    v2_epilogue<true, ACT, false, GN, false>(p, &tmY, base, tmem_base, warp, lane, tfull_bar(0), tempty_bar(0));
  } else {
    // ================= gather warps =================
    const int gtid = tid - 32 * (2 + p.ew);                // 0 .. 255
    const int gc = gtid & 7, grow = gtid >> 3;             // 16-byte channel chunk; pixel row inside a pass of 32 pixels
    DcnSample* par = reinterpret_cast<DcnSample*>(smem_raw + (base + dp.off_par - raw));
    const int hi_ = p.hm, wi_ = p.wm;
    uint32_t it = 0, git = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
      const int img = tile / per_img, r = tile - img * per_img;
      const int ty0 = (r / p.tiles_x) * V2_BH, tx0 = (r % p.tiles_x) * V2_BW;
      const int rs = it & 1;
      const uint32_t xs = base + dp.off_raw + (uint32_t)rs * dp.raw_stage_bytes;
      const bf16* oms = reinterpret_cast<const bf16*>(smem_raw + (xs + dp.om_off - raw));
      mbar_wait(rawfull(rs), (it >> 1) & 1u);
      asm volatile("bar.sync 2, 256;" ::: "memory");       // every gather warp is done with the previous tile's parameters
      // ---- sampling parameters: [tap][pixel]
      for (int idx = gtid; idx < DCN_TAPS * V2_BM; idx += 32 * DCN_GW) {
        const int t = idx >> 7, pix = idx & 127;
        const int py = ty0 + (pix >> 3), px = tx0 + (pix & 7);
        DcnSample sm;
        sm.pos = 0; sm.yx = 0; sm.w01 = 0u; sm.w23 = 0u;
        if (py < hi_ && px < wi_) {
          const bf16* o = oms + pix * 32;
          const float ody = __bfloat162float(o[2 * t]), odx = __bfloat162float(o[2 * t + 1]);
          const float mk = sigmoidf_(__bfloat162float(o[18 + t]));
          const float fy_ = (float)(py + t / 3 - 1) + ody, fx_ = (float)(px + t % 3 - 1) + odx;
          if (fy_ > -1.f && fx_ > -1.f && fy_ < (float)hi_ && fx_ < (float)wi_) {
            const float fy = floorf(fy_), fx = floorf(fx_);
            const int y0 = (int)fy, x0 = (int)fx;
            const float ly = fy_ - fy, lx = fx_ - fx;
            const bool vy0 = y0 >= 0, vy1 = y0 + 1 < hi_, vx0 = x0 >= 0, vx1 = x0 + 1 < wi_;
            const float w00 = (vy0 && vx0) ? (1.f - ly) * (1.f - lx) * mk : 0.f, w01 = (vy0 && vx1) ? (1.f - ly) * lx * mk : 0.f;
            const float w10 = (vy1 && vx0) ? ly * (1.f - lx) * mk : 0.f, w11 = (vy1 && vx1) ? ly * lx * mk : 0.f;
            __nv_bfloat162 a = __floats2bfloat162_rn(w00, w01), b = __floats2bfloat162_rn(w10, w11);
            sm.w01 = *reinterpret_cast<uint32_t*>(&a);
            sm.w23 = *reinterpret_cast<uint32_t*>(&b);
            sm.yx = (y0 << 16) | (x0 & 0xFFFF);
            const int ry = y0 - (ty0 - dp.R), rx = x0 - (tx0 - dp.R);
            sm.pos = (ry >= 0 && ry + 1 < dp.phr && rx >= 0 && rx + 1 < dp.pwr) ? ry * dp.pwr + rx : -1;
          }
        }
        par[idx] = sm;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      // ---- per tap: blended A tile
      for (int t = 0; t < DCN_TAPS; t++, git++) {
        const int s = (int)(git % (uint32_t)dp.aslots);
        const uint32_t a_s = base + dp.off_aslot + (uint32_t)s * 16384u;
        mbar_wait(aempty(s), ((git / (uint32_t)dp.aslots) & 1u) ^ 1u);
#pragma unroll
        for (int pass = 0; pass < 4; pass++) {
          const int pix = pass * 32 + grow;
          const DcnSample sm = par[t * V2_BM + pix];
          uint4 u00, u01, u10, u11;
          if (sm.pos >= 0) {
            const uint32_t r00 = (uint32_t)sm.pos, r10 = r00 + (uint32_t)dp.pwr;
            u00 = lds16(xs + r00 * 128u + ((((uint32_t)gc) ^ (r00 & 7u)) << 4));
            u01 = lds16(xs + (r00 + 1u) * 128u + ((((uint32_t)gc) ^ ((r00 + 1u) & 7u)) << 4));
            u10 = lds16(xs + r10 * 128u + ((((uint32_t)gc) ^ (r10 & 7u)) << 4));
            u11 = lds16(xs + (r10 + 1u) * 128u + ((((uint32_t)gc) ^ ((r10 + 1u) & 7u)) << 4));
          } else {  // a corner outside the staged halo (large learned offsets): global loads, corners clamped into the map (invalid ones have weight 0)
            const int y0 = sm.yx >> 16, x0 = (int)(short)(sm.yx & 0xFFFF);
            const int ya = max(y0, 0), yb = min(y0 + 1, hi_ - 1), xa = max(x0, 0), xb = min(x0 + 1, wi_ - 1);
            const bf16* xb_ = dp.x + (int64_t)img * hi_ * wi_ * dp.x_ld + gc * 8;
            u00 = __ldg(reinterpret_cast<const uint4*>(xb_ + (int64_t)(ya * wi_ + xa) * dp.x_ld));
            u01 = __ldg(reinterpret_cast<const uint4*>(xb_ + (int64_t)(ya * wi_ + xb) * dp.x_ld));
            u10 = __ldg(reinterpret_cast<const uint4*>(xb_ + (int64_t)(yb * wi_ + xa) * dp.x_ld));
            u11 = __ldg(reinterpret_cast<const uint4*>(xb_ + (int64_t)(yb * wi_ + xb) * dp.x_ld));
          }
          const __nv_bfloat162 wa = *reinterpret_cast<const __nv_bfloat162*>(&sm.w01), wb = *reinterpret_cast<const __nv_bfloat162*>(&sm.w23);
          const __nv_bfloat162 w00 = __low2bfloat162(wa), w01 = __high2bfloat162(wa), w10 = __low2bfloat162(wb), w11 = __high2bfloat162(wb);
          const __nv_bfloat162* h00 = reinterpret_cast<const __nv_bfloat162*>(&u00);
          const __nv_bfloat162* h01 = reinterpret_cast<const __nv_bfloat162*>(&u01);
          const __nv_bfloat162* h10 = reinterpret_cast<const __nv_bfloat162*>(&u10);
          const __nv_bfloat162* h11 = reinterpret_cast<const __nv_bfloat162*>(&u11);
          uint4 res;
          __nv_bfloat162* hr = reinterpret_cast<__nv_bfloat162*>(&res);
#pragma unroll
          for (int e = 0; e < 4; e++) hr[e] = __hfma2(w11, h11[e], __hfma2(w10, h10[e], __hfma2(w01, h01[e], __hmul2(w00, h00[e]))));
          sts16v(a_s + (uint32_t)pix * 128u + ((((uint32_t)gc) ^ ((uint32_t)pix & 7u)) << 4), res);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(afull(s));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(rawempty(rs));
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

int v2_make_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { yad_set_error("conv2d_v2: cuTensorMapEncodeTiled is unavailable"); return 1; }
  cuuint64_t d[5], st[5];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) st[i] = strides_bytes[i];
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { yad_set_error("conv2d_v2: cuTensorMapEncodeTiled failed with %d", (int)r); return 1; }
  return 0;
}

int v2_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && e[0] ? atoi(e) : dflt;
}

template <bool PATCH, int ACT, bool MULADD, bool GN, bool SCALE>
int v2_launch_t(const V2Params& p, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, int grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv2_kernel<PATCH, ACT, MULADD, GN, SCALE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      yad_set_error("conv2d_v2: cannot raise the dynamic shared memory limit");
      return 2;
    }
    attr_set = true;
  }
  YAD_LAUNCH((conv2_kernel<PATCH, ACT, MULADD, GN, SCALE>), grid, 64 + 32 * p.ew, smem, st, p, tmA, tmB, tmY);
  YAD_LAUNCH_CHECK("conv2d_v2");
  return 0;
}

template <bool PATCH>
int v2_dispatch(const V2Params& p, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, int grid, size_t smem, cudaStream_t st) {
  const bool muladd = p.mul || p.add || p.gate_h, gn = p.gn_stats != nullptr, scale = p.img_scale || p.pix_scale;
#define V2_CASE(ACT_, MA_, GN_, SC_) return v2_launch_t<PATCH, ACT_, MA_, GN_, SC_>(p, tmA, tmB, tmY, grid, smem, st)
  if (!gn && !scale) {
    if (p.act == YAD_ACT_SILU) { if (muladd) V2_CASE(YAD_ACT_SILU, true, false, false); else V2_CASE(YAD_ACT_SILU, false, false, false); }
    if (p.act == YAD_ACT_NONE) { if (muladd) V2_CASE(YAD_ACT_NONE, true, false, false); else V2_CASE(YAD_ACT_NONE, false, false, false); }
    if constexpr (!PATCH) {
      if (p.act == YAD_ACT_SIGMOID) { if (muladd) V2_CASE(YAD_ACT_SIGMOID, true, false, false); else V2_CASE(YAD_ACT_SIGMOID, false, false, false); }
      if (p.act == YAD_ACT_RELU && !muladd) V2_CASE(YAD_ACT_RELU, false, false, false);  // cls_prob_conv.0 of the head
    }
  }
  if constexpr (!PATCH) {
    // per-row scales without GroupNorm statistics / mul / add (the head's cv3 with cls_prob folded into the rows): before this variant existed
    // the call fell through to the generic all-features kernel (57.6 us at 64 -> 80 @80^2 against 26.9 us for the plain 64 -> 64)
    if (!gn && scale && !muladd && p.act == YAD_ACT_NONE) V2_CASE(YAD_ACT_NONE, false, false, true);
  }
  if (gn && !muladd && p.act == YAD_ACT_NONE) {
    if (!scale) V2_CASE(YAD_ACT_NONE, false, true, false);
    if constexpr (!PATCH) V2_CASE(YAD_ACT_NONE, false, true, true);
  }
  if constexpr (!PATCH) V2_CASE(ACT_GENERIC, true, true, true);
  else if (!scale) V2_CASE(ACT_GENERIC, true, true, false);
#undef V2_CASE
  yad_set_error("conv2d_v2: no kernel variant for this epilogue");
  return 1;
}

int pick_n_tile_v2(int cout) {
  const int c16 = (cout + 15) / 16 * 16;
  if (c16 <= 256) return c16;
  const int tiles = (c16 + 255) / 256;
  return ((c16 + tiles - 1) / tiles + 15) / 16 * 16;
}

}  // namespace

// One output phase (oy % 2, ox % 2) = (py, px) of the k3 s2 p1 op1 transposed convolution (nn.ConvTranspose2d of the neck, z-yaml layers 13 / 20):
// a stride-1 convolution over the INPUT grid with 1, 2 or 4 taps at offsets 0 / +1, written to every second pixel of every second output row.
struct V2Phase {
  int py, px, ntaps;
  int dy[4], dx[4], wtap[4];  // input offset of tap t and its index in the [cout][9][cin] weights
};

// 0: not eligible (the caller falls back to conv_tma_kernel / conv_tc_kernel); fills the launch plan otherwise
static int v2_plan(const yad_tensor* x, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, V2Params* out, size_t* smem_out,
                   const V2Phase* phase = nullptr) {
  if (phase) {
    if (d->mode != YAD_CONV_TRANSPOSED || y->h != 2 * x->h || y->w != 2 * x->w || y->n != x->n) return 0;
    if (e->mul || e->add || e->gate_h || e->gn_stats || e->img_scale || e->pix_scale) return 0;
  } else if (d->mode != YAD_CONV_NORMAL || d->stride != 1) {
    return 0;
  }
  const bool flat = !phase && d->kh == 1 && d->kw == 1 && d->pad_h == 0 && d->pad_w == 0;
  const bool patch = phase || (d->kh == 3 && d->kw == 3 && d->pad_h == 1 && d->pad_w == 1);
  if (!flat && !patch) return 0;
  if (x->c < 16 || (x->c % 8) || (y->c % 8) || (x->ld % 8) || (y->ld % 8)) return 0;
  if (((uintptr_t)x->ptr & 15) || ((uintptr_t)y->ptr & 15)) return 0;
  if (!phase && (y->h != x->h || y->w != x->w || y->n != x->n)) return 0;
  if (e->gn_stats && e->gn_groups == 0) return 0;  // fused batch statistics (train-mode BatchNorm) stay on conv_tma_kernel<true>
  if (patch && (e->img_scale || e->pix_scale)) return 0;
  if ((e->mul && ((uintptr_t)e->mul & 15 || e->mul_ld % 8)) || (e->add && ((uintptr_t)e->add & 15 || e->add_ld % 8))) return 0;
  if (e->gate_h && (!flat || e->mul || e->gn_stats || e->img_scale || e->pix_scale)) return 0;
  if (e->gate_h && (int64_t)x->n * (y->h > y->w ? y->h : y->w) * e->gate_ld >= (int64_t)1 << 31) return 0;  // 32-bit gate row offsets
  const int64_t M = (int64_t)x->n * x->h * x->w;
  if (M + V2_BM >= (int64_t)1 << 31) return 0;
  V2Params p;
  memset(&p, 0, sizeof(p));
  p.n = x->n; p.hm = x->h; p.wm = x->w; p.hw = x->h * x->w; p.cin = x->c; p.cout = y->c; p.m_total = (int)M;  // output grid = input grid (per phase)
  p.n_tile = pick_n_tile_v2(p.cout);
  {
    // 80 / 112 / ... output channels: the columns are dealt to the epilogue warps in units of 16, and one 16-column share forces 16-column (32-byte)
    // store boxes on every warp.  Padding the tile to the next multiple of 32 keeps 32-column boxes (the extra accumulator columns come from
    // zero-filled weight rows and are clipped by the store): by itself neutral (64 -> 80 @80^2: 33.2 us either way); together with the tile split 37 -> 29 us.  YAD_CONV2_NPAD=0 keeps the exact tile.
    static int npad_env = -1;
    if (npad_env < 0) npad_env = v2_env("YAD_CONV2_NPAD", 1);
    if (npad_env && !e->gn_stats && p.n_tile > 64 && (p.n_tile % 32) && p.n_tile + 16 <= 256 && p.cout <= p.n_tile) p.n_tile += 16;
  }
  p.tiles_n = (p.cout + p.n_tile - 1) / p.n_tile;
  p.ntaps = phase ? phase->ntaps : (patch ? 9 : 1);
  p.kpt = (p.cin + 63) / 64;
  if (p.tiles_n * p.ntaps * p.kpt > V2_MAX_BCHUNKS) return 0;
  if (e->gn_stats) {
    if (e->gn_groups <= 0 || y->c % e->gn_groups) return 0;
    p.cpg = y->c / e->gn_groups;
    if (p.cpg != 4 && p.cpg != 8 && p.cpg != 16) return 0;
  }
  {
    static int ks_env = -1;
    if (ks_env < 0) ks_env = v2_env("YAD_CONV2_KSPLIT", 1);
    p.ksplit = 1;
    if (patch && (ks_env == 2 || ks_env == 4) && 2 * ks_env * p.n_tile <= 512) p.ksplit = ks_env;
  }
  {
    static int dbg_env = -1;
    if (dbg_env < 0) dbg_env = (v2_env("YAD_CONV2_DBG", 0) & 7) | (v2_env("YAD_CONV2_PREFETCH", 1) ? 0 : 8);
    p.dbg = dbg_env;
  }
  static int ew_env = -1, ts_env = -1;
  if (ew_env < 0) ew_env = v2_env("YAD_CONV2_EW", 0);
  if (ts_env < 0) ts_env = v2_env("YAD_CONV2_TSPLIT", 1);
  // measured (profiles/r2_conv_tsplit.txt): 1x1 convolutions with <= 64 output channels gain 10-17 %; the 3x3 patch kernels are not bound by the
  // epilogue's latency and keep the column split (YAD_CONV2_TSPLIT=2 forces the tile split there too)
  // 1x1 tiles of 65 .. 128 columns (4 accumulator stages for the 4 warp groups): measured at batch 64 @80^2 after the warp groups stopped walking
  // the tiles of the other groups: 128 -> 128 plain 43.9 (column split) vs 45.6 us (tile split), + add 77.4 vs 68.4, + gate + add 97.0 vs 86.5,
  // 64 -> 80 (96-column tile) 37.3 vs 28.9 -- tile split when the epilogue carries operand rows or the column shares are uneven
  // (YAD_CONV2_TSPLIT=3: always, 0: never)
  const bool wide_ts = flat && p.n_tile > 64 && p.n_tile <= 128 && (ts_env == 3 || e->mul || e->add || e->gate_h || (p.n_tile % 64));
  const bool tsplit = ts_env && p.ksplit == 1 && ((p.n_tile <= 64 && (flat || ts_env == 2)) || wide_ts);
  p.acc_stages = (2 * p.ksplit * p.n_tile <= 512) ? 2 : 1;
  if (tsplit) p.acc_stages = 512 / p.n_tile < V2_MAX_ACC ? 512 / p.n_tile : V2_MAX_ACC;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.acc_stages * p.ksplit * p.n_tile) p.tmem_cols <<= 1;
  auto set_ew = [&](int ew) {
    p.ew = ew;
    if (tsplit) {  // whole tiles per 4-warp group; store boxes of at most 32 columns keep the staging tiles small
      p.tsplit = ew / 4;
      for (int w = 0; w <= 4; w++) p.bnd[w] = w == 0 ? 0 : p.n_tile;
      p.sc = (p.n_tile % 32) ? 16 : 32;
      return;
    }
    const int ways = p.ew / 4, u16 = p.n_tile / 16;  // columns are dealt in units of 16, the first ways take the remainder
    int at = 0;
    for (int w = 0; w < ways; w++) { p.bnd[w] = at; at += 16 * (u16 / ways + (w < u16 % ways ? 1 : 0)); }
    for (int w = ways; w <= 4; w++) p.bnd[w] = p.n_tile;
    int sc = 64;  // store box width: the widest of 64 / 32 / 16 columns that divides every warp's range
    for (int w = 0; w < ways; w++)
      while (sc > 16 && ((p.bnd[w + 1] - p.bnd[w]) % sc)) sc >>= 1;
    p.sc = sc;
  };
  set_ew(ew_env == 8 || ew_env == 16 ? ew_env : (p.n_tile > 64 || tsplit ? 16 : 8));
  if (patch) {
    static int pw_env = -1;
    if (pw_env < 0) pw_env = v2_env("YAD_CONV2_PW", 10);
    p.pw = pw_env == 16 ? 16 : 10;
    p.tiles_x = (p.wm + V2_BW - 1) / V2_BW;
    p.tiles_y = (p.hm + V2_BH - 1) / V2_BH;
    for (int t = 0; t < 9; t++) p.tap_row[t] = (t / 3) * p.pw + (t % 3);
    p.org = 1;
    if (phase) {
      p.custom = 1;
      p.org = 0;
      for (int t = 0; t < phase->ntaps; t++) { p.tap_row[t] = phase->dy[t] * p.pw + phase->dx[t]; p.wtap[t] = phase->wtap[t]; }
    }
    p.a_tx_bytes = (uint32_t)(p.pw * (V2_BH + 2) * 128);
    p.a_stage_bytes = (p.a_tx_bytes + 1023u) & ~1023u;
    p.total_tiles = p.n * p.tiles_x * p.tiles_y * p.tiles_n;
    // few tiles per CTA (20 x 20 maps at batch 64: 2.6): loading the whole weight matrix up front does not pay and the fixed 16 x 8 tile wastes half
    // of its rows; conv_tma_kernel with its fitted patch measured 8.1 us against 12.1 us here (profiles/r2_launch_floor_20.jsonl)
    if (p.total_tiles < 5 * 148 && d->impl != 4) return 0;  // (transposed phases alike: 20 -> 40 at batch 64 measured 51 us here, 43.5 us on conv_tma_kernel)
  } else {
    p.tiles_x = p.tiles_y = 1;
    p.a_tx_bytes = V2_BM * 128;
    p.a_stage_bytes = p.a_tx_bytes;
    p.total_tiles = (int)((M + V2_BM - 1) / V2_BM) * p.tiles_n;
  }
  p.b_chunk_bytes = (uint32_t)p.n_tile * 128u;
  const uint32_t b_total = (uint32_t)(p.tiles_n * p.ntaps * p.kpt) * p.b_chunk_bytes;
  const uint32_t bias_bytes = ((uint32_t)(p.tiles_n * p.n_tile) * 4u + 127u) & ~127u;
  const uint32_t budget = 227u * 1024u - 1024u;
  const uint32_t max_stages = patch ? 4u : 8u, min_stages = patch ? 2u : 3u;
  uint32_t stages = 0, stg_total = 0;
  auto total = [&](uint32_t s) { return b_total + s * p.a_stage_bytes + stg_total + bias_bytes + 8u * (2u * s + 2u * V2_MAX_ACC + V2_MAX_BCHUNKS) + 16u; };
  // preference order: (epilogue warps as chosen, 2 staging tiles) > (same, 1 tile) > (8 warps, 2 tiles) > (8 warps, 1 tile), each only if the
  // activation ring keeps min_stages
  for (int attempt = 0; attempt < 4; attempt++) {
    p.stg_bufs = (attempt & 1) ? 1 : 2;
    if (attempt == 2) {
      if (p.ew == 8) break;
      set_ew(8);
    }
    p.stg_warp_bytes = 32u * (uint32_t)p.sc * 2u;  // 1024 / 2048 / 4096: every staging tile starts on its swizzle period
    stg_total = (uint32_t)(p.ew * p.stg_bufs) * p.stg_warp_bytes;
    stages = max_stages;
    while (stages >= 2 && total(stages) > budget) stages--;
    if (stages >= min_stages) break;
  }
  if (stages < 2) return 0;
  // operand rows (add / mul) staged through shared memory by cp.async, issued ahead of the accumulator reads: taken out of the activation ring when
  // that keeps at least 4 (1x1) / 2 (3x3) stages (YAD_CONV2_PFSMEM=0: L2 prefetch hints + direct loads only)
  uint32_t pf_total = 0;
  {
    static int pf_env = -1;
    if (pf_env < 0) pf_env = v2_env("YAD_CONV2_PFSMEM", 1);
    const int nops = (e->add ? 1 : 0) + (e->mul ? 1 : 0);
    if (pf_env && nops && !e->gn_stats && !e->img_scale && !e->pix_scale && p.sc <= 32) {
      const uint32_t want = (uint32_t)(p.ew * nops) * 2048u, floor_stages = patch ? 2u : 4u;
      uint32_t st2 = stages;
      while (st2 >= floor_stages && total(st2) + want > budget) st2--;
      if (st2 >= floor_stages) { stages = st2; pf_total = want; p.pf_warp_bytes = (uint32_t)nops * 2048u; }
    }
  }
  p.stages = (int)stages;
  p.off_b = 0;
  p.off_a = b_total;
  p.off_stg = p.off_a + stages * p.a_stage_bytes;
  p.off_pf = p.off_stg + stg_total;
  p.off_bias = p.off_pf + pf_total;
  p.off_bars = p.off_bias + bias_bytes;
  *smem_out = 1024 + total(stages) + pf_total;
  p.bias = e->bias; p.img_scale = e->img_scale; p.pix_scale = (const bf16*)e->pix_scale; p.pix_scale_ld = e->pix_scale_ld;
  p.act = e->act; p.alpha = e->alpha;
  p.mul = (const bf16*)e->mul; p.mul_ld = e->mul_ld; p.add = (const bf16*)e->add; p.add_ld = e->add_ld;
  p.gate_h = (const bf16*)e->gate_h; p.gate_w = (const bf16*)e->gate_w; p.gate_ld = e->gate_ld;
  p.gn_stats = e->gn_stats; p.gn_groups = e->gn_groups;
  *out = p;
  return patch ? 2 : 1;
}

int yad_conv2d_v2_supported(const yad_tensor* x, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y) {
  static int on = -1;
  if (on < 0) on = v2_env("YAD_CONV_V2", 1);
  if (!on && d->impl != 4) return 0;
  if (!get_encode()) return 0;
  V2Params p;
  size_t smem;
  return v2_plan(x, d, e, y, &p, &smem);
}

static int v2_run(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream,
                  const V2Phase* phase) {
  V2Params p;
  size_t smem = 0;
  const int kind = v2_plan(x, d, e, y, &p, &smem, phase);
  YAD_CHECK(kind != 0, "conv2d_v2: shape / epilogue not supported by the resident-weight tcgen05 kernel");
  const bool patch = kind == 2;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmA, tmB, tmY;
  const int w_row = d->kh * d->kw * x->c;
  if (patch) {
    uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
    uint64_t strides[3] = {(uint64_t)x->ld * 2, (uint64_t)x->w * x->ld * 2, (uint64_t)x->h * x->w * x->ld * 2};
    uint32_t box[4] = {64, (uint32_t)p.pw, (uint32_t)(V2_BH + 2), 1};
    if (v2_make_map(&tmA, x->ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  } else {
    uint64_t dims[2] = {(uint64_t)x->c, (uint64_t)p.m_total}, strides[1] = {(uint64_t)x->ld * 2};
    uint32_t box[2] = {64, V2_BM};
    if (v2_make_map(&tmA, x->ptr, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    const int cout_rows = (y->c + 7) / 8 * 8;
    uint64_t dims[2] = {(uint64_t)w_row, (uint64_t)cout_rows}, strides[1] = {(uint64_t)w_row * 2};
    uint32_t box[2] = {64, (uint32_t)p.n_tile};
    if (v2_make_map(&tmB, w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    const CUtensorMapSwizzle sw = p.sc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.sc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    if (phase) {  // every second pixel of every second row, starting at (py, px): the phase's pixels as a dense (c, w / 2, h / 2, n) tensor
      uint64_t dims[4] = {(uint64_t)y->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)y->n};
      uint64_t strides[3] = {(uint64_t)y->ld * 4, (uint64_t)y->w * y->ld * 4, (uint64_t)y->h * y->w * y->ld * 2};
      uint32_t box[4] = {(uint32_t)p.sc, V2_BW, 4, 1};
      const bf16* y0 = reinterpret_cast<const bf16*>(y->ptr) + ((int64_t)phase->py * y->w + phase->px) * y->ld;
      if (v2_make_map(&tmY, y0, 4, dims, strides, box, sw)) return 1;
    } else if (patch) {
      uint64_t dims[4] = {(uint64_t)y->c, (uint64_t)y->w, (uint64_t)y->h, (uint64_t)y->n};
      uint64_t strides[3] = {(uint64_t)y->ld * 2, (uint64_t)y->w * y->ld * 2, (uint64_t)y->h * y->w * y->ld * 2};
      uint32_t box[4] = {(uint32_t)p.sc, V2_BW, 4, 1};
      if (v2_make_map(&tmY, y->ptr, 4, dims, strides, box, sw)) return 1;
    } else {
      uint64_t dims[2] = {(uint64_t)y->c, (uint64_t)p.m_total}, strides[1] = {(uint64_t)y->ld * 2};
      uint32_t box[2] = {(uint32_t)p.sc, 32};
      if (v2_make_map(&tmY, y->ptr, 2, dims, strides, box, sw)) return 1;
    }
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int grid = num_sms < p.total_tiles ? num_sms : p.total_tiles;
  if (patch) return v2_dispatch<true>(p, tmA, tmB, tmY, grid, smem, st);
  return v2_dispatch<false>(p, tmA, tmB, tmY, grid, smem, st);
}

// The caller (yad_conv2d_tc) has validated shapes and zeroed the GroupNorm statistics buffer.
int yad_conv2d_v2(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream) {
  return v2_run(x, w, d, e, y, stream, nullptr);
}

// k3 s2 p1 op1 transposed convolution as four launches of conv2_kernel, one per output phase (oy = 2 iy - 1 + ky: even rows take ky = 1 at
// iy = m, odd rows ky = 2 at iy = m and ky = 0 at iy = m + 1; columns alike).  Every phase reads the input through the haloed-patch pipeline
// with its 1, 2 or 4 weight taps resident and stores its pixels with a strided TMA map.  Returns -1 when the call is not covered (the caller
// keeps the per-phase conv_tma_kernel launches), 0 when the four launches were issued, > 0 on error.
int yad_conv2d_v2_transposed(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream) {
  static int on = -1;
  if (on < 0) on = v2_env("YAD_CONV_V2", 1) && v2_env("YAD_CONV2_TRANSPOSED", 1);
  if ((!on && d->impl != 4) || !get_encode()) return -1;
  if (d->kh != 3 || d->kw != 3 || d->stride != 2 || d->pad_h != 1 || d->pad_w != 1) return -1;
  V2Phase ph[4];
  for (int py = 0; py < 2; py++)
    for (int px = 0; px < 2; px++) {
      V2Phase& f = ph[py * 2 + px];
      f.py = py; f.px = px; f.ntaps = 0;
      for (int ky = 0; ky < 3; ky++) {
        if ((py + 1 - ky) & 1) continue;
        for (int kx = 0; kx < 3; kx++) {
          if ((px + 1 - kx) & 1) continue;
          f.dy[f.ntaps] = (py + 1 - ky) / 2; f.dx[f.ntaps] = (px + 1 - kx) / 2; f.wtap[f.ntaps] = ky * 3 + kx;
          f.ntaps++;
        }
      }
      V2Params p;
      size_t smem;
      if (!v2_plan(x, d, e, y, &p, &smem, &f)) return -1;  // all four phases or none
    }
  for (int i = 0; i < 4; i++) {
    const int r = v2_run(x, w, d, e, y, stream, &ph[i]);
    if (r) return r;
  }
  return 0;
}

// Small-channel 3x3 stride-1 convolution on conv3_kernel.  Returns -1 when the call is not covered (the caller goes on to conv_small_kernel /
// the generic kernels), 0 when the launch was issued, > 0 on error.
template <int ACT, bool MULADD>
static int c3_launch_t(const C3Params& cp, const CUtensorMap& tmA, const CUtensorMap& tmY, int grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv3_kernel<ACT, MULADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024) != cudaSuccess) {
      yad_set_error("conv2d (small-channel tcgen05): cannot raise the dynamic shared memory limit");
      return 2;
    }
    attr_set = true;
  }
  YAD_LAUNCH((conv3_kernel<ACT, MULADD>), grid, 320, smem, st, cp, tmA, tmY);
  YAD_LAUNCH_CHECK("conv2d (small-channel tcgen05)");
  return 0;
}

int yad_conv2d_c3(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream) {
  static int on = -1;
  if (on < 0) on = v2_env("YAD_CONV_C3", 1);
  if (!on || !get_encode()) return -1;
  if (d->mode != YAD_CONV_NORMAL || d->kh != 3 || d->kw != 3 || d->pad_h != 1 || d->pad_w != 1 || (d->stride != 1 && d->stride != 2)) return -1;
  if (!(x->c == 8 || x->c == 16 || x->c == 32) || !(y->c == 8 || y->c == 16 || y->c == 32)) return -1;
  const bool s2 = d->stride == 2;
  if (s2 && (x->c != 16 || x->ld != 16 || (x->h & 1) || (x->w & 1))) return -1;
  if ((x->ld % 8) || (y->ld % 8) || ((uintptr_t)x->ptr & 15) || ((uintptr_t)y->ptr & 15) || ((uintptr_t)w & 15)) return -1;
  if (y->h != x->h / d->stride || y->w != x->w / d->stride || y->n != x->n) return -1;
  if (e->img_scale || e->pix_scale || e->mul || e->gn_stats) return -1;
  if (e->add && (((uintptr_t)e->add & 15) || (e->add_ld % 8))) return -1;
  const int64_t M = (int64_t)y->n * y->h * y->w;
  if (M + V2_BM >= (int64_t)1 << 31) return -1;
  C3Params cp;
  memset(&cp, 0, sizeof(cp));
  V2Params& p = cp.v;
  p.n = x->n; p.hm = y->h; p.wm = y->w; p.hw = y->h * y->w; p.cin = x->c; p.cout = y->c; p.m_total = (int)M;
  p.n_tile = y->c < 16 ? 16 : y->c;
  p.tiles_n = 1; p.ntaps = 9; p.kpt = 1; p.ksplit = 1;
  p.tiles_x = (p.wm + V2_BW - 1) / V2_BW;
  p.tiles_y = (p.hm + V2_BH - 1) / V2_BH;
  p.total_tiles = p.n * p.tiles_x * p.tiles_y;
  if (p.total_tiles < 4 * 148) return -1;  // small maps: the per-CTA weight layout pass does not amortise
  p.ew = 8; p.tsplit = 2; p.stg_bufs = 2;
  p.acc_stages = 256 / p.n_tile < V2_MAX_ACC ? 256 / p.n_tile : V2_MAX_ACC;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.acc_stages * p.n_tile) p.tmem_cols <<= 1;
  for (int i = 0; i <= 4; i++) p.bnd[i] = i == 0 ? 0 : p.n_tile;
  p.sc = p.n_tile;  // 16 or 32 columns per store box
  p.stg_warp_bytes = 32u * (uint32_t)p.sc * 2u;
  {
    static int dbg_env = -1;
    if (dbg_env < 0) dbg_env = (v2_env("YAD_CONV2_DBG", 0) & 7) | (v2_env("YAD_CONV2_PREFETCH", 1) ? 0 : 8);
    p.dbg = dbg_env;
  }
  // A-operand layout by input width (see the kernel's header)
  const int nchunk = 9 * (x->c / 8);  // 16-byte K chunks of the weight rows
  cp.nmma = (nchunk + 1) / 2;
  if (s2) {
    cp.s2 = 1;
    for (int t = 0; t < 9; t++) {  // tap (dy, dx): input pixel (2 oy + dy - 1, 2 ox + dx - 1) -> row-parity plane, row / pair offset inside the box, x parity
      const int dy = t / 3, dx = t % 3;
      const int plane = dy == 1 ? 0 : 1, r0 = dy == 0 ? 0 : 1, c0 = dx == 0 ? 0 : 1, xpar = dx == 1 ? 0 : 1;
      cp.a_off[t] = (uint32_t)(plane * C3_S2_PLANE + (r0 * 9 + c0) * 64 + xpar * 32);
      cp.a_lbo[t] = 16u;
    }
    cp.a_hi = ((9u * 64u) >> 4) | (1u << 14) | (4u << 29);
    cp.tx_bytes = 2u * 17u * 9u * 64u;
    cp.stage_bytes = 2u * C3_S2_PLANE;
  } else if (x->c == 8) {
    cp.dense = x->ld == 8;
    for (int i = 0; i < cp.nmma; i++) {
      const int t0 = 2 * i, t1 = 2 * i + 1;
      cp.a_off[i] = (uint32_t)(((t0 / 3) * 10 + (t0 % 3)) * 16);
      cp.a_lbo[i] = t1 < 9 ? (uint32_t)(((t1 / 3) * 10 + (t1 % 3)) * 16) - cp.a_off[i] : 0u;  // the odd last step re-reads its own chunk against zero weights
    }
    cp.a_hi = (160u >> 4) | (1u << 14);
    cp.tx_bytes = C3_PLANE_BYTES;
    cp.stage_bytes = C3_PLANE_STRIDE;
  } else {
    const uint32_t rb = (uint32_t)x->c * 2u;  // 32 or 64 bytes per pixel row
    for (int i = 0; i < cp.nmma; i++) {
      const int t = x->c == 16 ? i : i / 2, half = x->c == 16 ? 0 : i % 2;
      cp.a_off[i] = (uint32_t)((t / 3) * 10 + (t % 3)) * rb + (uint32_t)half * 32u;
      cp.a_lbo[i] = 16u;  // unused by the swizzled K-major layouts
    }
    cp.a_hi = ((10u * rb) >> 4) | (1u << 14) | ((x->c == 16 ? 6u : 4u) << 29);
    cp.tx_bytes = 180u * rb;
    cp.stage_bytes = (cp.tx_bytes + 1023u) & ~1023u;
  }
  cp.b_lbo = (uint32_t)(p.n_tile / 8) * 128u;
  cp.b_mma_bytes = 2u * cp.b_lbo;
  cp.w = (const bf16*)w;
  cp.w_row = 9 * x->c;
  p.stages = 6;
  while (p.stages > 3 && 1024u + (uint32_t)p.stages * cp.stage_bytes + (uint32_t)cp.nmma * cp.b_mma_bytes + 1024u + (uint32_t)(p.ew * p.stg_bufs) * p.stg_warp_bytes + 512u > 112u * 1024u)
    p.stages--;
  p.off_a = 0;
  cp.off_w = p.off_a + (uint32_t)p.stages * cp.stage_bytes;
  cp.off_w = (cp.off_w + 127u) & ~127u;
  p.off_stg = (cp.off_w + (uint32_t)cp.nmma * cp.b_mma_bytes + 1023u) & ~1023u;
  p.off_bias = p.off_stg + (uint32_t)(p.ew * p.stg_bufs) * p.stg_warp_bytes;
  p.off_bars = p.off_bias + 128u;
  const size_t smem = 1024 + p.off_bars + 8u * (2u * (uint32_t)p.stages + 2u * V2_MAX_ACC) + 16u;
  if (smem > 112 * 1024) return -1;
  p.bias = e->bias; p.act = e->act; p.alpha = e->alpha;
  p.add = (const bf16*)e->add; p.add_ld = e->add_ld;
  CUtensorMap tmA, tmY;
  if (cp.s2) {
    const uint64_t rowb = (uint64_t)x->w * 32;  // bytes of one image row (16 channels)
    uint64_t dims[5] = {32, (uint64_t)x->w / 2, 2, (uint64_t)x->h / 2, (uint64_t)x->n};
    uint64_t strides[4] = {64, rowb, 2 * rowb, (uint64_t)x->h * rowb};
    uint32_t box[5] = {32, 9, 1, 17, 1};
    if (v2_make_map(&tmA, x->ptr, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  } else if (cp.dense) {
    uint64_t dims[3] = {(uint64_t)x->w * 8, (uint64_t)x->h, (uint64_t)x->n};
    uint64_t strides[2] = {(uint64_t)x->w * 16, (uint64_t)x->h * x->w * 16};
    uint32_t box[3] = {80, (uint32_t)(V2_BH + 2), 1};
    if (v2_make_map(&tmA, x->ptr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  } else {
    uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
    uint64_t strides[3] = {(uint64_t)x->ld * 2, (uint64_t)x->w * x->ld * 2, (uint64_t)x->h * x->w * x->ld * 2};
    uint32_t box[4] = {(uint32_t)x->c, 10, (uint32_t)(V2_BH + 2), 1};
    const CUtensorMapSwizzle sw = x->c == 8 ? CU_TENSOR_MAP_SWIZZLE_NONE : (x->c == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (v2_make_map(&tmA, x->ptr, 4, dims, strides, box, sw)) return 1;
  }
  {
    const CUtensorMapSwizzle sw = p.sc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    uint64_t dims[4] = {(uint64_t)y->c, (uint64_t)y->w, (uint64_t)y->h, (uint64_t)y->n};
    uint64_t strides[3] = {(uint64_t)y->ld * 2, (uint64_t)y->w * y->ld * 2, (uint64_t)y->h * y->w * y->ld * 2};
    uint32_t box[4] = {(uint32_t)p.sc, V2_BW, 4, 1};
    if (v2_make_map(&tmY, y->ptr, 4, dims, strides, box, sw)) return 1;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int grid = 2 * num_sms < p.total_tiles ? 2 * num_sms : p.total_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  const bool muladd = e->add != nullptr;
  if (e->act == YAD_ACT_SILU) return muladd ? c3_launch_t<YAD_ACT_SILU, true>(cp, tmA, tmY, grid, smem, st) : c3_launch_t<YAD_ACT_SILU, false>(cp, tmA, tmY, grid, smem, st);
  if (e->act == YAD_ACT_NONE) return muladd ? c3_launch_t<YAD_ACT_NONE, true>(cp, tmA, tmY, grid, smem, st) : c3_launch_t<YAD_ACT_NONE, false>(cp, tmA, tmY, grid, smem, st);
  return c3_launch_t<ACT_GENERIC, true>(cp, tmA, tmY, grid, smem, st);
}

// Fused deformable 3x3 convolution (dcn2_kernel).  Returns -1 when the call is not eligible (the caller falls back to conv_tc_kernel<DEFORM>).
int yad_conv2d_dcn_v2(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream) {
  static int on = -1;
  if (on < 0) on = v2_env("YAD_DCN_V2", 1);
  if (!on || !get_encode()) return -1;
  if (d->mode != YAD_CONV_DEFORM || x->c != 64 || (x->ld % 8) || (y->c % 8) || (y->ld % 8) || y->c > 256 || !d->offmask || (d->offmask_ld % 8)) return -1;
  if (((uintptr_t)x->ptr & 15) || ((uintptr_t)y->ptr & 15) || ((uintptr_t)d->offmask & 15)) return -1;
  if (e->mul || e->add || e->img_scale || e->pix_scale || (e->gn_stats && e->gn_groups == 0)) return -1;
  const int64_t M = (int64_t)x->n * x->h * x->w;
  if (M + V2_BM >= (int64_t)1 << 31) return -1;
  DcnParams dp;
  memset(&dp, 0, sizeof(dp));
  V2Params& p = dp.v;
  p.n = x->n; p.hm = y->h; p.wm = y->w; p.hw = y->h * y->w; p.cin = 64; p.cout = y->c; p.m_total = (int)M;
  p.n_tile = pick_n_tile_v2(p.cout);
  p.tiles_n = 1; p.ntaps = 9; p.kpt = 1; p.ksplit = 1; p.acc_stages = 2;
  if (e->gn_stats) {
    if (e->gn_groups <= 0 || y->c % e->gn_groups) return -1;
    p.cpg = y->c / e->gn_groups;
    if (p.cpg != 4 && p.cpg != 8 && p.cpg != 16) return -1;
  }
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.n_tile) p.tmem_cols <<= 1;
  p.ew = 8; p.stg_bufs = 1;
  {
    const int u16 = p.n_tile / 16;
    p.bnd[0] = 0; p.bnd[1] = 16 * ((u16 + 1) / 2);
    for (int i = 2; i <= 4; i++) p.bnd[i] = p.n_tile;
    int sc = 64;
    for (int wv = 0; wv < 2; wv++)
      while (sc > 16 && ((p.bnd[wv + 1] - p.bnd[wv]) % sc)) sc >>= 1;
    p.sc = sc;
  }
  p.tiles_x = (p.wm + V2_BW - 1) / V2_BW;
  p.tiles_y = (p.hm + V2_BH - 1) / V2_BH;
  p.total_tiles = p.n * p.tiles_x * p.tiles_y;
  p.b_chunk_bytes = (uint32_t)p.n_tile * 128u;
  p.stg_warp_bytes = 32u * (uint32_t)p.sc * 2u;
  const uint32_t b_total = 9u * p.b_chunk_bytes, stg_total = 8u * p.stg_warp_bytes, bias_bytes = ((uint32_t)p.n_tile * 4u + 127u) & ~127u;
  const uint32_t par_bytes = 9u * V2_BM * 16u, budget = 227u * 1024u - 1024u;
  bool fits = false;
  for (int R = 2; R >= 1 && !fits; R--)
    for (int as = 2; as >= 2 && !fits; as--) {
      dp.R = R; dp.pwr = V2_BW + 2 * R; dp.phr = V2_BH + 2 * R; dp.aslots = as;
      dp.om_off = ((uint32_t)(dp.pwr * dp.phr) * 128u + 1023u) & ~1023u;
      dp.raw_stage_bytes = dp.om_off + 8192u;
      fits = b_total + 2u * dp.raw_stage_bytes + (uint32_t)as * 16384u + par_bytes + stg_total + bias_bytes + 8u * 26u + 16u <= budget;
    }
  if (!fits) return -1;
  dp.raw_tx_bytes = (uint32_t)(dp.pwr * dp.phr) * 128u + 8192u;
  p.off_b = 0;
  dp.off_raw = b_total;
  dp.off_aslot = dp.off_raw + 2u * dp.raw_stage_bytes;
  dp.off_par = dp.off_aslot + (uint32_t)dp.aslots * 16384u;
  p.off_stg = dp.off_par + par_bytes;
  p.off_bias = p.off_stg + stg_total;
  p.off_bars = p.off_bias + bias_bytes;
  const size_t smem = 1024 + (size_t)p.off_bars + 8 * 26 + 16;
  p.bias = e->bias; p.act = e->act; p.alpha = e->alpha;
  p.gn_stats = e->gn_stats; p.gn_groups = e->gn_groups;
  dp.x = (const bf16*)x->ptr; dp.x_ld = x->ld;
  CUtensorMap tmX, tmOM, tmB, tmY;
  {
    uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
    uint64_t strides[3] = {(uint64_t)x->ld * 2, (uint64_t)x->w * x->ld * 2, (uint64_t)x->h * x->w * x->ld * 2};
    uint32_t box[4] = {64, (uint32_t)dp.pwr, (uint32_t)dp.phr, 1};
    if (v2_make_map(&tmX, x->ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {  // offsets / mask logits: 32 channels of the (n, h, w, >= 27) view (channels beyond the view are zero-filled), dense rows of 64 bytes
    uint64_t dims[4] = {(uint64_t)32, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
    uint64_t strides[3] = {(uint64_t)d->offmask_ld * 2, (uint64_t)x->w * d->offmask_ld * 2, (uint64_t)x->h * x->w * d->offmask_ld * 2};
    uint32_t box[4] = {32, V2_BW, V2_BH, 1};
    if (v2_make_map(&tmOM, d->offmask, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  }
  {
    const int cout_rows = (y->c + 7) / 8 * 8;
    uint64_t dims[2] = {(uint64_t)(9 * 64), (uint64_t)cout_rows}, strides[1] = {(uint64_t)(9 * 64) * 2};
    uint32_t box[2] = {64, (uint32_t)p.n_tile};
    if (v2_make_map(&tmB, w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    const CUtensorMapSwizzle sw = p.sc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.sc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    uint64_t dims[4] = {(uint64_t)y->c, (uint64_t)y->w, (uint64_t)y->h, (uint64_t)y->n};
    uint64_t strides[3] = {(uint64_t)y->ld * 2, (uint64_t)y->w * y->ld * 2, (uint64_t)y->h * y->w * y->ld * 2};
    uint32_t box[4] = {(uint32_t)p.sc, V2_BW, 4, 1};
    if (v2_make_map(&tmY, y->ptr, 4, dims, strides, box, sw)) return 1;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaFuncSetAttribute(dcn2_kernel<ACT_GENERIC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      yad_set_error("conv2d (deformable, fused): cannot raise the dynamic shared memory limit");
      num_sms = 0;
      return 2;
    }
  }
  const int grid = num_sms < p.total_tiles ? num_sms : p.total_tiles;
  YAD_LAUNCH((dcn2_kernel<ACT_GENERIC, true>), grid, 32 * (2 + 8 + DCN_GW), smem, (cudaStream_t)stream, dp, tmX, tmOM, tmB, tmY);
  YAD_LAUNCH_CHECK("conv2d (deformable, fused)");
  return 0;
}
