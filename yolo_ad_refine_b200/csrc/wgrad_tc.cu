// tcgen05 / TMEM weight-gradient kernel of the training path (autograd of F.conv2d w.r.t. its weight).
//
//   dW[co][tap][ci] = sum over output pixels m of dy[m][co] * x[pix(m, tap)][ci]        (stride 1 or 2; any cin >= 16, cout % 16 == 0, <= 256)
//
// As a GEMM the reduction axis is the PIXEL axis, which is the strided one in NHWC: both operands are "MN-major" for the tensor core
// (channels contiguous, pixels strided).  tcgen05.mma takes MN-major operands straight from shared memory (instruction descriptor bits 15/16),
// and an NHWC box fetched by TMA with SWIZZLE_128B -- 64 channels x (BH x BW pixels) -- IS the canonical MN-major SWIZZLE_128B layout
// (rows of 128 B = 64 channels of one pixel, 8-pixel atoms of 1024 B), so no transposition is ever materialised:
//
//   D[128 rows][N cols] (TMEM, fp32)  +=  A[128][K=16 px] * B[K=16 px][N]
//     A rows  = two "row groups" (tap, 64-channel group of cin): two x boxes of the SAME pixel patch (shifted by the tap; conv padding and
//               ragged edges = TMA zero fill), 64 channels each, LBO = one box apart
//     B cols  = the cout channels of the dy box of that pixel patch (64-channel groups, LBO = one box apart)
//
// CTA (mg, s): M-tile group mg (up to 512 / N accumulators of 128 x N live in TMEM for the whole kernel) x split s of the pixel tiles.
//   warp 0 / lane 0 : TMA producer -- per pixel tile one dy load + one x load per row group into a 2-3 stage mbarrier ring
//   warp 1 / lane 0 : tcgen05.mma issuer (KB/16 K-steps per M-tile and stage), tcgen05.commit frees the stage; warp 1 owns the TMEM allocation
//   warps 2-5       : epilogue after the last pixel tile: tcgen05.ld -> fp32 red.global.add into dW (lanes = consecutive ci: coalesced)
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int WT_THREADS = 192;
constexpr int WT_MAX_RG = 16;   // row groups (tap, 64-channel group) per CTA
constexpr int WT_MAX_TAPS = 9;

struct WtParams {
  int n, h, w, cin, cout;       // (n, h, w) = OUTPUT pixel grid; input pixel = stride * output pixel + tap offset
  int stride;
  int ntaps, G;                 // G = cin / 64
  int dy[WT_MAX_TAPS], dx[WT_MAX_TAPS];
  int bw, bh, kb;               // pixel patch of one K block: kb = bw * bh (multiple of 16, <= 128)
  int tiles_x, tiles_y, total_tiles;
  int RG;                       // total row groups = ntaps * G
  int rg_per_cta;               // row groups handled by one CTA (even, except possibly the tail group)
  int mgroups, splits;
  int n_groups;                 // 64-channel groups of the dy box (ceil(cout / 64))
  int stages, tmem_cols;
  float* dw;
};

// MN-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units): start >> 4 | LBO (bytes between 64-element groups of the M/N axis) >> 4 << 16 | SBO (bytes between 8-row K atoms = 1024) >> 4
// << 32 | version 1 << 46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format BF16 (1) @7/@10, A and B MN-major (1) @15/@16, N >> 3 @17, M = 128 >> 4 @24
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(WT_THREADS) wgrad_tc_kernel(const __grid_constant__ WtParams p, const __grid_constant__ CUtensorMap tmX,
                                                             const __grid_constant__ CUtensorMap tmDy) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t box_bytes = (uint32_t)p.kb * 128u;  // one 64-channel box of the pixel patch
  const int mg = blockIdx.x % p.mgroups, split = blockIdx.x / p.mgroups;
  const int rg0 = mg * p.rg_per_cta;
  const int nrg = min(p.rg_per_cta, p.RG - rg0);     // row groups of this CTA
  const int nmt = (nrg + 1) >> 1;                    // M-tiles (two row groups each)
  const int slots = 2 * ((p.rg_per_cta + 1) >> 1);   // x boxes reserved per stage (even: the last M-tile may read one stale box)
  const uint32_t stage_bytes = (uint32_t)(p.n_groups + slots) * box_bytes;
  const uint32_t bars = base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
  const uint32_t done_bar = bars + 8u * (2 * p.stages);
  const uint32_t tmem_ptr_addr = done_bar + 8u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // this CTA's pixel tiles
  const int t_begin = (int)(((int64_t)p.total_tiles * split) / p.splits), t_end = (int)(((int64_t)p.total_tiles * (split + 1)) / p.splits);
  const int per_img = p.tiles_x * p.tiles_y;

  if (tid == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDy) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_sync();  // on-chip prologue done; from here on the kernel reads what its predecessors wrote

  if (warp == 0) {
    {  // warp-uniform producer loop, the elected lane issues (up to 18 TMA boxes per stage: a divergent single thread pays ~100 cycles for each)
      const bool leader = elect_one();
      const uint32_t tx_bytes = (uint32_t)(p.n_groups + nrg) * box_bytes;
      // division-free walk (tc_ptx.cuh): consecutive tiles advance the (tile x, tile y, image) digits by one
      TileDigits ti;
      ti.init(t_begin, 1, 1, p.tiles_x, p.tiles_y);
      RingPos sr;
      for (int tile = t_begin; tile < t_end; tile++, ti.next(1, p.tiles_x, p.tiles_y), sr.next(p.stages)) {
        const int img = ti.img, ty0 = ti.ty * p.bh, tx0 = ti.tx * p.bw;
        const int s = sr.idx;
        const uint32_t ph = sr.ph;
        const uint32_t st = base + s * stage_bytes;
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (leader) mbar_expect_tx(full_bar(s), tx_bytes);
        for (int g = 0; g < p.n_groups; g++)
          if (leader) tma_load_4d(st + g * box_bytes, &tmDy, full_bar(s), g * 64, tx0, ty0, img);
        for (int j = 0; j < nrg; j++) {
          const int rg = rg0 + j, tap = rg / p.G, cg = rg - tap * p.G;
          if (leader) tma_load_4d(st + (p.n_groups + j) * box_bytes, &tmX, full_bar(s), cg * 64, p.stride * tx0 + p.dx[tap], p.stride * ty0 + p.dy[tap], img);
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop (the whole warp walks it, the lane elect.sync picks issues): descriptor words stay in uniform registers and the
    // per-MMA work is two 32-bit adds -- a single divergent thread needed ~200 issue cycles per tcgen05.mma
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc_mn(p.cout);
      const uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);        // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t lo_flags = ((box_bytes >> 4) & 0x3FFFu) << 16;               // LBO = one 64-channel box
      RingPos sr;
      bool it0 = true;  // the first tile overwrites the accumulators
      for (int tile = t_begin; tile < t_end; tile++, sr.next(p.stages), it0 = false) {
        const int s = sr.idx;
        mbar_wait(full_bar(s), sr.ph);
        tc_fence_after();
        const uint32_t st = base + s * stage_bytes;
        const uint32_t b_lo = (((st) & 0x3FFFFu) >> 4) | lo_flags, a_lo0 = (((st + p.n_groups * box_bytes) & 0x3FFFFu) >> 4) | lo_flags;
        const int ks = p.kb / 16;
        for (int mt = 0; mt < nmt; mt++) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(mt * p.cout);
          const uint32_t a_lo = a_lo0 + (uint32_t)(2 * mt) * (box_bytes >> 4);
          if (leader) {
            if (it0) umma_bf16<false>(d_tmem, pack64(a_lo, hi), pack64(b_lo, hi), idesc);
            else umma_bf16<true>(d_tmem, pack64(a_lo, hi), pack64(b_lo, hi), idesc);
            for (int k = 1; k < ks; k++) umma_bf16<true>(d_tmem, pack64(a_lo + 128u * k, hi), pack64(b_lo + 128u * k, hi), idesc);  // 2048 B per K step
          }
        }
        if (leader) umma_commit(empty_bar(s));
      }
      if (leader) umma_commit(done_bar);
    }
    __syncwarp();
    tc_fence_before();
  } else if (t_end > t_begin) {
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read (warps 2..5 -> 2, 3, 0, 1)
    const int row = quarter * 32 + lane;
    mbar_wait(done_bar, 0u);
    tc_fence_after();
    for (int mt = 0; mt < nmt; mt++) {
      const int j = 2 * mt + (row >> 6);  // row group inside the CTA
      const bool valid = j < nrg;
      const int rg = rg0 + (valid ? j : 0), tap = rg / p.G, cg = rg - tap * p.G;
      const int ci = cg * 64 + (row & 63);
      const bool ok = valid && ci < p.cin;  // rows beyond cin hold the TMA zero fill of a partial 64-channel group
      float* dst = p.dw + (int64_t)tap * p.cin + ci;  // + co * ntaps * cin
      for (int c0 = 0; c0 < p.cout; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * p.cout + c0), v);
        if (ok) {
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const float f = __uint_as_float(v[i]);
            if (f != 0.f) atomicAdd(dst + (int64_t)(c0 + i) * p.ntaps * p.cin, f);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// NHWC view (n, h, w, c; pixel stride ld) as a 4-D bf16 tensor map, SWIZZLE_128B, zero fill out of bounds.  The box delivers 64 channels x bw x bh
// pixels taken every `step`-th pixel in x and y (TMA traversal stride): a stride-2 convolution's input patch arrives already decimated.
int make_nhwc_map(CUtensorMap* tm, const void* ptr, int64_t n, int64_t h, int64_t w, int c, int ld, int bw, int bh, int step) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(bw * step), (cuuint32_t)(bh * step), 1}, es[4] = {1, (cuuint32_t)step, (cuuint32_t)step, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

// pixel patch (bw x bh, a multiple of 16 pixels, at most kb_max) that wastes the fewest zero-filled pixels on an (h, w) map
void pick_patch_k(int h, int w, int kb_max, int* bw_out, int* bh_out) {
  double best = -1;
  int best_area = 0;
  *bw_out = 16; *bh_out = 1;
  for (int bw = 1; bw <= (w < 256 ? w : 256); bw++) {
    for (int bh = 1; bh <= (h < 256 ? h : 256) && bw * bh <= kb_max; bh++) {
      const int area = bw * bh;
      if (area % 16) continue;
      const int tx = (w + bw - 1) / bw, ty = (h + bh - 1) / bh;
      // useful pixels per fetched pixel, discounted for small K blocks (per-stage TMA / barrier overhead)
      const double eff = (double)h * w / ((double)tx * ty * area) * (area >= 64 ? 1.0 : 0.5 + area / 128.0);
      if (eff > best + 1e-9 || (eff > best - 1e-9 && area > best_area)) { best = eff; best_area = area; *bw_out = bw; *bh_out = bh; }
    }
  }
}

}  // namespace

// Returns 0 when the launch was issued, 1 on error, -1 when the geometry is not covered by this kernel (the caller falls back to mma.sync).
int yad_conv_wgrad_tc(const yad_tensor* x, const yad_tensor* dy, const yad_conv_desc* d, float* dw, void* stream) {
  const int st = d->stride;
  if ((st != 1 && st != 2) || d->kh * d->kw > WT_MAX_TAPS || d->pad_h != d->kh / 2 || d->pad_w != d->kw / 2) return -1;
  if (dy->c > 256 && dy->c % 16 == 0) {  // more than one N tile: one launch per channel window of dy (a view) and the matching rows of dW
    for (int c0 = 0; c0 < dy->c;) {
      int cn = dy->c - c0;
      if (cn > 256) cn = (cn >= 384) ? 192 : cn / 2 / 16 * 16;
      yad_tensor part = *dy;
      part.ptr = (void*)((bf16*)dy->ptr + c0);
      part.c = cn;
      const int r = yad_conv_wgrad_tc(x, &part, d, dw + (int64_t)c0 * d->kh * d->kw * x->c, stream);
      if (r != 0) return c0 == 0 ? r : 1;
      c0 += cn;
    }
    return 0;
  }
  if (x->c < 16 || x->c % 8 != 0 || dy->c % 16 != 0 || dy->c > 256 || dy->c < 16) return -1;
  if (dy->h != (x->h + 2 * d->pad_h - d->kh) / st + 1 || dy->w != (x->w + 2 * d->pad_w - d->kw) / st + 1) return -1;
  if (((uintptr_t)x->ptr & 15) || ((uintptr_t)dy->ptr & 15)) return -1;
  static int sm100 = -1;
  if (sm100 < 0) {
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    sm100 = major == 10 ? 1 : 0;
  }
  if (!sm100 || !get_encode()) return -1;
  WtParams p;
  p.cin = x->c; p.cout = dy->c; p.stride = st;
  p.ntaps = d->kh * d->kw;
  p.G = (x->c + 63) / 64;
  for (int t = 0; t < p.ntaps; t++) { p.dy[t] = t / d->kw - d->pad_h; p.dx[t] = t % d->kw - d->pad_w; }
  // 1x1: the pixel axis is one flat axis (no neighbourhood) -> (1, 1, n*h*w) grid, full 128-pixel K blocks whatever the map size
  const bool flat = p.ntaps == 1 && st == 1;
  int64_t gn = dy->n, gh = dy->h, gw = dy->w, xn = x->n, xh = x->h, xw = x->w;
  if (flat) { gw = gn * gh * gw; gn = 1; gh = 1; xw = gw; xn = 1; xh = 1; }
  if (gh * gw < 16 || gw > (1ll << 31) - 1) return -1;
  p.n = (int)gn; p.h = (int)gh; p.w = (int)gw;
  p.RG = p.ntaps * p.G;
  p.n_groups = (p.cout + 63) / 64;
  // M-tiles per CTA: all accumulators (128 x cout fp32 each) stay in TMEM (512 columns)
  const int mtiles = (p.RG + 1) / 2;
  int mpc = 512 / p.cout;
  if (mpc > WT_MAX_RG / 2) mpc = WT_MAX_RG / 2;
  if (mpc > mtiles) mpc = mtiles;
  p.mgroups = (mtiles + mpc - 1) / mpc;
  mpc = (mtiles + p.mgroups - 1) / p.mgroups;  // balance the groups
  p.rg_per_cta = 2 * mpc;
  int cols = mpc * p.cout, tc = 32;
  while (tc < cols) tc <<= 1;
  p.tmem_cols = tc;
  // pixel patch / pipeline depth: stage = (dy groups + x boxes) * kb * 128 B
  const int boxes = p.n_groups + p.rg_per_cta;
  int kb_max = 128;
  while (kb_max > 32 && 2 * boxes * kb_max * 128 > 200 * 1024) kb_max >>= 1;
  if (flat) { p.bw = kb_max; p.bh = 1; }
  else pick_patch_k(p.h, p.w < 128 ? p.w : 128, kb_max, &p.bw, &p.bh);
  if (st == 2 && (p.bw > 128 || p.bh > 128)) return -1;  // TMA box dimension limit (256) with traversal stride 2
  p.kb = p.bw * p.bh;
  if (p.kb % 16 || 2 * boxes * p.kb * 128 > 200 * 1024) return -1;
  p.stages = (200 * 1024) / (boxes * p.kb * 128);
  if (p.stages > 4) p.stages = 4;
  p.tiles_x = (p.w + p.bw - 1) / p.bw;
  p.tiles_y = (p.h + p.bh - 1) / p.bh;
  p.total_tiles = p.n * p.tiles_x * p.tiles_y;
  p.splits = 148 / p.mgroups;
  if (p.splits > p.total_tiles) p.splits = p.total_tiles;
  if (p.splits < 1) p.splits = 1;
  p.dw = dw;
  CUtensorMap tmX, tmDy;
  if (make_nhwc_map(&tmX, x->ptr, xn, xh, xw, x->c, x->ld, p.bw, p.bh, st) || make_nhwc_map(&tmDy, dy->ptr, gn, gh, gw, dy->c, dy->ld, p.bw, p.bh, 1))
    return -1;
  const size_t smem = (size_t)p.stages * boxes * p.kb * 128 + 8 * (2 * p.stages + 1) + 16 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  YAD_LAUNCH(wgrad_tc_kernel, p.mgroups * p.splits, WT_THREADS, smem, (cudaStream_t)stream, p, tmX, tmDy);
  YAD_LAUNCH_CHECK("conv_wgrad (tcgen05)");
  return 0;
}
