// Small-channel 3x3 convolution (cin, cout in {8, 16, 32}; stride 1 or 2; pad 1): the first backbone layers, their input gradients and the
// narrow convs of the C3k2 bottlenecks.  These layers are HBM-bound (K = 9 * cin <= 288, N <= 32): the implicit-GEMM tcgen05 kernel spends a
// 128 x N x 64 tile (and a TMA box per 16-32-byte row) on them, so this kernel reads every input byte once instead:
//   a persistent CTA walks tiles of 8 x 32 output pixels, stages the haloed input patch in shared memory with 128-bit loads (zero-filled
//   padding), keeps the whole weight matrix [cout][9 * cin] in shared memory, and each warp computes two 16-pixel row segments with
//   mma.sync.m16n8k16 (bf16 -> fp32): A fragments come straight from the patch through ldmatrix.x4 with per-row addresses (a K step of 16 =
//   two (tap, 8-channel group) slices, possibly of different taps), B fragments from the weight rows.  Epilogue: bias, activation, residual
//   add; the tile is transposed through shared memory so that global stores are 128-bit and coalesced.
#include "common.cuh"

namespace {

constexpr int CS_TH = 8, CS_TW = 32, CS_THREADS = 256;
__host__ __device__ constexpr int cs_pitch(int c) { return ((c >> 3) & 1) ? c + 16 : c + 8; }  // odd number of 16-byte slots per row

struct CsGeom {
  int n, hi, wi, cin, ho, wo, cout, stride;
  int x_ld, y_ld;
  int tiles_x, tiles_y, total_tiles;
  int ksteps;  // ceil(9 * cin / 16)
};

__device__ __forceinline__ void cs_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cs_ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cs_ldsm2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

template <int NT>  // n-tiles of 8 output channels (cout = 8 * NT)
__global__ void __launch_bounds__(CS_THREADS) conv_small_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                                                                    CsGeom g, const float* __restrict__ bias, int act, const bf16* __restrict__ add,
                                                                    int add_ld) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t cs_smem[];
  const int s = g.stride;
  const int PH = (CS_TH - 1) * s + 3, PW = (CS_TW - 1) * s + 3;
  const int cpitch = cs_pitch(g.cin);
  const int K = 9 * g.cin, kpad = g.ksteps * 16, wpitch = kpad + 8;  // (kpad + 8) * 2 B = odd multiple of 16 B: conflict-free ldmatrix rows
  bf16* patch = reinterpret_cast<bf16*>(cs_smem);              // two patch buffers: tile i + 1 streams in (cp.async) while tile i is computed
  const uint32_t patch_elems = (uint32_t)PH * PW * cpitch;
  bf16* ws = patch + 2 * (size_t)patch_elems;
  bf16* outs = ws + (size_t)NT * 8 * wpitch;  // [256 px][cout + 8]
  const int opitch = NT * 8 + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // weights -> shared memory once per CTA: row co holds k = tap * cin + ci, zero beyond 9 * cin
  for (int i = tid; i < NT * 8 * (kpad / 8); i += CS_THREADS) {
    const int co = i / (kpad / 8), k8 = (i % (kpad / 8)) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (k8 < K) v = *reinterpret_cast<const uint4*>(w + (size_t)co * K + k8);
    *reinterpret_cast<uint4*>(ws + (size_t)co * wpitch + k8) = v;
  }
  const uint32_t patch_s0 = (uint32_t)__cvta_generic_to_shared(patch), ws_s = (uint32_t)__cvta_generic_to_shared(ws);
  const int coct = g.cin >> 3, kgroups = 9 * coct;  // number of 8-wide K groups that carry data
  const int g4 = lane >> 2, q4 = lane & 3;
  // The chunk geometry of the staging loop does not depend on the tile: each thread keeps its (row, column, offsets) slots in registers, so
  // that per tile a slot costs two adds, the bounds test and one cp.async (the kernel was issue-bound on index arithmetic before).
  constexpr int MAXS = 9;  // 17 x 65 x 2 chunks / 256 threads
  int sl_py[MAXS], sl_px[MAXS], sl_g[MAXS];
  uint32_t sl_s[MAXS];
  const int nchunks = PH * PW * coct;
#pragma unroll
  for (int j = 0; j < MAXS; j++) {
    const int ch = tid + j * CS_THREADS;
    const int oc = ch % coct, pp = ch / coct, px = pp % PW, py = pp / PW;
    sl_py[j] = ch < nchunks ? py : -100000;  // fails every bounds test
    sl_px[j] = px;
    sl_g[j] = (py * g.wi + px) * g.x_ld + oc * 8;
    sl_s[j] = (uint32_t)((pp * cpitch + oc * 8) * 2);
  }
  auto stage = [&](int tile, int buf) {  // haloed input patch of `tile` -> patch buffer `buf`, zero-filled outside the image (src-size 0)
    const int tx = tile % g.tiles_x, ty = (tile / g.tiles_x) % g.tiles_y, img = tile / (g.tiles_x * g.tiles_y);
    const int iy0 = ty * CS_TH * s - 1, ix0 = tx * CS_TW * s - 1;
    const uint32_t dst0 = patch_s0 + (uint32_t)buf * patch_elems * 2u;
    const bf16* xb = x + (((int64_t)img * g.hi + iy0) * g.wi + ix0) * g.x_ld;
#pragma unroll
    for (int j = 0; j < MAXS; j++) {
      if (j * CS_THREADS < nchunks) {
        const int iy = iy0 + sl_py[j], ix = ix0 + sl_px[j];
        const bool ok = (unsigned)iy < (unsigned)g.hi && (unsigned)ix < (unsigned)g.wi;
        const bf16* src = ok ? xb + sl_g[j] : x;
        if (sl_py[j] > -100000)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + sl_s[j]), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ldmatrix.x4 row addresses of the A fragments: matrix m = lane >> 3 -> pixel half (m & 1), K group (m >> 1); row r = lane & 7.  The offset of
  // K step ks (its (tap, 8-channel group) for this lane's K half) is tile-independent: one register per K step.
  constexpr int MAXK = 18;  // 9 * 32 / 16
  const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8, mk = lane >> 4;
  uint32_t koff[MAXK];
#pragma unroll
  for (int ks = 0; ks < MAXK; ks++) {
    int kg = ks * 2 + mk;
    if (kg >= kgroups) kg = kgroups - 1;  // K padding: the weight rows are zero there, any finite operand will do
    const int tap = kg / coct, cg = kg - tap * coct, ky = tap / 3, kx = tap - ky * 3;
    koff[ks] = (uint32_t)(((ky * PW + kx) * cpitch + cg * 8) * 2);
  }
  if ((int)blockIdx.x < g.total_tiles) stage(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, it++) {
    const int tx = tile % g.tiles_x, ty = (tile / g.tiles_x) % g.tiles_y, img = tile / (g.tiles_x * g.tiles_y);
    const int oy0 = ty * CS_TH, ox0 = tx * CS_TW;
    const int buf = it & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // this tile's patch has landed; the previous tile (other patch buffer, output staging) is fully drained
    if (tile + (int)gridDim.x < g.total_tiles) stage(tile + gridDim.x, buf ^ 1);
    const uint32_t patch_s = patch_s0 + (uint32_t)buf * patch_elems * 2u;
    // each warp: two 16-pixel segments (tile row = warp, x halves 0 / 1)
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const int py = warp, px0 = half * 16;
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) acc[j][i] = 0.f;
      if (oy0 + py < g.ho && ox0 + px0 < g.wo) {
        const uint32_t a_row = patch_s + (uint32_t)(((py * s * PW) + (px0 + mrow) * s) * cpitch * 2);
#pragma unroll
        for (int ks = 0; ks < MAXK; ks++) {
          if (ks >= g.ksteps) break;
          uint32_t a[4];
          cs_ldsm4(a, a_row + koff[ks]);
#pragma unroll
          for (int j = 0; j < NT; j++) {
            uint32_t b0, b1;
            cs_ldsm2(b0, b1, ws_s + (uint32_t)(((j * 8 + (lane & 7)) * wpitch + ks * 16 + ((lane >> 3) & 1) * 8) * 2));
            cs_mma(acc[j], a, b0, b1);
          }
        }
      }
      // epilogue into the transposition buffer: rows g4 / g4 + 8 of this segment, channels j * 8 + 2 * q4 (+1)
#pragma unroll
      for (int j = 0; j < NT; j++) {
        const int co = j * 8 + 2 * q4;
        float v[4] = {acc[j][0], acc[j][1], acc[j][2], acc[j][3]};
        if (bias) { v[0] += bias[co]; v[1] += bias[co + 1]; v[2] += bias[co]; v[3] += bias[co + 1]; }
        apply_act_n<4>(v, act);
        const int p0 = py * CS_TW + px0 + g4;
        *reinterpret_cast<__nv_bfloat162*>(outs + (size_t)p0 * opitch + co) = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(outs + (size_t)(p0 + 8) * opitch + co) = __floats2bfloat162_rn(v[2], v[3]);
      }
    }
    __syncthreads();
    // coalesced 128-bit stores (+ residual add, applied to the bf16-rounded conv output like the tcgen05 epilogue does): slot j of a thread is
    // chunk tid + 256 j = (pixel, 8-channel group); NT is a power of two, so the decomposition is shifts
#pragma unroll
    for (int j = 0; j < NT; j++) {
      const int ch = tid + j * CS_THREADS;
      const int oc = ch & (NT - 1), pp = ch / NT, px = pp & (CS_TW - 1), py = pp / CS_TW;
      const int oy = oy0 + py, ox = ox0 + px;
      if (oy >= g.ho || ox >= g.wo) continue;
      const int64_t pix = ((int64_t)img * g.ho + oy) * g.wo + ox;
      uint4 u = *reinterpret_cast<const uint4*>(outs + (size_t)pp * opitch + oc * 8);
      if (add) {
        float a8[8], v8[8];
        load8(reinterpret_cast<const bf16*>(&u), v8);
        load8(add + pix * add_ld + oc * 8, a8);
#pragma unroll
        for (int i = 0; i < 8; i++) v8[i] += a8[i];
        store8(y + pix * g.y_ld + oc * 8, v8);
      } else {
        *reinterpret_cast<uint4*>(y + pix * g.y_ld + oc * 8) = u;
      }
    }
  }
}

}  // namespace

// Returns 0 when the launch was issued, 1 on error, -1 when the convolution is not covered (the caller uses the generic kernels).
int yad_conv2d_small(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y, void* stream) {
  if (d->mode != YAD_CONV_NORMAL || d->kh != 3 || d->kw != 3 || d->pad_h != 1 || d->pad_w != 1 || (d->stride != 1 && d->stride != 2)) return -1;
  if (!(x->c == 8 || x->c == 16 || x->c == 32) || !(y->c == 8 || y->c == 16 || y->c == 32)) return -1;
  // measured against the tcgen05 kernel (batch 128, tools/conv_probe.py): 1.5 - 2.4x faster everywhere except 16 -> 32 at stride 2 and
  // 32 -> 32, where the implicit-GEMM tile is well filled and the two tie
  if (x->c >= 16 && y->c == 32 && (d->stride == 2 || x->c == 32)) return -1;
  if (e->img_scale || e->pix_scale || e->mul || e->gn_stats || e->alpha != 1.0f) return -1;
  if (((uintptr_t)x->ptr & 15) || ((uintptr_t)y->ptr & 15) || ((uintptr_t)w & 15) || (e->add && ((uintptr_t)e->add & 15))) return -1;
  CsGeom g;
  g.n = x->n; g.hi = x->h; g.wi = x->w; g.cin = x->c; g.ho = y->h; g.wo = y->w; g.cout = y->c; g.stride = d->stride;
  g.x_ld = x->ld; g.y_ld = y->ld;
  g.tiles_x = cdiv(g.wo, CS_TW); g.tiles_y = cdiv(g.ho, CS_TH); g.total_tiles = g.n * g.tiles_x * g.tiles_y;
  g.ksteps = (9 * g.cin + 15) / 16;
  const int s = d->stride, NT = y->c / 8;
  const int PH = (CS_TH - 1) * s + 3, PW = (CS_TW - 1) * s + 3;
  const size_t smem = 2 * (size_t)PH * PW * cs_pitch(x->c) * 2 + (size_t)y->c * (g.ksteps * 16 + 8) * 2 + (size_t)CS_TH * CS_TW * (y->c + 8) * 2;
  if (smem > 160 * 1024) return -1;
  int grid = 148 * 2;
  if (grid > g.total_tiles) grid = g.total_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* xp = (const bf16*)x->ptr;
  const bf16* wp = (const bf16*)w;
  bf16* yp = (bf16*)y->ptr;
  const bf16* addp = (const bf16*)e->add;
#define CS_LAUNCH(N)                                                                                                        \
  {                                                                                                                         \
    static bool attr = false;                                                                                               \
    if (!attr) { cudaFuncSetAttribute(conv_small_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; } \
    YAD_LAUNCH(conv_small_kernel<N>, grid, CS_THREADS, smem, st, xp, wp, yp, g, e->bias, e->act, addp, e->add_ld);                 \
  }
  if (NT == 1) CS_LAUNCH(1) else if (NT == 2) CS_LAUNCH(2) else CS_LAUNCH(4)
#undef CS_LAUNCH
  YAD_LAUNCH_CHECK("conv2d (small-channel)");
  return 0;
}
