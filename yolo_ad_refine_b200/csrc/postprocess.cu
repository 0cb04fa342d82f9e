// a9: fused DFL softmax-expectation + make_anchors + dist2bbox + stride + class sigmoid (nn/modules/head.py:1181-1204,1236-1252)
// a10: batched NMS replacing utils/ops.py:163-312 + torchvision.ops.nms: candidate compaction in anchor order, exact top-max_nms
//      selection (radix select), per-image bitonic sort by (score desc, index asc), greedy suppression in 64-wide chunks whose
//      intra-chunk IoU bitmask is built with warp ballots.  All box arithmetic uses explicit round-to-nearest intrinsics (no FMA
//      contraction) so that keep decisions are bit-identical to the reference's fp32 CPU arithmetic.
#include "common.cuh"

namespace {

// =====================================================================================================================
// decode
// =====================================================================================================================
constexpr int MAX_LEVELS = 4;
struct DecodeLevels {
  const void* ptr[MAX_LEVELS];
  int64_t sb[MAX_LEVELS], sc[MAX_LEVELS], sa[MAX_LEVELS];
  int h[MAX_LEVELS], w[MAX_LEVELS], start[MAX_LEVELS + 1];
  float stride[MAX_LEVELS];
  int nl;
};

constexpr int DA = 32;  // anchors per block

template <typename T>
__global__ void __launch_bounds__(128) decode_kernel(DecodeLevels L, int nc, int reg_max, const float* __restrict__ proj, float* __restrict__ y,
                                                     int N) {
  pdl_sync();
  extern __shared__ float tile[];  // [DA anchors][no + 1] (odd pitch: conflict-free column reads), then dist[4][DA]
  const int no = 4 * reg_max + nc, pitch = no + 1;
  float* dist = tile + DA * pitch;
  const int b = blockIdx.y, a0 = blockIdx.x * DA;
  const int na = min(DA, N - a0);
  for (int l = 0; l < L.nl; l++) {  // levels this block touches
    const int lo = max(a0, L.start[l]), hi = min(a0 + na, L.start[l + 1]);
    if (lo >= hi) continue;
    const T* src = reinterpret_cast<const T*>(L.ptr[l]) + (int64_t)b * L.sb[l];
    const int cnt = hi - lo;
    if (L.sc[l] == 1 && (no & 7) == 0 && (L.sa[l] & 7) == 0) {  // channel-contiguous (NHWC): 128-bit loads, consecutive lanes -> consecutive channels
      const int vpr = no >> 3;
      for (int idx = threadIdx.x; idx < cnt * vpr; idx += blockDim.x) {
        const int al = idx / vpr, v8 = (idx - al * vpr) * 8;
        float v[8];
        load8(src + (int64_t)(lo - L.start[l] + al) * L.sa[l] + v8, v);
        float* dst = tile + ((lo - a0) + al) * pitch + v8;
#pragma unroll
        for (int i = 0; i < 8; i++) dst[i] = v[i];
      }
    } else if (L.sc[l] == 1) {
      for (int idx = threadIdx.x; idx < cnt * no; idx += blockDim.x) {
        int al = idx / no, ch = idx - al * no;
        tile[((lo - a0) + al) * pitch + ch] = ld1(src + (int64_t)(lo - L.start[l] + al) * L.sa[l] + ch);
      }
    } else {  // anchor-contiguous ((B, no, N) layout)
      for (int idx = threadIdx.x; idx < cnt * no; idx += blockDim.x) {
        int ch = idx / cnt, al = idx - ch * cnt;
        tile[((lo - a0) + al) * pitch + ch] = ld1(src + (int64_t)(lo - L.start[l] + al) * L.sa[l] + (int64_t)ch * L.sc[l]);
      }
    }
  }
  __syncthreads();
  // DFL: thread = (side, anchor)
  for (int idx = threadIdx.x; idx < 4 * DA; idx += blockDim.x) {
    int side = idx / DA, al = idx % DA;
    if (al < na) {
      const float* row = tile + al * pitch + side * reg_max;
      float m = -INFINITY;
      for (int k = 0; k < reg_max; k++) m = fmaxf(m, row[k]);
      float s = 0.f, e = 0.f;
      for (int k = 0; k < reg_max; k++) {
        float p = expf(row[k] - m);
        s += p;
        e = fmaf(p, proj[k], e);
      }
      dist[side * DA + al] = e / s;
    }
  }
  __syncthreads();
  float* yb = y + (int64_t)b * (4 + nc) * N;
  if (threadIdx.x < na) {
    const int al = threadIdx.x, a = a0 + al;
    int l = 0;
    while (l + 1 < L.nl && a >= L.start[l + 1]) l++;
    const int i = a - L.start[l];
    const float ax = (float)(i % L.w[l]) + 0.5f, ay = (float)(i / L.w[l]) + 0.5f, st = L.stride[l];
    const float x1 = ax - dist[0 * DA + al], y1 = ay - dist[1 * DA + al], x2 = ax + dist[2 * DA + al], y2 = ay + dist[3 * DA + al];
    yb[0 * (int64_t)N + a] = ((x1 + x2) / 2.0f) * st;
    yb[1 * (int64_t)N + a] = ((y1 + y2) / 2.0f) * st;
    yb[2 * (int64_t)N + a] = (x2 - x1) * st;
    yb[3 * (int64_t)N + a] = (y2 - y1) * st;
  }
  for (int idx = threadIdx.x; idx < nc * DA; idx += blockDim.x) {
    int c = idx / DA, al = idx % DA;
    if (al < na) yb[(int64_t)(4 + c) * N + a0 + al] = sigmoidf_(tile[al * pitch + 4 * reg_max + c]);
  }
}

// Channel-contiguous bf16 levels (the engine's NHWC head outputs), reg_max 16, nc % 8 == 0: 128 anchors per CTA of 256 threads.  The tile stays
// bf16 in shared memory (row pitch 4 * reg_max + nc + 8 halves = an odd number of 16-byte chunks: every phase below is bank-conflict free); all
// of a thread's 128-bit global loads are issued before the first one is used; the DFL rows are read as two 128-bit shared loads; the class phase
// keeps lanes = consecutive anchors so that every channel row of y is written in 128-byte segments.  decode_kernel (32 anchors, fp32 tile,
// scalar shared stores) measured 131 us for the 336 MB of the batch-64 step = 2.6 TB/s.
constexpr int DB = 128;  // anchors per block
template <int NL>
__global__ void __launch_bounds__(256) decode_nhwc_kernel(DecodeLevels L, int nc, const float* __restrict__ proj, float* __restrict__ y, int N) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t dsm[];
  constexpr int RM = 16;
  const int no = 4 * RM + nc, vpr = no >> 3, pitch = (no + 8) * 2;  // bytes
  float* dist = reinterpret_cast<float*>(dsm + DB * pitch);           // [4][DB]
  const int b = blockIdx.y, a0 = blockIdx.x * DB, na = min(DB, N - a0);
  {
    const int total = na * vpr;
    constexpr int MAXV = 10;  // 128 anchors x up to 20 vectors / 256 threads
    uint4 v[MAXV];
#pragma unroll
    for (int k = 0; k < MAXV; k++) {
      const int idx = threadIdx.x + k * 256;
      if (idx < total) {
        const int al = idx / vpr, v8 = (idx - al * vpr) * 8, a = a0 + al;
        int l = 0;
#pragma unroll
        for (int j = 1; j < NL; j++) l += (j < L.nl && a >= L.start[j]) ? 1 : 0;
        const bf16* src = reinterpret_cast<const bf16*>(L.ptr[l]) + (int64_t)b * L.sb[l] + (int64_t)(a - L.start[l]) * L.sa[l] + v8;
        v[k] = __ldg(reinterpret_cast<const uint4*>(src));
      }
    }
#pragma unroll
    for (int k = 0; k < MAXV; k++) {
      const int idx = threadIdx.x + k * 256;
      if (idx < total) {
        const int al = idx / vpr, v8 = (idx - al * vpr) * 8;
        *reinterpret_cast<uint4*>(dsm + al * pitch + v8 * 2) = v[k];
      }
    }
    for (int idx = threadIdx.x + MAXV * 256; idx < total; idx += 256) {  // nc > 96
      const int al = idx / vpr, v8 = (idx - al * vpr) * 8, a = a0 + al;
      int l = 0;
      for (int j = 1; j < L.nl; j++) l += a >= L.start[j] ? 1 : 0;
      const bf16* src = reinterpret_cast<const bf16*>(L.ptr[l]) + (int64_t)b * L.sb[l] + (int64_t)(a - L.start[l]) * L.sa[l] + v8;
      *reinterpret_cast<uint4*>(dsm + al * pitch + v8 * 2) = __ldg(reinterpret_cast<const uint4*>(src));
    }
  }
  __syncthreads();
  // DFL: (side, anchor) pairs, lanes = consecutive anchors
  for (int idx = threadIdx.x; idx < 4 * DB; idx += 256) {
    const int side = idx / DB, al = idx - side * DB;
    if (al < na) {
      float r[16];
      const uint4 u0 = *reinterpret_cast<const uint4*>(dsm + al * pitch + side * 32), u1 = *reinterpret_cast<const uint4*>(dsm + al * pitch + side * 32 + 16);
      const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&u0);
      const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&u1);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 f0 = __bfloat1622float2(h0[i]), f1 = __bfloat1622float2(h1[i]);
        r[2 * i] = f0.x; r[2 * i + 1] = f0.y; r[8 + 2 * i] = f1.x; r[8 + 2 * i + 1] = f1.y;
      }
      float m = r[0];
#pragma unroll
      for (int k = 1; k < 16; k++) m = fmaxf(m, r[k]);
      float sden = 0.f, e = 0.f;
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const float pk = __expf(r[k] - m);  // ex2.approx: ~2 ulp; the kernel was instruction-bound on 144 precise exp / div per anchor
        sden += pk;
        e = fmaf(pk, proj[k], e);
      }
      dist[side * DB + al] = e / sden;
    }
  }
  __syncthreads();
  float* yb = y + (int64_t)b * (4 + nc) * N;
  if (threadIdx.x < na) {
    const int al = threadIdx.x, a = a0 + al;
    int l = 0;
    while (l + 1 < L.nl && a >= L.start[l + 1]) l++;
    const int i = a - L.start[l];
    const float ax = (float)(i % L.w[l]) + 0.5f, ay = (float)(i / L.w[l]) + 0.5f, st = L.stride[l];
    const float x1 = ax - dist[0 * DB + al], y1 = ay - dist[1 * DB + al], x2 = ax + dist[2 * DB + al], y2 = ay + dist[3 * DB + al];
    yb[0 * (int64_t)N + a] = ((x1 + x2) / 2.0f) * st;
    yb[1 * (int64_t)N + a] = ((y1 + y2) / 2.0f) * st;
    yb[2 * (int64_t)N + a] = (x2 - x1) * st;
    yb[3 * (int64_t)N + a] = (y2 - y1) * st;
  }
  // classes: item = (class octet, anchor), lanes = consecutive anchors
  const int coct = nc >> 3;
  for (int idx = threadIdx.x; idx < coct * DB; idx += 256) {
    const int co = idx / DB, al = idx - co * DB;
    if (al < na) {
      const uint4 u = *reinterpret_cast<const uint4*>(dsm + al * pitch + (4 * RM + co * 8) * 2);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
      float* dst = yb + (int64_t)(4 + co * 8) * N + a0 + al;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 f = __bfloat1622float2(h[i]);
        dst[(int64_t)(2 * i) * N] = __fdividef(1.0f, 1.0f + __expf(-f.x));
        dst[(int64_t)(2 * i + 1) * N] = __fdividef(1.0f, 1.0f + __expf(-f.y));
      }
    }
  }
}

// =====================================================================================================================
// NMS
// =====================================================================================================================
constexpr int CH = 256;  // anchors per chunk

struct NmsWs {
  int32_t* chunk_cnt;   // [B][nchunks]      pass counts
  int32_t* chunk_cnt2;  // [B][nchunks][2]   (key > T, key == T)
  int32_t* meta;        // [B][4]: total_pass, Tkey, r (equals to take), n_final
  float* best_sc;       // [B][N]  single-label: best score (or -1)
  int32_t* best_cls;    // [B][N]
  uint64_t* key;        // [B][cap_pad]
  float4* box;          // [B][cap]
  float4* sbox;         // [B][cap]  class-offset boxes in sorted order (written by the sort kernel)
  float* sc;            // [B][cap]
  int32_t* cls;         // [B][cap]
  int32_t* anchor;      // [B][cap]
  int cap, cap_pad, nchunks;
};

__device__ __forceinline__ uint32_t fkey(float f) {  // order-preserving float -> uint
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ int block_sum_int(int v, int* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  int r = (lane < nw) ? red[lane] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// exclusive block scan (blockDim <= 1024); returns the exclusive prefix for this thread, *total gets the block sum
__device__ __forceinline__ int block_exscan(int v, int* red, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) red[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < nw) ? red[lane] : 0, winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    red[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) red[32] = winc;
  }
  __syncthreads();
  *total = red[32];
  return inc - v + red[wid];
}

__device__ __forceinline__ bool class_ok(const uint8_t* cm, int c) { return cm == nullptr || cm[c] != 0; }

// K1: per-chunk pass counts; single-label mode also caches (best score, best class) per anchor
template <bool MULTI>
__global__ void __launch_bounds__(CH) nms_count_kernel(const float* __restrict__ pred, int nc, int N, float conf, const uint8_t* __restrict__ cm,
                                                       NmsWs ws) {
  pdl_sync();
  __shared__ int red[33];
  const int b = blockIdx.y, a = blockIdx.x * CH + threadIdx.x;
  const float* pb = pred + (int64_t)b * (4 + nc) * N;
  int cnt = 0;
  if (a < N) {
    if (MULTI) {
      for (int c = 0; c < nc; c++) cnt += (pb[(int64_t)(4 + c) * N + a] > conf && class_ok(cm, c)) ? 1 : 0;
    } else {
      float best = pb[(int64_t)4 * N + a];
      int bc = 0;
      for (int c = 1; c < nc; c++) {
        float s = pb[(int64_t)(4 + c) * N + a];
        if (s > best) { best = s; bc = c; }
      }
      const bool ok = best > conf && class_ok(cm, bc);
      ws.best_sc[(int64_t)b * N + a] = ok ? best : -1.0f;
      ws.best_cls[(int64_t)b * N + a] = bc;
      cnt = ok ? 1 : 0;
    }
  }
  cnt = block_sum_int(cnt, red);
  if (threadIdx.x == 0) ws.chunk_cnt[b * ws.nchunks + blockIdx.x] = cnt;
}

// iterate over the candidate (score) stream of image b in a block-strided fashion
template <bool MULTI, typename F>
__device__ __forceinline__ void for_each_score(const float* pb, const float* best, int nc, int N, float conf, const uint8_t* cm, F f) {
  if (MULTI) {
    const int64_t total = (int64_t)nc * N;
    for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
      int c = (int)(i / N);
      float s = pb[(int64_t)4 * N + i];
      if (s > conf && class_ok(cm, c)) f(s);
    }
  } else {
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
      float s = best[a];
      if (s >= 0.f) f(s);
    }
  }
}

// K2: one block per image.  If more than max_nms candidates pass, find the exact key T of the max_nms-th largest score (4 x 8-bit
// radix select) and r = how many candidates equal to T are still taken (in index order).
template <bool MULTI>
__global__ void __launch_bounds__(1024) nms_select_kernel(const float* __restrict__ pred, int nc, int N, float conf,
                                                          const uint8_t* __restrict__ cm, int max_nms, NmsWs ws) {
  pdl_sync();
  __shared__ int hist[256];
  __shared__ int red[33];
  __shared__ uint32_t s_prefix;
  __shared__ int s_k;
  const int b = blockIdx.x;
  int tot = 0;
  for (int i = threadIdx.x; i < ws.nchunks; i += blockDim.x) tot += ws.chunk_cnt[b * ws.nchunks + i];
  tot = block_sum_int(tot, red);
  int32_t* meta = ws.meta + b * 4;
  if (tot <= max_nms) {
    if (threadIdx.x == 0) { meta[0] = tot; meta[1] = 0; meta[2] = 0; meta[3] = tot; }
    return;
  }
  const float* pb = pred + (int64_t)b * (4 + nc) * N;
  const float* best = ws.best_sc + (int64_t)b * N;
  if (threadIdx.x == 0) { s_prefix = 0; s_k = max_nms; }
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for_each_score<MULTI>(pb, best, nc, N, conf, cm, [&](float s) {
      uint32_t k = fkey(s);
      if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
    });
    __syncthreads();
    if (threadIdx.x == 0) {
      int k = s_k, bin = 255;
      for (; bin > 0; bin--) {
        if (hist[bin] >= k) break;
        k -= hist[bin];
      }
      s_k = k;
      s_prefix = prefix | ((uint32_t)bin << shift);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { meta[0] = tot; meta[1] = (int32_t)s_prefix; meta[2] = s_k; meta[3] = max_nms; }
}

// K3: per-chunk counts of (key > T) and (key == T)
template <bool MULTI>
__global__ void __launch_bounds__(CH) nms_count2_kernel(const float* __restrict__ pred, int nc, int N, float conf, const uint8_t* __restrict__ cm,
                                                        NmsWs ws) {
  pdl_sync();
  __shared__ int red[33];
  const int b = blockIdx.y, a = blockIdx.x * CH + threadIdx.x;
  const uint32_t Tk = (uint32_t)ws.meta[b * 4 + 1];
  const float* pb = pred + (int64_t)b * (4 + nc) * N;
  int g = 0, e = 0;
  if (a < N) {
    if (MULTI) {
      for (int c = 0; c < nc; c++) {
        float s = pb[(int64_t)(4 + c) * N + a];
        if (s > conf && class_ok(cm, c)) { uint32_t k = fkey(s); g += k > Tk; e += k == Tk; }
      }
    } else {
      float s = ws.best_sc[(int64_t)b * N + a];
      if (s >= 0.f) { uint32_t k = fkey(s); g = k > Tk; e = k == Tk; }
    }
  }
  g = block_sum_int(g, red);
  e = block_sum_int(e, red);
  if (threadIdx.x == 0) {
    ws.chunk_cnt2[(b * ws.nchunks + blockIdx.x) * 2] = g;
    ws.chunk_cnt2[(b * ws.nchunks + blockIdx.x) * 2 + 1] = e;
  }
}

// K4: ordered gather of the selected candidates (anchor-major, class-minor)
template <bool MULTI>
__global__ void __launch_bounds__(CH) nms_gather_kernel(const float* __restrict__ pred, int nc, int N, float conf, const uint8_t* __restrict__ cm,
                                                        NmsWs ws) {
  pdl_sync();
  __shared__ int red[33];
  __shared__ int s_base[2];
  const int b = blockIdx.y, a = blockIdx.x * CH + threadIdx.x;
  const uint32_t Tk = (uint32_t)ws.meta[b * 4 + 1];
  const int r = ws.meta[b * 4 + 2];
  const float* pb = pred + (int64_t)b * (4 + nc) * N;
  // prefix over earlier chunks
  int pg = 0, pe = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += blockDim.x) {
    pg += ws.chunk_cnt2[(b * ws.nchunks + i) * 2];
    pe += ws.chunk_cnt2[(b * ws.nchunks + i) * 2 + 1];
  }
  pg = block_sum_int(pg, red);
  pe = block_sum_int(pe, red);
  int g = 0, e = 0;
  if (a < N) {
    if (MULTI) {
      for (int c = 0; c < nc; c++) {
        float s = pb[(int64_t)(4 + c) * N + a];
        if (s > conf && class_ok(cm, c)) { uint32_t k = fkey(s); g += k > Tk; e += k == Tk; }
      }
    } else {
      float s = ws.best_sc[(int64_t)b * N + a];
      if (s >= 0.f) { uint32_t k = fkey(s); g = k > Tk; e = k == Tk; }
    }
  }
  int tot_e, tot_i;
  int e_off = block_exscan(e, red, &tot_e);
  int eq_rank = pe + e_off;                       // global rank of this thread's first "equal" item
  int inc_e = max(0, min(e, r - eq_rank));        // equals still taken
  int my = g + inc_e;
  int off = block_exscan(my, red, &tot_i);
  int pos = pg + min(pe, r) + off;
  if (my == 0) return;
  // box (xywh -> xyxy, utils/ops.py:412-431)
  const float cx = pb[a], cy = pb[(int64_t)N + a], w = pb[(int64_t)2 * N + a], h = pb[(int64_t)3 * N + a];
  const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
  const float4 bx = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
  const int64_t base = (int64_t)b * ws.cap;
  auto emit = [&](float s, int c) {
    if (pos < ws.cap) {
      ws.box[base + pos] = bx;
      ws.sc[base + pos] = s;
      ws.cls[base + pos] = c;
      ws.anchor[base + pos] = a;
      ws.key[(int64_t)b * ws.cap_pad + pos] = ((uint64_t)(~fkey(s)) << 32) | (uint32_t)pos;
    }
    pos++;
  };
  if (MULTI) {
    int er = eq_rank;
    for (int c = 0; c < nc; c++) {
      float s = pb[(int64_t)(4 + c) * N + a];
      if (s > conf && class_ok(cm, c)) {
        uint32_t k = fkey(s);
        if (k > Tk) emit(s, c);
        else if (k == Tk) { if (er < r) emit(s, c); er++; }
      }
    }
  } else {
    emit(ws.best_sc[(int64_t)b * N + a], ws.best_cls[(int64_t)b * N + a]);
  }
  (void)tot_e; (void)tot_i; (void)s_base;
}

// K5: per-image bitonic sort of the 64-bit keys (ascending = score descending, index ascending), then the gather of the class-offset boxes in
// sorted order (ws.sbox) so that the greedy pass streams them with one coalesced, prefetchable load per 64-candidate chunk.
//   np <= 1024          one key per thread, compare-exchange through shared memory
//   2048 .. 16384       E = np / 1024 keys per thread IN REGISTERS: strides inside a thread are register swaps, strides inside a warp are
//                       shuffles, and only the five top index bits need shared memory -- two transpositions per merge size (layout A: index =
//                       thread * E + r; layout B: index = r << 10 | lane << 5 | warp) instead of one block-wide pass per stride (105 passes of
//                       16 K keys measured 208 us per image batch; see profiles/r2_nms.md)
//   larger              in global memory / L2 (multi_label lists of up to max_nms = 30000 keys)
constexpr int SORT_SMEM = 16384;  // keys held in (dynamic) shared memory: 128 KB
__device__ __forceinline__ int sort_swz(int i) { return i ^ ((i >> 5) & 15); }  // conflict-free for both layouts (8-byte banks)
__device__ __forceinline__ uint64_t shfl_xor64(uint64_t v, int m) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return ((uint64_t)hi << 32) | lo;
}

// compare-exchange helpers: `up` / `keep_min` are uniform over a thread's registers wherever the merge size exceeds the thread's own span
__device__ __forceinline__ void cmpx(uint64_t& lo, uint64_t& hi, bool up) {  // up: the lower index receives the minimum
  const uint64_t x = lo, y = hi;
  const bool sw = (y < x) == up;
  lo = sw ? y : x;
  hi = sw ? x : y;
}
__device__ __forceinline__ void cmps(uint64_t& v, int lane_mask, bool keep_min) {  // partner in another lane of the warp
  const uint64_t y = shfl_xor64(v, lane_mask);
  if ((y < v) == keep_min) v = y;
}

// layout A (index = thread * E + r): stages on index bits hi_bit .. 0 of merge size 2^sb, sb >= LOGE (direction uniform per thread)
template <int E, int LOGE>
__device__ __forceinline__ void sort_stages_a(uint64_t (&v)[E], int sb, int hi_bit, int t, int lane) {
  const bool up = ((t >> (sb - LOGE)) & 1) == 0;
  for (int jb = hi_bit; jb >= LOGE; jb--) {  // across lanes
    const int lm = 1 << (jb - LOGE);
    const bool keep_min = ((lane & lm) != 0) != up;
#pragma unroll
    for (int r = 0; r < E; r++) cmps(v[r], lm, keep_min);
  }
#pragma unroll
  for (int jb = LOGE - 1; jb >= 0; jb--) {  // inside the thread (hi_bit >= LOGE - 1 for every sb >= LOGE)
#pragma unroll
    for (int r = 0; r < E; r++)
      if (!(r & (1 << jb))) cmpx(v[r], v[r | (1 << jb)], up);
  }
}

template <int E>
__device__ __forceinline__ void sort_regs(uint64_t* sk, const uint64_t* gk) {
  constexpr int LOGE = E == 2 ? 1 : E == 4 ? 2 : E == 8 ? 3 : 4;
  constexpr int LOGN = 10 + LOGE, NP = 1024 * E;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  uint64_t v[E];
  for (int i = t; i < NP; i += 1024) sk[sort_swz(i)] = gk[i];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < E; r++) v[r] = sk[sort_swz(t * E + r)];
  // merge sizes below the thread's span: directions are compile-time constants
#pragma unroll
  for (int sb = 1; sb < LOGE; sb++) {
#pragma unroll
    for (int jb = sb - 1; jb >= 0; jb--) {
#pragma unroll
      for (int r = 0; r < E; r++)
        if (!(r & (1 << jb))) cmpx(v[r], v[r | (1 << jb)], ((r >> sb) & 1) == 0);
    }
  }
  for (int sb = LOGE; sb <= LOGN; sb++) {
    if (sb - 1 > LOGE + 4) {
      // index bits sb-1 .. LOGE+5 in layout B (index = r << 10 | lane << 5 | warp: bits 10.. are registers, bits 5..9 lanes)
      __syncthreads();
#pragma unroll
      for (int r = 0; r < E; r++) sk[sort_swz(t * E + r)] = v[r];
      __syncthreads();
#pragma unroll
      for (int r = 0; r < E; r++) v[r] = sk[sort_swz((r << 10) | (lane << 5) | warp)];
      const bool lane_up = sb >= 10 ? true : ((lane >> (sb - 5)) & 1) == 0;
      const int rbit = sb >= 10 && sb - 10 < LOGE ? (1 << (sb - 10)) : 0;  // register bit that flips the direction (0: none)
#pragma unroll
      for (int jb = LOGN - 1; jb >= 10; jb--) {
        if (jb > sb - 1) continue;
#pragma unroll
        for (int r = 0; r < E; r++)
          if (!(r & (1 << (jb - 10)))) cmpx(v[r], v[r | (1 << (jb - 10))], lane_up && !(r & rbit));
      }
      for (int jb = (sb - 1 < 9 ? sb - 1 : 9); jb >= LOGE + 5; jb--) {
        const int lm = 1 << (jb - 5);
        const bool upper = (lane & lm) != 0;
#pragma unroll
        for (int r = 0; r < E; r++) cmps(v[r], lm, upper != (lane_up && !(r & rbit)));
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < E; r++) sk[sort_swz((r << 10) | (lane << 5) | warp)] = v[r];
      __syncthreads();
#pragma unroll
      for (int r = 0; r < E; r++) v[r] = sk[sort_swz(t * E + r)];
      sort_stages_a<E, LOGE>(v, sb, LOGE + 4, t, lane);
    } else {
      sort_stages_a<E, LOGE>(v, sb, sb - 1, t, lane);
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < E; r++) sk[sort_swz(t * E + r)] = v[r];
  __syncthreads();
}

__global__ void __launch_bounds__(1024) nms_sort_kernel(NmsWs ws, int agnostic, float max_wh) {
  pdl_sync();
  extern __shared__ uint64_t sk[];
  const int b = blockIdx.x;
  const int n = ws.meta[b * 4 + 3];
  if (n <= 0) return;
  int np = 2;
  while (np < n) np <<= 1;
  uint64_t* gk = ws.key + (int64_t)b * ws.cap_pad;
  const bool in_smem = np <= SORT_SMEM;
  bool swizzled = false;
  if (n > 1) {
    for (int i = n + threadIdx.x; i < np; i += blockDim.x) gk[i] = ~0ull;
    __syncthreads();
    if (in_smem && np >= 2048) {
      swizzled = true;
      switch (np >> 10) {
        case 2: sort_regs<2>(sk, gk); break;
        case 4: sort_regs<4>(sk, gk); break;
        case 8: sort_regs<8>(sk, gk); break;
        default: sort_regs<16>(sk, gk); break;
      }
    } else {
      uint64_t* k = gk;
      if (in_smem) {
        for (int i = threadIdx.x; i < np; i += blockDim.x) sk[i] = gk[i];
        k = sk;
        __syncthreads();
      }
      for (int size = 2; size <= np; size <<= 1) {
        for (int j = size >> 1; j > 0; j >>= 1) {
          for (int i = threadIdx.x; i < np; i += blockDim.x) {
            int p = i ^ j;
            if (p > i) {
              uint64_t x = k[i], y = k[p];
              bool up = (i & size) == 0;
              if ((x > y) == up) { k[i] = y; k[p] = x; }
            }
          }
          __syncthreads();
        }
      }
    }
  }
  // sorted keys back to global memory + the class-offset boxes in sorted order
  const int64_t base = (int64_t)b * ws.cap;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    uint64_t key = gk[i];
    if (n > 1 && in_smem) {
      key = sk[swizzled ? sort_swz(i) : i];
      gk[i] = key;
    }
    const int ci = (int)(key & 0xFFFFFFFFu);
    float4 bx = ws.box[base + ci];
    if (!agnostic) {
      float off = __fmul_rn((float)ws.cls[base + ci], max_wh);  // utils/ops.py:285 (fp32 class offset)
      bx.x = __fadd_rn(bx.x, off); bx.y = __fadd_rn(bx.y, off); bx.z = __fadd_rn(bx.z, off); bx.w = __fadd_rn(bx.w, off);
    }
    ws.sbox[base + i] = bx;
  }
}

__device__ __forceinline__ bool iou_gt(const float4& a, float area_a, const float4& b, float area_b, float thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return ovr > thr;
}

// K6: greedy suppression, one block of GREEDY_T threads per image, 64 sorted candidates per round.  Inside the loop there is no dependent global
// access: the next chunk's boxes are prefetched into registers, the serial part works on two 64-bit masks in shared memory, and the rows of
// the kept candidates are written by all threads after the loop.
constexpr int GREEDY_T = 1024;
__global__ void __launch_bounds__(GREEDY_T) nms_greedy_kernel(NmsWs ws, float iou_thres, int max_det, float* __restrict__ out,
                                                              int32_t* __restrict__ out_idx, int32_t* __restrict__ out_count) {
  pdl_sync();
  extern __shared__ float4 kept[];                 // [max_det] offset boxes, float areas[max_det], int sorted rank[max_det]
  float* kept_area = reinterpret_cast<float*>(kept + max_det);
  int* kept_rank = reinterpret_cast<int*>(kept_area + max_det);
  __shared__ float4 cb[64];
  __shared__ float ca[64];
  __shared__ unsigned int dead[2];                 // suppressed-by-kept bitmask
  __shared__ unsigned long long sup[64];           // sup[i] = mask of j > i suppressed by i
  __shared__ int s_nk;
  __shared__ unsigned int s_keep[2];
  const int b = blockIdx.x, n = ws.meta[b * 4 + 3];
  const uint64_t* key = ws.key + (int64_t)b * ws.cap_pad;
  const int64_t base = (int64_t)b * ws.cap;
  if (threadIdx.x == 0) s_nk = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int PARTS = GREEDY_T / 64, IPW = 64 / (GREEDY_T / 32);
  float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x < 64 && threadIdx.x < n) nxt = ws.sbox[base + threadIdx.x];
  __syncthreads();
  for (int s = 0; s < n; s += 64) {
    const int m = min(64, n - s);
    const int nk = s_nk;
    if (nk >= max_det) break;
    if (threadIdx.x < 64) {
      const float4 bx = nxt;
      if (s + 64 + threadIdx.x < n) nxt = ws.sbox[base + s + 64 + threadIdx.x];
      cb[threadIdx.x] = bx;
      ca[threadIdx.x] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
      if (threadIdx.x < 2) dead[threadIdx.x] = 0u;
    }
    __syncthreads();
    // A: against the kept list
    {
      const int j = threadIdx.x & 63, part = threadIdx.x >> 6;
      bool d = false;
      if (j < m) {
        const float4 bj = cb[j];
        const float aj = ca[j];
        for (int k = part; k < nk && !d; k += PARTS) d = iou_gt(kept[k], kept_area[k], bj, aj, iou_thres);
      }
      if (d) atomicOr(&dead[j >> 5], 1u << (j & 31));
    }
    // B: intra-chunk bitmask with warp ballots: warp w handles i in [IPW w, IPW w + IPW)
    for (int ii = 0; ii < IPW; ii++) {
      const int i = warp * IPW + ii;
      for (int half = 0; half < 2; half++) {
        const int j = half * 32 + lane;
        bool p = false;
        if (i < m && j < m && j > i) p = iou_gt(cb[i], ca[i], cb[j], ca[j], iou_thres);
        unsigned int bal = __ballot_sync(0xffffffffu, p);
        if (lane == 0) reinterpret_cast<unsigned int*>(sup)[i * 2 + half] = bal;
      }
    }
    __syncthreads();
    // C: serial resolve on the bitmasks alone: one thread hops from kept candidate to kept candidate
    if (threadIdx.x == 0) {
      const uint64_t valid = m == 64 ? ~0ull : ((1ull << m) - 1ull);
      uint64_t remv = ((uint64_t)dead[1] << 32) | dead[0], keep = 0ull;
      uint64_t avail = ~remv & valid;
      int k = nk;
      while (avail && k < max_det) {
        const int i = __ffsll((long long)avail) - 1;
        keep |= 1ull << i;
        k++;
        remv |= sup[i];
        avail = ~remv & valid & ~((2ull << i) - 1ull);
      }
      s_keep[0] = (unsigned int)keep;
      s_keep[1] = (unsigned int)(keep >> 32);
      s_nk = k;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int i = threadIdx.x;
      const uint64_t keep = ((uint64_t)s_keep[1] << 32) | s_keep[0];
      if ((keep >> i) & 1ull) {
        const int k = nk + __popcll(keep & ((1ull << i) - 1ull));
        kept[k] = cb[i];
        kept_area[k] = ca[i];
        kept_rank[k] = s + i;
      }
    }
    __syncthreads();
  }
  const int nk = s_nk;
  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    const int ci = (int)(key[kept_rank[k]] & 0xFFFFFFFFu);
    const float4 ob = ws.box[base + ci];
    const int cl = ws.cls[base + ci];
    float* o = out + ((int64_t)b * max_det + k) * 6;
    o[0] = ob.x; o[1] = ob.y; o[2] = ob.z; o[3] = ob.w; o[4] = ws.sc[base + ci]; o[5] = (float)cl;
    out_idx[((int64_t)b * max_det + k) * 2] = ws.anchor[base + ci];
    out_idx[((int64_t)b * max_det + k) * 2 + 1] = cl;
  }
  if (threadIdx.x == 0) out_count[b] = nk;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int nms_cap(int N, int nc, int multi_label, int max_nms) {
  int64_t c = (int64_t)N * (multi_label ? nc : 1);
  return (int)(c < max_nms ? c : max_nms);
}

size_t carve(NmsWs& ws, char* p, int B, int N, int nc, int multi_label, int max_nms) {
  ws.nchunks = (N + CH - 1) / CH;
  ws.cap = nms_cap(N, nc, multi_label, max_nms);
  ws.cap_pad = 2;
  while (ws.cap_pad < ws.cap) ws.cap_pad <<= 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
  ws.chunk_cnt = (int32_t*)take(sizeof(int32_t) * B * ws.nchunks);
  ws.chunk_cnt2 = (int32_t*)take(sizeof(int32_t) * B * ws.nchunks * 2);
  ws.meta = (int32_t*)take(sizeof(int32_t) * B * 4);
  ws.best_sc = (float*)take(sizeof(float) * (size_t)B * N);
  ws.best_cls = (int32_t*)take(sizeof(int32_t) * (size_t)B * N);
  ws.key = (uint64_t*)take(sizeof(uint64_t) * (size_t)B * ws.cap_pad);
  ws.box = (float4*)take(sizeof(float4) * (size_t)B * ws.cap);
  ws.sbox = (float4*)take(sizeof(float4) * (size_t)B * ws.cap);
  ws.sc = (float*)take(sizeof(float) * (size_t)B * ws.cap);
  ws.cls = (int32_t*)take(sizeof(int32_t) * (size_t)B * ws.cap);
  ws.anchor = (int32_t*)take(sizeof(int32_t) * (size_t)B * ws.cap);
  return off;
}

}  // namespace

extern "C" {

int yad_decode(const void* const* lvl_ptr_host, const int64_t* lvl_sb_host, const int64_t* lvl_sc_host, const int64_t* lvl_sa_host,
               const int32_t* lvl_h_host, const int32_t* lvl_w_host, const float* lvl_stride_host, int nl, int batch, int nc,
               int reg_max, const float* proj, float* y, int dtype, void* stream) {
  YAD_CHECK(nl >= 1 && nl <= MAX_LEVELS, "decode: %d levels unsupported (max %d)", nl, MAX_LEVELS);
  YAD_CHECK(reg_max >= 1 && nc >= 1, "decode: bad reg_max %d / nc %d", reg_max, nc);
  DecodeLevels L;
  L.nl = nl;
  L.start[0] = 0;
  for (int l = 0; l < nl; l++) {
    L.ptr[l] = lvl_ptr_host[l]; L.sb[l] = lvl_sb_host[l]; L.sc[l] = lvl_sc_host[l]; L.sa[l] = lvl_sa_host[l];
    L.h[l] = lvl_h_host[l]; L.w[l] = lvl_w_host[l]; L.stride[l] = lvl_stride_host[l];
    L.start[l + 1] = L.start[l] + lvl_h_host[l] * lvl_w_host[l];
  }
  const int N = L.start[nl];
  if (N == 0 || batch == 0) return 0;
  const int no = 4 * reg_max + nc;
  size_t smem = (size_t)(DA * (no + 1) + 4 * DA) * sizeof(float);
  YAD_CHECK(smem <= 48 * 1024, "decode: %d channels per anchor need %zu B of shared memory", no, smem);
  cudaStream_t st = (cudaStream_t)stream;
  {
    static int fast = -1;
    if (fast < 0) { const char* ev = getenv("YAD_DECODE_NHWC"); fast = (ev && ev[0] == '0') ? 0 : 1; }
    bool ok = fast && dtype == YAD_BF16 && reg_max == 16 && (nc % 8) == 0 && nl <= MAX_LEVELS;
    for (int l = 0; l < nl && ok; l++) ok = L.sc[l] == 1 && (L.sa[l] % 8) == 0 && (L.sb[l] % 8) == 0 && ((uintptr_t)L.ptr[l] & 15) == 0;
    const size_t fsmem = (size_t)DB * (no + 8) * 2 + 4 * DB * sizeof(float);
    if (ok && fsmem <= 100 * 1024) {
      static bool attr = false;
      if (!attr) {
        if (cudaFuncSetAttribute(decode_nhwc_kernel<MAX_LEVELS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
          yad_set_error("decode: cannot raise the dynamic shared memory limit");
          return 2;
        }
        attr = true;
      }
      dim3 gridf((N + DB - 1) / DB, batch);
      YAD_LAUNCH(decode_nhwc_kernel<MAX_LEVELS>, gridf, 256, fsmem, st, L, nc, proj, y, N);
      YAD_LAUNCH_CHECK("decode");
      return 0;
    }
  }
  dim3 grid((N + DA - 1) / DA, batch);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(decode_kernel<T>, grid, 128, smem, st, L, nc, reg_max, proj, y, N);)
  YAD_LAUNCH_CHECK("decode");
  return 0;
}

int64_t yad_nms_workspace_bytes(int batch, int n_anchors, int nc, int multi_label, int max_nms) {
  NmsWs ws;
  return (int64_t)carve(ws, nullptr, batch, n_anchors, nc, multi_label, max_nms);
}

int yad_nms(const float* pred, int batch, int nc, int n_anchors, float conf_thres, float iou_thres, const uint8_t* classes_mask,
            int agnostic, int multi_label, int max_det, int max_nms, float max_wh, float* out, int32_t* out_idx, int32_t* out_count,
            void* workspace, void* stream) {
  YAD_CHECK(conf_thres >= 0.f && conf_thres <= 1.f, "Invalid Confidence threshold %f, valid values are between 0.0 and 1.0", conf_thres);
  YAD_CHECK(iou_thres >= 0.f && iou_thres <= 1.f, "Invalid IoU %f, valid values are between 0.0 and 1.0", iou_thres);
  YAD_CHECK(max_det >= 1 && max_det <= 2048, "nms: max_det %d outside [1, 2048]", max_det);
  YAD_CHECK(max_nms >= 1, "nms: max_nms must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (batch == 0) return 0;
  if (n_anchors == 0) {
    cudaMemsetAsync(out_count, 0, sizeof(int32_t) * batch, st);
    return 0;
  }
  multi_label = multi_label && nc > 1;  // utils/ops.py:235
  NmsWs ws;
  carve(ws, (char*)workspace, batch, n_anchors, nc, multi_label, max_nms);
  dim3 gc(ws.nchunks, batch);
  if (multi_label) {
    YAD_LAUNCH(nms_count_kernel<true>, gc, CH, 0, st, pred, nc, n_anchors, conf_thres, classes_mask, ws);
    YAD_LAUNCH(nms_select_kernel<true>, batch, 1024, 0, st, pred, nc, n_anchors, conf_thres, classes_mask, max_nms, ws);
    YAD_LAUNCH(nms_count2_kernel<true>, gc, CH, 0, st, pred, nc, n_anchors, conf_thres, classes_mask, ws);
    YAD_LAUNCH(nms_gather_kernel<true>, gc, CH, 0, st, pred, nc, n_anchors, conf_thres, classes_mask, ws);
  } else {
    YAD_LAUNCH(nms_count_kernel<false>, gc, CH, 0, st, pred, nc, n_anchors, conf_thres, classes_mask, ws);
    YAD_LAUNCH(nms_select_kernel<false>, batch, 1024, 0, st, pred, nc, n_anchors, conf_thres, classes_mask, max_nms, ws);
    YAD_LAUNCH(nms_count2_kernel<false>, gc, CH, 0, st, pred, nc, n_anchors, conf_thres, classes_mask, ws);
    YAD_LAUNCH(nms_gather_kernel<false>, gc, CH, 0, st, pred, nc, n_anchors, conf_thres, classes_mask, ws);
  }
  {
    static bool attr = false;
    if (!attr) {
      if (cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM * 8) != cudaSuccess ||
          cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 24) != cudaSuccess) {
        yad_set_error("nms: cannot raise the dynamic shared memory limit of the sort / greedy kernels");
        return 2;
      }
      attr = true;
    }
    const int np = ws.cap_pad < SORT_SMEM ? ws.cap_pad : SORT_SMEM;
    YAD_LAUNCH(nms_sort_kernel, batch, 1024, (size_t)np * 8, st, ws, agnostic, max_wh);
  }
  size_t smem = (size_t)max_det * (sizeof(float4) + sizeof(float) + sizeof(int));
  YAD_LAUNCH(nms_greedy_kernel, batch, GREEDY_T, smem, st, ws, iou_thres, max_det, out, out_idx, out_count);
  YAD_LAUNCH_CHECK("nms");
  return 0;
}

}  // extern "C"
