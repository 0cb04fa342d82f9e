// a11-a14: training-side kernels.
//   TaskAlignedAssigner.forward (utils/tal.py:38-88): tal_topk_kernel, one CTA per (image, gt), + tal_assign_kernel, one CTA per image.
//   v8DetectionLoss pieces (utils/loss.py:410-417 bbox_decode, :264-311 BboxLoss = 0.5(1-CIoU)+0.5(1-NWD) and DFL, :18-42 SlideLoss-BCE)
//   with analytic gradients w.r.t. pred_distri / pred_scores (forward-mode duals through CIoU / NWD; alpha detached as in
//   utils/metrics.py:123-124).
#include <limits.h>

#include "common.cuh"

namespace {

constexpr float kPi = 3.14159265358979323846f;
constexpr int MAX_TOPK = 16;

// ---- scalar abstraction: plain float or float with 4 partial derivatives (w.r.t. the predicted box x1,y1,x2,y2) ----------------
struct Dual {
  float v, d[4];
};
__device__ __forceinline__ Dual mk(float v) { return Dual{v, {0.f, 0.f, 0.f, 0.f}}; }
__device__ __forceinline__ Dual var(float v, int i) { Dual r = mk(v); r.d[i] = 1.f; return r; }
__device__ __forceinline__ float val(float a) { return a; }
__device__ __forceinline__ float val(const Dual& a) { return a.v; }
#define DUAL_BIN(op, expr_v, expr_d)                                   \
  __device__ __forceinline__ Dual op(const Dual& a, const Dual& b) {   \
    Dual r;                                                            \
    r.v = expr_v;                                                      \
    _Pragma("unroll") for (int i = 0; i < 4; i++) r.d[i] = expr_d;     \
    return r;                                                          \
  }
DUAL_BIN(operator+, a.v + b.v, a.d[i] + b.d[i])
DUAL_BIN(operator-, a.v - b.v, a.d[i] - b.d[i])
DUAL_BIN(operator*, a.v* b.v, a.d[i] * b.v + a.v * b.d[i])
DUAL_BIN(operator/, a.v / b.v, (a.d[i] * b.v - a.v * b.d[i]) / (b.v * b.v))
__device__ __forceinline__ Dual operator+(const Dual& a, float b) { Dual r = a; r.v += b; return r; }
__device__ __forceinline__ Dual operator-(const Dual& a, float b) { Dual r = a; r.v -= b; return r; }
__device__ __forceinline__ Dual operator-(float a, const Dual& b) { Dual r; r.v = a - b.v; for (int i = 0; i < 4; i++) r.d[i] = -b.d[i]; return r; }
__device__ __forceinline__ Dual operator*(const Dual& a, float b) { Dual r; r.v = a.v * b; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] * b; return r; }
__device__ __forceinline__ Dual operator/(const Dual& a, float b) { Dual r; r.v = a.v / b; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] / b; return r; }
__device__ __forceinline__ Dual operator/(float a, const Dual& b) { return mk(a) / b; }
__device__ __forceinline__ float s_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float s_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ Dual s_min(const Dual& a, float b) { return a.v <= b ? a : mk(b); }
__device__ __forceinline__ Dual s_max(const Dual& a, float b) { return a.v >= b ? a : mk(b); }
__device__ __forceinline__ float s_clamp0(float a) { return fmaxf(a, 0.f); }
__device__ __forceinline__ Dual s_clamp0(const Dual& a) { return a.v > 0.f ? a : mk(0.f); }
__device__ __forceinline__ float s_atan(float a) { return atanf(a); }
__device__ __forceinline__ Dual s_atan(const Dual& a) { Dual r; r.v = atanf(a.v); float g = 1.f / (1.f + a.v * a.v); for (int i = 0; i < 4; i++) r.d[i] = a.d[i] * g; return r; }
__device__ __forceinline__ float s_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ Dual s_sqrt(const Dual& a) { Dual r; r.v = sqrtf(a.v); float g = 0.5f / r.v; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] * g; return r; }
__device__ __forceinline__ float s_exp(float a) { return expf(a); }
__device__ __forceinline__ Dual s_exp(const Dual& a) { Dual r; r.v = expf(a.v); for (int i = 0; i < 4; i++) r.d[i] = a.d[i] * r.v; return r; }
__device__ __forceinline__ float s_neg(float a) { return -a; }
__device__ __forceinline__ Dual s_neg(const Dual& a) { Dual r; r.v = -a.v; for (int i = 0; i < 4; i++) r.d[i] = -a.d[i]; return r; }

// utils/metrics.py:94-125, xywh=False, CIoU=True.  (a*) = box1 may carry derivatives, (b*) = box2 is constant.
template <typename S>
__device__ __forceinline__ S ciou_box1(S ax1, S ay1, S ax2, S ay2, float bx1, float by1, float bx2, float by2, float* iou_out = nullptr) {
  const float eps = 1e-7f;
  S w1 = ax2 - ax1, h1 = ay2 - ay1 + eps;
  float w2 = bx2 - bx1, h2 = by2 - by1 + eps;
  S inter = s_clamp0(s_min(ax2, bx2) - s_max(ax1, bx1)) * s_clamp0(s_min(ay2, by2) - s_max(ay1, by1));
  S uni = w1 * h1 + w2 * h2 - inter + eps;
  S iou = inter / uni;
  S cw = s_max(ax2, bx2) - s_min(ax1, bx1), ch = s_max(ay2, by2) - s_min(ay1, by1);
  S c2 = cw * cw + ch * ch + eps;
  S dx = (bx1 + bx2) - ax1 - ax2, dy = (by1 + by2) - ay1 - ay2;
  S rho2 = (dx * dx + dy * dy) / 4.0f;
  S da = atanf(w2 / h2) - s_atan(w1 / h1);
  S v = da * da * (4.0f / (kPi * kPi));
  const float alpha = val(v) / (val(v) - val(iou) + (1.0f + eps));  // no_grad
  if (iou_out) *iou_out = val(iou);
  return iou - (rho2 / c2 + v * alpha);
}

// plain-float CIoU with box1 constant (gt) and box2 = prediction, as TaskAlignedAssigner calls it (utils/tal.py:123-125)
__device__ __forceinline__ float ciou_f(float ax1, float ay1, float ax2, float ay2, float bx1, float by1, float bx2, float by2) {
  return ciou_box1<float>(ax1, ay1, ax2, ay2, bx1, by1, bx2, by2);
}

// utils/metrics.py:539-564 wasserstein_loss (returns exp(-sqrt(W2)/C)); pred may carry derivatives
template <typename S>
__device__ __forceinline__ S nwd_sim(S ax1, S ay1, S ax2, S ay2, float bx1, float by1, float bx2, float by2) {
  const float eps = 1e-7f, constant = 12.8f;
  S w1 = ax2 - ax1, h1 = ay2 - ay1 + eps;
  float w2 = bx2 - bx1, h2 = by2 - by1 + eps;
  S cx1 = ax1 + w1 / 2.0f, cy1 = ay1 + h1 / 2.0f;
  float cx2 = bx1 + w2 / 2.0f, cy2 = by1 + h2 / 2.0f;
  S ex = cx1 - cx2, ey = cy1 - cy2;
  S cd = ex * ex + ey * ey + eps;
  S ew = w1 - w2, eh = h1 - h2;
  S whd = (ew * ew + eh * eh) / 4.0f;
  return s_exp(s_neg(s_sqrt(cd + whd)) / constant);
}

// =====================================================================================================================
// loss_decode: (B,N,4*reg_max) logits -> boxes (grid units and pixels); class logits -> sigmoid
// =====================================================================================================================
// 4 lanes per anchor (one per box side): each lane reads its 16 contiguous DFL logits as four 128-bit loads, so a warp sweeps 2 KB of
// contiguous memory; the four expectations meet through shuffles and lane 0 of the quad writes both boxes.
__global__ void loss_decode_kernel(const float* __restrict__ distri, const float* __restrict__ anc, const float* __restrict__ stride_t, int64_t total,
                                   int N, int reg_max, float* __restrict__ boxes, float* __restrict__ boxes_px) {
  pdl_sync();
  const int64_t quads = total * 4;
  for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x; q0 < quads; q0 += (int64_t)gridDim.x * blockDim.x) {  // block-uniform trip count
    const int64_t q = q0 + threadIdx.x;
    const bool ok = q < quads;
    const int64_t i = ok ? q >> 2 : total - 1;
    const int s = (int)(q & 3);
    const float* d = distri + (i * 4 + s) * reg_max;
    float m = -INFINITY;
    for (int k = 0; k < reg_max; k++) m = fmaxf(m, d[k]);
    float sum = 0.f, ex = 0.f;
    for (int k = 0; k < reg_max; k++) { float p = expf(d[k] - m); sum += p; ex = fmaf(p, (float)k, ex); }
    const float e = ex / sum;
    const int lane = threadIdx.x & 31, base = lane & ~3;
    const float e0 = __shfl_sync(0xffffffffu, e, base), e1 = __shfl_sync(0xffffffffu, e, base + 1);
    const float e2 = __shfl_sync(0xffffffffu, e, base + 2), e3 = __shfl_sync(0xffffffffu, e, base + 3);
    if (ok && s == 0) {
      const int a = (int)(i % N);
      const float ax = anc[2 * a], ay = anc[2 * a + 1], st = stride_t[a];
      const float b0 = ax - e0, b1 = ay - e1, b2 = ax + e2, b3 = ay + e3;
      reinterpret_cast<float4*>(boxes)[i] = make_float4(b0, b1, b2, b3);
      reinterpret_cast<float4*>(boxes_px)[i] = make_float4(b0 * st, b1 * st, b2 * st, b3 * st);
    }
  }
}

__global__ void loss_sigmoid_kernel(const float* __restrict__ logits, int64_t count, float* __restrict__ sig) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) sig[i] = sigmoidf_(logits[i]);
}

// =====================================================================================================================
// TaskAlignedAssigner
// =====================================================================================================================
struct TalWs {
  int32_t* tk_cnt;   // [B][M]
  int32_t* tk_idx;   // [B][M][MAX_TOPK]
  int32_t* fg_cnt;   // [B][N]
  int32_t* first_m;  // [B][N]
  int32_t* am;       // [B][N] final gt index
  float* aal;        // [B][N] align metric of the final pair
};

__device__ __forceinline__ bool in_gt(float ax, float ay, const float4& g, float eps) {
  return fminf(fminf(ax - g.x, ay - g.y), fminf(g.z - ax, g.w - ay)) > eps;
}

// align metric and overlap of pair (gt, anchor); zero when the pair is not valid (tal.py:102-121)
__device__ __forceinline__ void pair_metric(const float* __restrict__ pd_scores, const float* __restrict__ pd_bboxes, int64_t row, int nc, int label,
                                            const float4& g, float ax, float ay, float alpha, float beta, float eps, float* align, float* ov) {
  *align = 0.f;
  *ov = 0.f;
  if (!in_gt(ax, ay, g, eps)) return;
  const float4 p = reinterpret_cast<const float4*>(pd_bboxes)[row];
  const float o = fmaxf(ciou_f(g.x, g.y, g.z, g.w, p.x, p.y, p.z, p.w), 0.f);
  const float s = pd_scores[row * nc + label];
  const float sa = alpha == 0.5f ? sqrtf(s) : powf(s, alpha);
  *ov = o;
  *align = sa * powf(o, beta);
}

// grid (M, B): top-k anchors of one gt by align metric; ties -> lower anchor index
__global__ void __launch_bounds__(256) tal_topk_kernel(const float* __restrict__ pd_scores, const float* __restrict__ pd_bboxes,
                                                       const float* __restrict__ anc, const float* __restrict__ gt_labels,
                                                       const float* __restrict__ gt_bboxes, const float* __restrict__ mask_gt, int N, int nc, int M,
                                                       int topk, float alpha, float beta, float eps, TalWs ws) {
  pdl_sync();
  extern __shared__ float metric[];  // [N]
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ int sel[MAX_TOPK];
  const int m = blockIdx.x, b = blockIdx.y;
  const int64_t gi = (int64_t)b * M + m;
  if (mask_gt[gi] == 0.f) {  // padded gt: its top-k indices collapse onto anchor 0 and are dropped (tal.py:147-158)
    if (threadIdx.x == 0) ws.tk_cnt[gi] = 0;
    return;
  }
  const float4 g = reinterpret_cast<const float4*>(gt_bboxes)[gi];
  const int label = max((int)gt_labels[gi], 0);
  for (int a = threadIdx.x; a < N; a += blockDim.x) {
    float al, ov;
    pair_metric(pd_scores, pd_bboxes, (int64_t)b * N + a, nc, label, g, anc[2 * a], anc[2 * a + 1], alpha, beta, eps, &al, &ov);
    metric[a] = al;
  }
  __syncthreads();
  const int k = min(topk, N);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int r = 0; r < k; r++) {
    float bv = -1.f;
    int bi = INT_MAX;
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
      float v = metric[a];
      if (v > bv) { bv = v; bi = a; }  // ascending a within a thread: first max kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { rv[wid] = bv; ri[wid] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; w++)
        if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
      sel[r] = bi;
      metric[bi] = -2.f;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int cnt = 0;
    for (int r = 0; r < k; r++) {
      int a = sel[r];
      if (in_gt(anc[2 * a], anc[2 * a + 1], g, eps)) ws.tk_idx[gi * MAX_TOPK + cnt++] = a;  // mask_topk * mask_in_gts * mask_gt
    }
    ws.tk_cnt[gi] = cnt;
  }
}

// one CTA per image: conflict resolution (tal.py:234-265), targets (162-208) and score normalisation (80-86)
__global__ void __launch_bounds__(512) tal_assign_kernel(const float* __restrict__ pd_scores, const float* __restrict__ pd_bboxes,
                                                         const float* __restrict__ anc, const float* __restrict__ gt_labels,
                                                         const float* __restrict__ gt_bboxes, const float* __restrict__ mask_gt, int N, int nc, int M,
                                                         float alpha, float beta, float eps, TalWs ws, int64_t* __restrict__ target_labels,
                                                         float* __restrict__ target_bboxes, float* __restrict__ target_scores,
                                                         uint8_t* __restrict__ fg_mask, int64_t* __restrict__ target_gt_idx, double* __restrict__ sums) {
  pdl_sync();
  extern __shared__ int pos[];  // pos_align[M], pos_ov[M] as float bits
  __shared__ float red[32];
  int* pos_align = pos;
  int* pos_ov = pos + M;
  const int b = blockIdx.x;
  int32_t* fg_cnt = ws.fg_cnt + (int64_t)b * N;
  int32_t* first_m = ws.first_m + (int64_t)b * N;
  for (int i = threadIdx.x; i < 2 * M; i += blockDim.x) pos[i] = 0;
  for (int i = threadIdx.x; i < M * MAX_TOPK; i += blockDim.x) {
    int m = i / MAX_TOPK, j = i % MAX_TOPK;
    if (j < ws.tk_cnt[(int64_t)b * M + m]) {
      int a = ws.tk_idx[((int64_t)b * M + m) * MAX_TOPK + j];
      atomicAdd(&fg_cnt[a], 1);
      atomicMin(&first_m[a], m);
    }
  }
  __syncthreads();
  const float4* gts = reinterpret_cast<const float4*>(gt_bboxes) + (int64_t)b * M;
  for (int a = threadIdx.x; a < N; a += blockDim.x) {
    const int c = fg_cnt[a];
    if (c == 0) continue;
    const float ax = anc[2 * a], ay = anc[2 * a + 1];
    const int64_t row = (int64_t)b * N + a;
    int m = first_m[a];
    if (c > 1) {  // argmax over ALL gts of the overlaps row (first max), tal.py:256-259
      float best = -1.f;
      for (int q = 0; q < M; q++) {
        float al, ov = 0.f;
        if (mask_gt[(int64_t)b * M + q] != 0.f)
          pair_metric(pd_scores, pd_bboxes, row, nc, max((int)gt_labels[(int64_t)b * M + q], 0), gts[q], ax, ay, alpha, beta, eps, &al, &ov);
        if (ov > best) { best = ov; m = q; }
      }
    }
    float al = 0.f, ov = 0.f;
    if (mask_gt[(int64_t)b * M + m] != 0.f)
      pair_metric(pd_scores, pd_bboxes, row, nc, max((int)gt_labels[(int64_t)b * M + m], 0), gts[m], ax, ay, alpha, beta, eps, &al, &ov);
    ws.am[row] = m;
    ws.aal[row] = al;
    atomicMax(&pos_align[m], __float_as_int(al));
    atomicMax(&pos_ov[m], __float_as_int(ov));
  }
  __syncthreads();
  const int label0 = M > 0 ? max((int)gt_labels[(int64_t)b * M], 0) : 0;
  const float4 box0 = gts[0];
  float tsum = 0.f, nfg = 0.f;
  for (int a = threadIdx.x; a < N; a += blockDim.x) {
    const int64_t row = (int64_t)b * N + a;
    const bool fg = fg_cnt[a] > 0;
    int m = 0, label = label0;
    float4 bx = box0;
    if (fg) {
      m = ws.am[row];
      label = max((int)gt_labels[(int64_t)b * M + m], 0);
      bx = gts[m];
      const float norm = ws.aal[row] * __int_as_float(pos_ov[m]) / (__int_as_float(pos_align[m]) + eps);
      target_scores[row * nc + label] = norm;
      tsum += norm;
      nfg += 1.f;
    }
    target_labels[row] = label;
    target_gt_idx[row] = m;
    fg_mask[row] = fg ? 1 : 0;
    reinterpret_cast<float4*>(target_bboxes)[row] = bx;
  }
  tsum = block_sum(tsum, red);
  nfg = block_sum(nfg, red);
  if (threadIdx.x == 0 && sums) { atomicAdd(&sums[5], (double)tsum); atomicAdd(&sums[4], (double)nfg); }
}

__global__ void tal_empty_kernel(int64_t total, int nc, int64_t* target_labels, int64_t* target_gt_idx, uint8_t* fg_mask) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    target_labels[i] = nc;  // bg_idx (tal.py:62-70)
    target_gt_idx[i] = 0;
    fg_mask[i] = 0;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
size_t tal_carve(TalWs& ws, char* p, int B, int N, int M) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
  ws.fg_cnt = (int32_t*)take(sizeof(int32_t) * (size_t)B * N);   // fg_cnt and first_m first: they are memset together
  ws.first_m = (int32_t*)take(sizeof(int32_t) * (size_t)B * N);
  ws.tk_cnt = (int32_t*)take(sizeof(int32_t) * (size_t)B * (M > 0 ? M : 1));
  ws.tk_idx = (int32_t*)take(sizeof(int32_t) * (size_t)B * (M > 0 ? M : 1) * MAX_TOPK);
  ws.am = (int32_t*)take(sizeof(int32_t) * (size_t)B * N);
  ws.aal = (float*)take(sizeof(float) * (size_t)B * N);
  return off;
}

// =====================================================================================================================
// box / DFL loss over foreground anchors, forward + backward
// =====================================================================================================================
// sums: [0] sum (1-ciou) w, [1] sum (1-nwd) w, [2] sum dfl w, [3] sum ciou, [4] n_fg (written by the assigner), [5] sum target_scores
// (assigner), [6] sum slide-bce.
__global__ void __launch_bounds__(256) loss_bbox_kernel(const float* __restrict__ distri, const float* __restrict__ boxes, const float* __restrict__ anc,
                                                        const float* __restrict__ stride_t, const float* __restrict__ tboxes_px,
                                                        const float* __restrict__ tscores, const uint8_t* __restrict__ fg, int64_t total, int N,
                                                        int nc, int reg_max, double* __restrict__ sums, float box_gain, float dfl_gain,
                                                        float batch_scale, float* __restrict__ grad_distri) {
  pdl_sync();
  __shared__ float red[32];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const float tss = fmaxf((float)sums[5], 1.0f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (!fg[i]) continue;
    const int a = (int)(i % N);
    float w = 0.f;
    for (int c = 0; c < nc; c++) w += tscores[i * nc + c];
    const float st = stride_t[a];
    const float4 pb = reinterpret_cast<const float4*>(boxes)[i];
    float4 tb = reinterpret_cast<const float4*>(tboxes_px)[i];
    tb.x /= st; tb.y /= st; tb.z /= st; tb.w /= st;
    Dual c = ciou_box1<Dual>(var(pb.x, 0), var(pb.y, 1), var(pb.z, 2), var(pb.w, 3), tb.x, tb.y, tb.z, tb.w);
    Dual q = nwd_sim<Dual>(var(pb.x, 0), var(pb.y, 1), var(pb.z, 2), var(pb.w, 3), tb.x, tb.y, tb.z, tb.w);
    s0 += (1.0f - c.v) * w;
    s1 += (1.0f - q.v) * w;
    s3 += c.v;
    // DFL (utils/loss.py:246-261): targets ltrb clamped to [0, reg_max-1-0.01]
    const float ax = anc[2 * a], ay = anc[2 * a + 1];
    const float tgt[4] = {ax - tb.x, ay - tb.y, tb.z - ax, tb.w - ay};
    // d(box loss)/d(box coords) -> d/d(ltrb): x1 = ax - l, y1 = ay - t, x2 = ax + r, y2 = ay + b
    const float gscale = grad_distri ? batch_scale * box_gain * 0.5f * w / tss : 0.f;
    float gl[4];
    gl[0] = (c.d[0] + q.d[0]) * gscale;   // dL/dl = -dL/dx1 and dL/dx1 = -(dc+dq)*gscale
    gl[1] = (c.d[1] + q.d[1]) * gscale;
    gl[2] = -(c.d[2] + q.d[2]) * gscale;
    gl[3] = -(c.d[3] + q.d[3]) * gscale;
    float dfl = 0.f;
    const float* d = distri + i * 4 * reg_max;
    for (int s = 0; s < 4; s++) {
      const float t = fminf(fmaxf(tgt[s], 0.f), (float)(reg_max - 1) - 0.01f);
      const int tl = (int)t;
      const float wl = (float)(tl + 1) - t, wr = 1.0f - wl;
      float m = -INFINITY;
      for (int k = 0; k < reg_max; k++) m = fmaxf(m, d[s * reg_max + k]);
      float sum = 0.f, ex = 0.f;
      for (int k = 0; k < reg_max; k++) { float p = expf(d[s * reg_max + k] - m); sum += p; ex = fmaf(p, (float)k, ex); }
      const float lse = m + logf(sum);
      dfl += (lse - d[s * reg_max + tl]) * wl + (lse - d[s * reg_max + tl + 1]) * wr;
      if (grad_distri) {
        const float e = ex / sum;
        const float gd = batch_scale * dfl_gain * w / tss * 0.25f;
        float* go = grad_distri + i * 4 * reg_max + s * reg_max;
        for (int k = 0; k < reg_max; k++) {
          const float p = expf(d[s * reg_max + k] - m) / sum;
          float gk = gd * (p - (k == tl ? wl : 0.f) - (k == tl + 1 ? wr : 0.f));
          gk += gl[s] * p * ((float)k - e);  // through the softmax expectation
          go[k] = gk;
        }
      }
    }
    s2 += dfl * 0.25f * w;
  }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  s3 = block_sum(s3, red);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[0], (double)s0);
    atomicAdd(&sums[1], (double)s1);
    atomicAdd(&sums[2], (double)s2);
    atomicAdd(&sums[3], (double)s3);
  }
}

// SlideLoss(BCEWithLogits) over (B,N,nc), forward + backward (utils/loss.py:25-42, 510-515)
__global__ void __launch_bounds__(256) loss_cls_kernel(const float* __restrict__ logits, const float* __restrict__ targets, int64_t count,
                                                       double* __restrict__ sums, float cls_gain, float batch_scale, float* __restrict__ grad) {
  pdl_sync();
  __shared__ float red[32];
  float a = sums[4] > 0.0 ? (float)(sums[3] / sums[4]) : -1.0f;
  if (a < 0.2f) a = 0.2f;
  const float a2 = expf(1.0f - a), lo = a - 0.1f;
  const float tss = fmaxf((float)sums[5], 1.0f);
  const float gs = batch_scale * cls_gain / tss;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = logits[i], t = targets[i];
    const float bce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
    const float mod = (t <= lo) ? 1.0f : ((t < a) ? a2 : expf(-(t - 1.0f)));
    acc += bce * mod;
    if (grad) grad[i] = (sigmoidf_(x) - t) * mod * gs;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(&sums[6], (double)acc);
}

// loss_items[3] = (box, cls, dfl) with gains; total[0] = sum * batch  (utils/loss.py:419-424, 517-519)
__global__ void loss_finalize_kernel(const double* __restrict__ sums, float box_gain, float cls_gain, float dfl_gain, float batch, float* __restrict__ out) {
  pdl_sync();
  const double tss = sums[5] > 1.0 ? sums[5] : 1.0;
  // the reference keeps target_scores_sum in fp32
  const float tssf = fmaxf((float)sums[5], 1.0f);
  (void)tss;
  float lb = 0.f, ld = 0.f;
  if (sums[4] > 0.0) {
    lb = (0.5f * (float)(sums[0]) / tssf + 0.5f * (float)(sums[1]) / tssf) * box_gain;
    ld = (float)(sums[2]) / tssf * dfl_gain;
  }
  const float lc = (float)(sums[6]) / tssf * cls_gain;
  out[0] = lb; out[1] = lc; out[2] = ld;
  out[3] = (lb + lc + ld) * batch;
}

int grid_for(int64_t items, int tpb) {
  int64_t g = (items + tpb - 1) / tpb;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace

extern "C" {

int64_t yad_tal_workspace_bytes(int batch, int n_anchors, int n_max_boxes) {
  TalWs ws;
  return (int64_t)tal_carve(ws, nullptr, batch, n_anchors, n_max_boxes);
}

int yad_tal_assign(const float* pd_scores, const float* pd_bboxes, const float* anc, const float* gt_labels, const float* gt_bboxes,
                   const float* mask_gt, int batch, int n_anchors, int nc, int n_max_boxes, int topk, float alpha, float beta, float eps,
                   int64_t* target_labels, float* target_bboxes, float* target_scores, uint8_t* fg_mask, int64_t* target_gt_idx,
                   double* sums, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int B = batch, N = n_anchors, M = n_max_boxes;
  if (B == 0 || N == 0) return 0;
  YAD_CHECK(topk >= 1 && topk <= MAX_TOPK, "tal: topk %d outside [1, %d]", topk, MAX_TOPK);
  const int64_t total = (int64_t)B * N;
  cudaMemsetAsync(target_scores, 0, sizeof(float) * total * nc, st);
  if (M == 0) {  // tal.py:62-70
    cudaMemsetAsync(target_bboxes, 0, sizeof(float) * total * 4, st);
    YAD_LAUNCH(tal_empty_kernel, grid_for(total, 256), 256, 0, st, total, nc, target_labels, target_gt_idx, fg_mask);
    YAD_LAUNCH_CHECK("tal_empty");
    return 0;
  }
  TalWs ws;
  tal_carve(ws, (char*)workspace, B, N, M);
  cudaMemsetAsync(ws.fg_cnt, 0, sizeof(int32_t) * total, st);
  cudaMemsetAsync(ws.first_m, 0x7f, sizeof(int32_t) * total, st);
  size_t smem = sizeof(float) * N;
  YAD_CHECK(smem <= 200 * 1024, "tal: %d anchors do not fit in shared memory", N);
  if (smem > 48 * 1024) cudaFuncSetAttribute(tal_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 g1(M, B);
  YAD_LAUNCH(tal_topk_kernel, g1, 256, smem, st, pd_scores, pd_bboxes, anc, gt_labels, gt_bboxes, mask_gt, N, nc, M, topk, alpha, beta, eps, ws);
  YAD_LAUNCH_CHECK("tal_topk");
  YAD_LAUNCH(tal_assign_kernel, B, 512, sizeof(int) * 2 * M, st, pd_scores, pd_bboxes, anc, gt_labels, gt_bboxes, mask_gt, N, nc, M, alpha, beta, eps, ws,
                                                          target_labels, target_bboxes, target_scores, fg_mask, target_gt_idx, sums);
  YAD_LAUNCH_CHECK("tal_assign");
  return 0;
}

int yad_loss_decode(const float* pred_distri, const float* pred_logits, const float* anc, const float* stride_t, int batch,
                    int n_anchors, int nc, int reg_max, float* pred_bboxes, float* pred_bboxes_px, float* pred_scores_sig, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)batch * n_anchors;
  if (total == 0) return 0;
  YAD_LAUNCH(loss_decode_kernel, grid_for(total * 4, 128), 128, 0, st, pred_distri, anc, stride_t, total, n_anchors, reg_max, pred_bboxes, pred_bboxes_px);
  if (pred_scores_sig) YAD_LAUNCH(loss_sigmoid_kernel, grid_for(total * nc, 256), 256, 0, st, pred_logits, total * nc, pred_scores_sig);
  YAD_LAUNCH_CHECK("loss_decode");
  return 0;
}

int yad_loss_bbox(const float* pred_distri, const float* pred_bboxes, const float* anc, const float* stride_t,
                  const float* target_bboxes_px, const float* target_scores, const uint8_t* fg_mask, int batch, int n_anchors, int nc,
                  int reg_max, double* sums, float box_gain, float dfl_gain, float* grad_distri, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)batch * n_anchors;
  if (total == 0) return 0;
  YAD_CHECK(reg_max >= 2, "loss_bbox: DFL needs reg_max >= 2");
  if (grad_distri) cudaMemsetAsync(grad_distri, 0, sizeof(float) * total * 4 * reg_max, st);
  YAD_LAUNCH(loss_bbox_kernel, grid_for(total, 256), 256, 0, st, pred_distri, pred_bboxes, anc, stride_t, target_bboxes_px, target_scores, fg_mask, total,
                                                         n_anchors, nc, reg_max, sums, box_gain, dfl_gain, (float)batch, grad_distri);
  YAD_LAUNCH_CHECK("loss_bbox");
  return 0;
}

int yad_loss_cls(const float* pred_logits, const float* target_scores, int batch, int n_anchors, int nc, double* sums, float cls_gain,
                 float* grad_logits, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t count = (int64_t)batch * n_anchors * nc;
  if (count == 0) return 0;
  YAD_LAUNCH(loss_cls_kernel, grid_for(count, 256), 256, 0, st, pred_logits, target_scores, count, sums, cls_gain, (float)batch, grad_logits);
  YAD_LAUNCH_CHECK("loss_cls");
  return 0;
}

int yad_loss_finalize(const double* sums, float box_gain, float cls_gain, float dfl_gain, int batch, float* out4, void* stream) {
  YAD_LAUNCH(loss_finalize_kernel, 1, 1, 0, (cudaStream_t)stream, sums, box_gain, cls_gain, dfl_gain, (float)batch, out4);
  YAD_LAUNCH_CHECK("loss_finalize");
  return 0;
}

}  // extern "C"
