// Validator statistics on the device: the per-batch half of DetectionValidator.update_metrics and the per-dataset
// half of get_stats.  Everything the reference does in numpy per image / per class on the host runs here on the batched NMS output.
//   val_labels_kernel : _prepare_batch (models/yolo/detect/val.py:104-116): normalised xywh labels -> xyxy pixels of the native image
//                       (xywh2xyxy utils/ops.py:425-431, * imgsz, scale_boxes with ratio_pad :112-123, clip_boxes :327-331), fp32, reference op order.
//   val_match_kernel  : _process_batch (val.py:209-227) = box_iou (utils/metrics.py:52-71) + BaseValidator.match_predictions
//                       (engine/validator.py:221-261, numpy branch) in closed form, one CTA per image:
//                         best(d) = first argmax over labels of iou * (class match);  correct[d, t] = iou_best(d) >= thr[t] and d is the
//                         lowest detection index with that label and iou_best >= thr[t]   (an atomicMin per (label, threshold) in shared memory).
//   ap_per_class (utils/metrics.py:1144-1231, compute_ap :1112-1141, smooth :1054-1059) in fp64:
//     val_keys_kernel  : sort keys (class << 32 | descending-confidence bits) + the label histogram nt;  cub::DeviceRadixSort (stable) orders them;
//     val_class_kernel : one CTA per class over its segment: cumulative TP counts (packed block scan), precision envelope (reverse max scan),
//                        recall / precision curves at the 1000 confidence abscissae, 101-point interpolated AP per IoU threshold (np.interp's
//                        index rule, np.trapz with numpy's pairwise summation order);
//     val_summary_kernel: F1 curves, their class mean, the 101-tap box filter, its argmax, and p / r / f1 / tp / fp at that index.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

constexpr int VT = 256;           // threads per CTA of the per-image / per-class kernels
constexpr int MAX_NIOU = 12;      // two 64-bit words of six 10-bit counters
constexpr int NPX = 1000;         // confidence abscissae of the curves (utils/metrics.py:1183)
constexpr int NAP = 101;          // recall abscissae of compute_ap (:1132)

__global__ void val_labels_kernel(const float* __restrict__ bb, const int32_t* __restrict__ bidx, int m, float img_w, float img_h,
                                  const yad_image_desc* __restrict__ desc, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float4 b = reinterpret_cast<const float4*>(bb)[i];
  const yad_image_desc d = desc[bidx[i]];
  const float hw = __fdiv_rn(b.z, 2.f), hh = __fdiv_rn(b.w, 2.f);
  float x1 = __fmul_rn(__fsub_rn(b.x, hw), img_w), y1 = __fmul_rn(__fsub_rn(b.y, hh), img_h);
  float x2 = __fmul_rn(__fadd_rn(b.x, hw), img_w), y2 = __fmul_rn(__fadd_rn(b.y, hh), img_h);
  x1 = __fdiv_rn(__fsub_rn(x1, d.pad_x), d.gain);
  y1 = __fdiv_rn(__fsub_rn(y1, d.pad_y), d.gain);
  x2 = __fdiv_rn(__fsub_rn(x2, d.pad_x), d.gain);
  y2 = __fdiv_rn(__fsub_rn(y2, d.pad_y), d.gain);
  const float w0 = (float)d.src_w, h0 = (float)d.src_h;
  reinterpret_cast<float4*>(out)[i] =
      make_float4(fminf(fmaxf(x1, 0.f), w0), fminf(fmaxf(y1, 0.f), h0), fminf(fmaxf(x2, 0.f), w0), fminf(fmaxf(y2, 0.f), h0));
}

// dynamic shared memory: int first[nl_max * niou]; int best[max_det]; float miou[max_det]
__global__ void __launch_bounds__(VT) val_match_kernel(const float* __restrict__ det, int row_ld, const int32_t* __restrict__ count, int max_det,
                                                       const float* __restrict__ gt, const float* __restrict__ gt_cls,
                                                       const int32_t* __restrict__ gt_off, const float* __restrict__ iouv, int niou, int nl_max,
                                                       uint8_t* __restrict__ correct, float* __restrict__ stat_conf,
                                                       float* __restrict__ stat_cls) {
  extern __shared__ int sm_i[];
  int* first = sm_i;
  int* best = sm_i + nl_max * niou;
  float* miou = reinterpret_cast<float*>(best + max_det);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int k = count ? min(count[b], max_det) : max_det;
  const int l0 = gt_off[b], nl = gt_off[b + 1] - l0;
  uint8_t* out = correct + (int64_t)b * max_det * niou;
  if (stat_conf) {  // the rows DetectionValidator.update_metrics appends to its stats (val.py:152-153); padding rows get class -1 = "no detection"
    for (int d = tid; d < max_det; d += VT) {
      const float* r = det + ((int64_t)b * max_det + d) * row_ld;
      stat_conf[(int64_t)b * max_det + d] = d < k ? r[4] : 0.f;
      stat_cls[(int64_t)b * max_det + d] = d < k ? r[5] : -1.f;
    }
    for (int i = k * niou + tid; i < max_det * niou; i += VT) out[i] = 0;
  }
  if (nl <= 0 || k <= 0) {  // update_metrics keeps tp all-false for an image without labels (val.py:131-135, 161)
    for (int i = tid; i < k * niou; i += VT) out[i] = 0;
    return;
  }
  for (int i = tid; i < nl * niou; i += VT) first[i] = INT_MAX;
  __syncthreads();
  const float4* g4 = reinterpret_cast<const float4*>(gt) + l0;
  for (int d = tid; d < k; d += VT) {
    const float* r = det + ((int64_t)b * max_det + d) * row_ld;
    const float x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3], cls = r[5];
    const float area2 = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    float m = 0.f;
    int bl = 0;
    for (int l = 0; l < nl; l++) {
      const float4 g = g4[l];
      const float iw = fmaxf(__fsub_rn(fminf(g.z, x2), fmaxf(g.x, x1)), 0.f), ih = fmaxf(__fsub_rn(fminf(g.w, y2), fmaxf(g.y, y1)), 0.f);
      const float inter = __fmul_rn(iw, ih);
      const float area1 = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
      const float iou = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area1, area2), inter), 1e-7f));
      const float v = gt_cls[l0 + l] == cls ? iou : __fmul_rn(iou, 0.f);
      if (l == 0 || v > m) { m = v; bl = l; }
    }
    best[d] = bl;
    miou[d] = m;
    for (int t = 0; t < niou; t++)
      if (m >= iouv[t]) atomicMin(&first[bl * niou + t], d);
  }
  __syncthreads();
  for (int i = tid; i < k * niou; i += VT) {
    const int d = i / niou, t = i - d * niou;
    out[i] = (miou[d] >= iouv[t] && first[best[d] * niou + t] == d) ? 1 : 0;
  }
}

__device__ __forceinline__ uint32_t descending_bits(float f) {
  uint32_t u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending order of the float as an unsigned integer
  return ~u;
}

__global__ void val_keys_kernel(const float* __restrict__ conf, const float* __restrict__ pred_cls, int64_t n, const float* __restrict__ target_cls,
                                int64_t m, int nc, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int32_t* __restrict__ nt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int c = (int)pred_cls[i];
    c = c < 0 || c >= nc ? nc : c;  // a class outside [0, nc) has no labels: parked behind every real segment
    keys[i] = ((uint64_t)c << 32) | descending_bits(conf[i]);
    vals[i] = (uint32_t)i;
  }
  if (i < m) {
    const int c = (int)target_cls[i];
    if (c >= 0 && c < nc) atomicAdd(&nt[c], 1);
  }
}

__device__ __forceinline__ int64_t lower_bound_key(const uint64_t* keys, int64_t n, uint64_t k) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// np.trapz over the 101 abscissae with numpy's pairwise summation of the 100 products (8 interleaved partial sums, then the tail)
__device__ double trapz101(const double* y) {
  double r[8];
  double res = 0.0;
  for (int i = 0; i < NAP - 1; i++) {
    const double x0 = i == 0 ? 0.0 : __dmul_rn((double)i, 0.01), x1 = i + 1 == NAP - 1 ? 1.0 : __dmul_rn((double)(i + 1), 0.01);
    const double a = __ddiv_rn(__dmul_rn(__dsub_rn(x1, x0), __dadd_rn(y[i + 1], y[i])), 2.0);
    if (i < 8) r[i] = a;
    else if (i < 96) r[i & 7] = __dadd_rn(r[i & 7], a);
    else {
      if (i == 96) res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
      res = __dadd_rn(res, a);
    }
  }
  return res;
}

struct ClassSeg {
  const int32_t* tpc;   // [n_p][niou] cumulative true positives of this class
  const double* env;    // [n_p][niou] precision envelope
  const float* conf;    // [n_p] descending
  int64_t np;
  int niou;
  double denom;         // n_l + eps
};
__device__ __forceinline__ double seg_recall(const ClassSeg& s, int64_t k, int j) { return __ddiv_rn((double)s.tpc[k * s.niou + j], s.denom); }
__device__ __forceinline__ double seg_precision(const ClassSeg& s, int64_t k, int j) { return __ddiv_rn((double)s.tpc[k * s.niou + j], (double)(k + 1)); }
// compute_ap's padded arrays: index 0 = (0, 1), 1..np = detections, np + 1 = (1, 0)
__device__ __forceinline__ double seg_mrec(const ClassSeg& s, int64_t i, int j) { return i == 0 ? 0.0 : i == s.np + 1 ? 1.0 : seg_recall(s, i - 1, j); }
__device__ __forceinline__ double seg_mpre(const ClassSeg& s, int64_t i, int j) { return i == 0 ? 1.0 : i == s.np + 1 ? 0.0 : s.env[(i - 1) * s.niou + j]; }

// np.interp(-px, -conf, f, left) for one abscissa; which = 0: recall[:, 0] (left 0), 1: precision[:, 0] (left 1)
__device__ double curve_at(const ClassSeg& s, double px, int which) {
  int64_t lo = 0, hi = s.np;  // count of confidences >= px (descending order)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((double)s.conf[mid] >= px) lo = mid + 1; else hi = mid;
  }
  const int64_t j = lo - 1;
  if (j < 0) return which ? 1.0 : 0.0;
  const double fj = which ? seg_precision(s, j, 0) : seg_recall(s, j, 0);
  if (j == s.np - 1 || (double)s.conf[j] == px) return fj;
  const double fj1 = which ? seg_precision(s, j + 1, 0) : seg_recall(s, j + 1, 0);
  const double xj = -(double)s.conf[j], xj1 = -(double)s.conf[j + 1];
  const double slope = __ddiv_rn(__dsub_rn(fj1, fj), __dsub_rn(xj1, xj));
  return __dadd_rn(__dmul_rn(slope, __dsub_rn(-px, xj)), fj);
}

__global__ void __launch_bounds__(VT) val_class_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ order, int64_t n,
                                                       const uint8_t* __restrict__ tp, const float* __restrict__ conf, const int32_t* __restrict__ nt,
                                                       int niou, double eps, float* __restrict__ conf_sorted, int32_t* __restrict__ tpc,
                                                       double* __restrict__ env, double* __restrict__ ap, double* __restrict__ p_curve,
                                                       double* __restrict__ r_curve) {
  __shared__ int64_t s_seg[2];
  __shared__ unsigned long long s_warp[2][VT / 32];
  __shared__ double s_wmax[VT / 32];
  __shared__ double s_y[MAX_NIOU][NAP];
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < 2) s_seg[tid] = lower_bound_key(keys, n, (uint64_t)(c + tid) << 32);
  __syncthreads();
  const int64_t lo = s_seg[0], np_ = s_seg[1] - lo;
  const int nl = nt[c];
  if (np_ == 0 || nl == 0) {  // "continue" of utils/metrics.py:1192-1193: the rows stay zero
    for (int i = tid; i < NPX; i += VT) p_curve[(int64_t)c * NPX + i] = r_curve[(int64_t)c * NPX + i] = 0.0;
    if (tid < niou) ap[c * niou + tid] = 0.0;
    return;
  }
  // ---- 1. cumulative true positives per IoU threshold: six 10-bit counters per 64-bit word, one block scan per word and chunk of VT detections
  const int nwords = (niou + 5) / 6;
  int run[MAX_NIOU];
#pragma unroll
  for (int j = 0; j < MAX_NIOU; j++) run[j] = 0;
  for (int64_t base = 0; base < np_; base += VT) {
    const int64_t k = base + tid;
    unsigned long long w[2] = {0ull, 0ull};
    if (k < np_) {
      const uint32_t src = order[lo + k];
      conf_sorted[lo + k] = conf[src];
      const uint8_t* row = tp + (int64_t)src * niou;
      for (int j = 0; j < niou; j++) w[j / 6] |= (unsigned long long)(row[j] != 0) << (10 * (j % 6));
    }
    for (int q = 0; q < nwords; q++) {
      unsigned long long v = w[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (lane == 31) s_warp[q][wid] = v;
      w[q] = v;
    }
    __syncthreads();
    unsigned long long tot[2] = {0ull, 0ull};
    for (int q = 0; q < nwords; q++)
      for (int x = 0; x < VT / 32; x++) {
        const unsigned long long sv = s_warp[q][x];
        if (x < wid) w[q] += sv;
        tot[q] += sv;
      }
    __syncthreads();
    if (k < np_)
      for (int j = 0; j < niou; j++) tpc[(lo + k) * niou + j] = run[j] + (int)((w[j / 6] >> (10 * (j % 6))) & 1023ull);
    for (int j = 0; j < niou; j++) run[j] += (int)((tot[j / 6] >> (10 * (j % 6))) & 1023ull);
  }
  __syncthreads();
  // ---- 2. precision envelope = running maximum from the right (np.flip(np.maximum.accumulate(np.flip(mpre))), utils/metrics.py:1127)
  for (int j = 0; j < niou; j++) {
    double carry = 0.0;  // the trailing sentinel of mpre
    for (int64_t base = 0; base < np_; base += VT) {
      const int64_t q = base + tid, k = np_ - 1 - q;
      double v = q < np_ ? __ddiv_rn((double)tpc[(lo + k) * niou + j], (double)(k + 1)) : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = fmax(v, t);
      }
      if (lane == 31) s_wmax[wid] = v;
      __syncthreads();
      double all = carry;
      for (int x = 0; x < VT / 32; x++) {
        const double sv = s_wmax[x];
        if (x < wid) v = fmax(v, sv);
        all = fmax(all, sv);
      }
      v = fmax(v, carry);
      __syncthreads();
      if (q < np_) env[(lo + k) * niou + j] = v;
      carry = all;
    }
  }
  __syncthreads();  // the block's own global writes are visible to it from here on
  ClassSeg s{tpc + lo * niou, env + lo * niou, conf_sorted + lo, np_, niou, __dadd_rn((double)nl, eps)};
  // ---- 3. recall / precision against confidence (utils/metrics.py:1199-1204)
  const double stepx = __ddiv_rn(1.0, (double)(NPX - 1));
  for (int i = tid; i < NPX; i += VT) {
    const double px = i == NPX - 1 ? 1.0 : __dmul_rn((double)i, stepx);
    r_curve[(int64_t)c * NPX + i] = curve_at(s, px, 0);
    p_curve[(int64_t)c * NPX + i] = curve_at(s, px, 1);
  }
  // ---- 4. AP per threshold: 101-point interpolation of the envelope over recall (compute_ap)
  for (int i = tid; i < niou * NAP; i += VT) {
    const int j = i / NAP, xi = i - j * NAP;
    const double x = xi == NAP - 1 ? 1.0 : __dmul_rn((double)xi, 0.01);
    int64_t a = 0, b = np_ + 2;  // last padded index with mrec <= x (mrec[0] = 0 <= x always)
    while (b - a > 1) {
      const int64_t mid = (a + b) >> 1;
      if (seg_mrec(s, mid, j) <= x) a = mid; else b = mid;
    }
    double y;
    const double xa = seg_mrec(s, a, j), ya = seg_mpre(s, a, j);
    if (a == np_ + 1 || xa == x) y = ya;
    else {
      const double slope = __ddiv_rn(__dsub_rn(seg_mpre(s, a + 1, j), ya), __dsub_rn(seg_mrec(s, a + 1, j), xa));
      y = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xa)), ya);
    }
    s_y[j][xi] = y;
  }
  __syncthreads();
  if (tid < niou) ap[c * niou + tid] = trapz101(s_y[tid]);
}

// one CTA: F1 curves, mean over the classes that have labels, smooth(., 0.1), argmax, per-class p / r / f1 / tp / fp at that index
__global__ void __launch_bounds__(1024) val_summary_kernel(const double* __restrict__ p_curve, const double* __restrict__ r_curve,
                                                           const int32_t* __restrict__ nt, int nc, double eps, double* __restrict__ f1_curve,
                                                           double* __restrict__ summary, int32_t* __restrict__ f1_index) {
  __shared__ double s_mean[NPX];
  __shared__ double s_val[1024];
  __shared__ int s_idx[1024];
  const int tid = threadIdx.x;
  if (tid < NPX) {
    double acc = 0.0;
    int nu = 0;
    for (int c = 0; c < nc; c++) {
      const double p = p_curve[(int64_t)c * NPX + tid], r = r_curve[(int64_t)c * NPX + tid];
      const double f1 = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, p), r), __dadd_rn(__dadd_rn(p, r), eps));
      f1_curve[(int64_t)c * NPX + tid] = f1;
      if (nt[c] > 0) { acc = __dadd_rn(acc, f1); nu++; }
    }
    s_mean[tid] = nu ? __ddiv_rn(acc, (double)nu) : 0.0;
  }
  __syncthreads();
  constexpr int NF = 101;  // round(1000 * 0.1 * 2) // 2 + 1
  double v = -1.0;
  if (tid < NPX) {
    const double wgt = __ddiv_rn(1.0, (double)NF);
    v = 0.0;
    for (int k = 0; k < NF; k++) v = __dadd_rn(v, __dmul_rn(s_mean[min(max(tid + k - NF / 2, 0), NPX - 1)], wgt));
  }
  s_val[tid] = v;
  s_idx[tid] = tid;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (tid < o) {
      const double a = s_val[tid], b = s_val[tid + o];
      if (b > a || (b == a && s_idx[tid + o] < s_idx[tid])) { s_val[tid] = b; s_idx[tid] = s_idx[tid + o]; }
    }
    __syncthreads();
  }
  const int best = s_idx[0];
  if (tid == 0) *f1_index = best;
  for (int c = tid; c < nc; c += 1024) {
    const double p = p_curve[(int64_t)c * NPX + best], r = r_curve[(int64_t)c * NPX + best], f1 = f1_curve[(int64_t)c * NPX + best];
    const double tpn = rint(__dmul_rn(r, (double)nt[c]));
    const double fpn = rint(__dsub_rn(__ddiv_rn(tpn, __dadd_rn(p, eps)), tpn));
    double* o = summary + (int64_t)c * 5;
    o[0] = p; o[1] = r; o[2] = f1; o[3] = tpn; o[4] = fpn;
  }
}

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct ApWorkspace {
  int64_t keys_in, keys_out, vals_in, vals_out, conf_sorted, tpc, env, cub, cub_bytes, total;
};
inline int sort_end_bit(int nc) {
  int bits = 1;
  while ((1 << bits) <= nc) bits++;  // classes 0..nc (nc = the parking class)
  return 32 + bits;
}
int ap_workspace(int64_t n, int nc, int niou, ApWorkspace* w) {
  size_t cub_bytes = 0;
  if (n > 0) {
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                                    (uint32_t*)nullptr, n, 0, sort_end_bit(nc), (cudaStream_t)0);
    if (e != cudaSuccess) return 1;
  }
  int64_t off = 0;
  w->keys_in = off;     off += align256(n * 8);
  w->keys_out = off;    off += align256(n * 8);
  w->vals_in = off;     off += align256(n * 4);
  w->vals_out = off;    off += align256(n * 4);
  w->conf_sorted = off; off += align256(n * 4);
  w->tpc = off;         off += align256(n * niou * 4);
  w->env = off;         off += align256(n * niou * 8);
  w->cub = off;         off += align256((int64_t)cub_bytes);
  w->cub_bytes = (int64_t)cub_bytes;
  w->total = off + 256;
  return 0;
}

}  // namespace

extern "C" {

int yad_val_labels(const float* bboxes_xywhn, const int32_t* batch_idx, int m, int img_h, int img_w, const yad_image_desc* desc, float* out_xyxy,
                   void* stream) {
  YAD_CHECK(m >= 0, "val_labels: negative label count");
  if (m == 0) return 0;
  YAD_CHECK(bboxes_xywhn && batch_idx && desc && out_xyxy, "val_labels: null argument");
  YAD_CHECK((((uintptr_t)bboxes_xywhn | (uintptr_t)out_xyxy) & 15) == 0, "val_labels: boxes must be 16-byte aligned");
  val_labels_kernel<<<cdiv(m, 256), 256, 0, (cudaStream_t)stream>>>(bboxes_xywhn, batch_idx, m, (float)img_w, (float)img_h, desc, out_xyxy);
  YAD_LAUNCH_CHECK("val_labels");
  return 0;
}

int yad_val_match(const float* det, int row_ld, const int32_t* count, int batch, int max_det, const float* gt_xyxy, const float* gt_cls,
                  const int32_t* gt_offset, int max_labels_per_image, const float* iouv, int niou, uint8_t* correct, float* stat_conf,
                  float* stat_cls, void* stream) {
  YAD_CHECK(batch >= 0 && max_det >= 0, "val_match: batch %d / max_det %d out of range", batch, max_det);
  if (batch == 0 || max_det == 0) return 0;
  YAD_CHECK(det && gt_offset && iouv && correct, "val_match: null argument");
  YAD_CHECK((stat_conf == nullptr) == (stat_cls == nullptr), "val_match: stat_conf and stat_cls come together");
  YAD_CHECK(row_ld >= 6, "val_match: rows must hold x1, y1, x2, y2, conf, cls (row_ld %d)", row_ld);
  YAD_CHECK(niou >= 1 && niou <= MAX_NIOU, "val_match: niou %d outside [1, %d]", niou, MAX_NIOU);
  YAD_CHECK(max_labels_per_image >= 0, "val_match: negative max_labels_per_image");
  YAD_CHECK(max_labels_per_image == 0 || (gt_xyxy && gt_cls && ((uintptr_t)gt_xyxy & 15) == 0), "val_match: gt_xyxy must be a 16-byte aligned array");
  const size_t smem = ((size_t)max_labels_per_image * niou + 2 * (size_t)max_det) * 4;
  YAD_CHECK(smem <= 200 * 1024, "val_match: %d labels per image x %d thresholds + %d detections need %zu bytes of shared memory (limit 200 KB)",
            max_labels_per_image, niou, max_det, smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(val_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    YAD_CHECK(e == cudaSuccess, "val_match: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  val_match_kernel<<<batch, VT, smem, (cudaStream_t)stream>>>(det, row_ld, count, max_det, gt_xyxy, gt_cls, gt_offset, iouv, niou,
                                                             max_labels_per_image, correct, stat_conf, stat_cls);
  YAD_LAUNCH_CHECK("val_match");
  return 0;
}

int64_t yad_val_ap_workspace_bytes(int64_t n, int nc, int niou) {
  ApWorkspace w;
  if (n < 0 || nc < 1 || niou < 1 || ap_workspace(n, nc, niou, &w)) return -1;
  return w.total;
}

int yad_val_ap(const uint8_t* tp, const float* conf, const float* pred_cls, int64_t n, const float* target_cls, int64_t m, int nc, int niou,
               double eps, double* ap, double* p_curve, double* r_curve, double* f1_curve, int32_t* nt, double* summary, int32_t* f1_index,
               void* workspace, void* stream) {
  YAD_CHECK(n >= 0 && n < ((int64_t)1 << 31) && m >= 0, "val_ap: n %lld / m %lld out of range", (long long)n, (long long)m);
  YAD_CHECK(nc >= 1 && nc <= 65535, "val_ap: nc %d outside [1, 65535]", nc);
  YAD_CHECK(niou >= 1 && niou <= MAX_NIOU, "val_ap: niou %d outside [1, %d]", niou, MAX_NIOU);
  YAD_CHECK(ap && p_curve && r_curve && f1_curve && nt && summary && f1_index, "val_ap: null output");
  YAD_CHECK(n == 0 || (tp && conf && pred_cls && workspace), "val_ap: null input");
  YAD_CHECK(m == 0 || target_cls, "val_ap: null target_cls");
  cudaStream_t st = (cudaStream_t)stream;
  ApWorkspace w;
  YAD_CHECK(ap_workspace(n, nc, niou, &w) == 0, "val_ap: cub workspace query failed");
  char* base = reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  uint64_t* keys_in = reinterpret_cast<uint64_t*>(base + w.keys_in);
  uint64_t* keys_out = reinterpret_cast<uint64_t*>(base + w.keys_out);
  uint32_t* vals_in = reinterpret_cast<uint32_t*>(base + w.vals_in);
  uint32_t* vals_out = reinterpret_cast<uint32_t*>(base + w.vals_out);
  cudaError_t e = cudaMemsetAsync(nt, 0, sizeof(int32_t) * nc, st);
  YAD_CHECK(e == cudaSuccess, "val_ap: memset: %s", cudaGetErrorString(e));
  const int64_t nm = n > m ? n : m;
  if (nm > 0) {
    val_keys_kernel<<<cdiv(nm, 256), 256, 0, st>>>(conf, pred_cls, n, target_cls, m, nc, keys_in, vals_in, nt);
    YAD_LAUNCH_CHECK("val_keys");
  }
  if (n > 0) {
    size_t cub_bytes = (size_t)w.cub_bytes;
    e = cub::DeviceRadixSort::SortPairs(base + w.cub, cub_bytes, (const uint64_t*)keys_in, keys_out, (const uint32_t*)vals_in, vals_out, n, 0,
                                        sort_end_bit(nc), st);
    YAD_CHECK(e == cudaSuccess, "val_ap: radix sort: %s", cudaGetErrorString(e));
  }
  val_class_kernel<<<nc, VT, 0, st>>>(keys_out, vals_out, n, tp, conf, nt, niou, eps, reinterpret_cast<float*>(base + w.conf_sorted),
                                      reinterpret_cast<int32_t*>(base + w.tpc), reinterpret_cast<double*>(base + w.env), ap, p_curve, r_curve);
  YAD_LAUNCH_CHECK("val_class");
  val_summary_kernel<<<1, 1024, 0, st>>>(p_curve, r_curve, nt, nc, eps, f1_curve, summary, f1_index);
  YAD_LAUNCH_CHECK("val_summary");
  return 0;
}

}  // extern "C"
