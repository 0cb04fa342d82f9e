/*
 * yad.h -- C ABI of libyad.so: the B200 (sm_100a) kernels behind the YOLO-AD-Refine detection hot path.
 *
 * The reference (wcq99681-svg/YOLO-AD-Refine, an Ultralytics fork) is pure Python/PyTorch on this path and has no FFI
 * of its own; every entry point below therefore names the reference *operator* it replaces (file:line relative to
 * /root/reference/ultralytics) -- that is what a reference-side binding (ctypes) hooks up.
 *
 * Conventions
 *   - plain pointers + sizes; no torch types.  All pointers are DEVICE pointers unless the name ends in _host.
 *   - the caller owns every buffer (outputs and workspaces are pre-allocated by the caller).
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises internally.
 *   - return 0 on success; non-zero on error, message via yad_last_error().
 *   - activations are NHWC "views": element (n,y,x,ch) lives at ptr[((n*h + y)*w + x)*ld + ch], ld >= c, so a
 *     channel slice of a wider buffer (concat elimination) is just another view.  dtype: 0 = fp32, 1 = bf16
 *     (the arithmetic is always fp32-accumulate).  Channel counts and ld must be multiples of 8.
 */
#ifndef YAD_H_
#define YAD_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define YAD_F32 0
#define YAD_BF16 1

enum { YAD_ACT_NONE = 0, YAD_ACT_SILU = 1, YAD_ACT_RELU = 2, YAD_ACT_SIGMOID = 3, YAD_ACT_GELU = 4, YAD_ACT_HARDSWISH = 5 };
enum { YAD_CONV_NORMAL = 0, YAD_CONV_TRANSPOSED = 1, YAD_CONV_DEFORM = 2 };

typedef struct {
  void* ptr;
  int32_t n, h, w, c;
  int32_t ld; /* pixel stride in elements */
} yad_tensor;

/* y = act(acc * img_scale[n] * pix_scale[pixel] + bias[co]) * alpha ; y *= mul[pixel][co] ; y += add[pixel][co] */
typedef struct {
  const float* bias;      /* [cout] fp32 or NULL */
  const float* img_scale; /* [n] fp32 or NULL */
  const void* pix_scale;  /* per-pixel scalar (activation dtype), element at pixel*pix_scale_ld, or NULL */
  int32_t pix_scale_ld;
  int32_t act;
  float alpha;
  const void* mul; /* NHWC view with the output's n,h,w,c; or NULL */
  int32_t mul_ld;
  const void* add;
  int32_t add_ld;
} yad_epilogue;

typedef struct {
  int32_t mode;        /* YAD_CONV_* */
  int32_t k;           /* 1 or 3 (square) */
  int32_t stride;      /* 1 or 2.  TRANSPOSED: k=3, stride=2, pad=1, output_padding=1 */
  int32_t pad;
  const void* offmask; /* DEFORM: NHWC view (activation dtype) with >= 27 channels: 18 offsets (dy,dx per tap) + 9 mask logits */
  int32_t offmask_ld;
  int32_t impl;        /* 0 = auto, 1 = SIMT fp32-accumulate kernel, 2 = tcgen05/TMEM kernel (bf16 only) */
} yad_conv_desc;

const char* yad_last_error(void);
int yad_version(void);
/* 1 when the running device is sm_100 (tcgen05 path usable) */
int yad_device_is_sm100(void);

/* -- a1/a7/a8: dense convolution as implicit GEMM, fused epilogue.  Replaces Conv.forward_fuse (nn/modules/conv.py:52-54),
 *    nn.Conv2d / nn.ConvTranspose2d of the neck (z-yaml L12,13,15,20,22), Linear layers of layer 10, the per-image dynamic 1x1
 *    of TaskDecomposition (nn/modules/head.py:651-669) and mmcv ModulatedDeformConv2d (head.py:772-779).
 *    w: [cout][k*k][cin] in the activation dtype (taps row-major), cin = x->c, cout = y->c. */
int yad_conv2d(const yad_tensor* x, const void* w, const yad_conv_desc* d, const yad_epilogue* e, const yad_tensor* y,
               int dtype, void* stream);

/* -- depthwise k x k (3 or 7), stride 1, pad k/2.  w: [k*k][c] fp32, bias fp32 or NULL; then optional per-channel affine
 *    (folded eval BatchNorm), activation, optional residual add.  gate_split > 0: x has 2*gate_split channels and
 *    y[ch] = gelu(dw(x)[ch]) * dw(x)[ch + gate_split]   (EDFFN, nn/modules/block.py:2395-2396).
 *    Replaces ProgressiveFeatureFusion's conv/norm/activation and spatial_mix (block.py:2605-2619). */
int yad_dwconv(const yad_tensor* x, const float* w, const float* bias, const float* scale, const float* shift, int k, int act,
               int gate_split, const void* add, int add_ld, const yad_tensor* y, int dtype, void* stream);

/* -- GroupNorm over NHWC (nn/modules/head.py:1276 Conv_GN.gn, :774 DyDCNv2.norm). stats: double [n][groups][2] (sum, sumsq),
 *    zeroed by yad_gn_stats itself.  yad_gn_apply: y = act((x-mean)*rstd*gamma+beta) (+ add). */
int yad_gn_stats(const yad_tensor* x, int groups, double* stats, int dtype, void* stream);
int yad_gn_apply(const yad_tensor* x, const double* stats, int groups, const float* gamma, const float* beta, float eps, int act,
                 const void* add, int add_ld, const yad_tensor* y, int dtype, void* stream);

/* -- pooling helpers */
/* SPPF (block.py:177-196): the three chained 5x5/s1/p2 max-pools = 5x5, 9x9, 13x13 windows of x; y1,y2,y3 views. */
int yad_sppf_pool(const yad_tensor* x, const yad_tensor* y1, const yad_tensor* y2, const yad_tensor* y3, int dtype, void* stream);
/* global average pool -> out fp32 [n][c] */
int yad_gap(const yad_tensor* x, float* out, int dtype, void* stream);
/* row / column means -> rowmean fp32 [n][h][c], colmean fp32 [n][w][c]  (ELA_HSFPN block.py:1421-1422, CoordAtt head.py:691-692) */
int yad_rowcol_mean(const yad_tensor* x, float* rowmean, float* colmean, int dtype, void* stream);
/* out = x * gh[n][y][ch] * gw[n][x][ch] (x may be NULL: pure outer product, ELA flag=False) */
int yad_rowcol_gate(const yad_tensor* x, const float* gh, const float* gw, const yad_tensor* y, int dtype, void* stream);
/* adaptive_avg_pool to (h/s, w/s) followed by bilinear upsample back, align_corners=False (block.py:2451-2457) */
int yad_pool_upsample(const yad_tensor* x, int s, const yad_tensor* y, int dtype, void* stream);

/* -- MLCA (block.py:1540-1584). local: fp32 [n][25][c] workspace, att: fp32 [n][25][c].
 *    yad_mlca_apply: y = x*att_unpooled (+ add) */
int yad_mlca_pool(const yad_tensor* x, float* local, int dtype, void* stream);
int yad_mlca_att(const float* local, const float* w_global, const float* w_local, int ksize, float local_weight, int n, int c,
                 float* att, void* stream);
int yad_mlca_apply(const yad_tensor* x, const float* att, const void* add, int add_ld, const yad_tensor* y, int dtype, void* stream);

/* -- ELA_HSFPN gate (block.py:1411-1423): v fp32 [n][L][c] (row or column means) -> sigmoid(GN16(conv1d_k7(v)+b)) fp32 [n][L][c].
 *    w: fp32 [cout][7][cin]. */
int yad_ela_gate(const float* v, const float* w, const float* bias, const float* gamma, const float* beta, int n, int L, int c,
                 int groups, float eps, float* out, void* stream);

/* -- CoordAtt gates (head.py:687-705): pooled fp32 [n][L][c] -> sigmoid(conv_x(hardswish(bn(conv1(pooled))))) fp32 [n][L][c]
 *    w1 [mip][c], b1/bn_scale/bn_shift [mip], w2 [c][mip], b2 [c]. */
int yad_coordatt_gate(const float* pooled, const float* w1, const float* b1, const float* bn_scale, const float* bn_shift,
                      const float* w2, const float* b2, int n, int L, int c, int mip, float* out, void* stream);

/* -- tiny per-image MLPs on a GAP vector g fp32 [n][c]:
 *    kind 0: sigmoid(w2 . relu(w1 g + b1) + b2)        -> out [n]      (TaskDecomposition la_conv1/2, head.py:655-656)
 *    kind 1: softmax(w2 . relu(w1 g + b1) + b2) over 3 -> out [n][3]   (AdaptiveDynamicTanh.importance_gate, block.py:2520-2530) */
int yad_gate_mlp(const float* g, const float* w1, const float* b1, const float* w2, const float* b2, int n, int c, int hidden,
                 int nout, int kind, float* out, void* stream);
/* AdaptiveDynamicTanh apply (block.py:2541-2577): y = (sum_i tanh(alpha_i x) imp[n][i]) * weight[ch] + bias[ch] */
int yad_adt_apply(const yad_tensor* x, const float* imp, const float* alphas, const float* weight, const float* bias,
                  const yad_tensor* y, int dtype, void* stream);

/* -- elementwise: op 0: y = alpha*a + beta*b ; op 1: y = a*b ; op 2: y = a * s[pixel] (s: per-pixel scalar view, ld = b_ld);
 *    op 3: y = alpha*a + beta*b + gamma*c3 + d4   (ProgressiveFeatureFusion tail, block.py:2628-2630) */
int yad_eltwise(int op, const yad_tensor* a, const void* b, int b_ld, const void* c3, int c3_ld, const void* d4, int d4_ld,
                float alpha, float beta, float gamma, const yad_tensor* y, int dtype, void* stream);

/* -- CrossScaleAttentionTSSA token statistics (block.py:2459-2474) for one scale: qkv view (n, 1, T, 3c) -> out view (n,1,T,c)
 *    temps: fp32 [heads]. */
int yad_tssa(const yad_tensor* qkv, const float* temps, int heads, const yad_tensor* out, int dtype, void* stream);
/* -- multi-head self attention core (nn.MultiheadAttention, block.py:2432-2434,2479-2486): qkv view (n,1,T,3c) -> out (n,1,T,c) */
int yad_mha(const yad_tensor* qkv, int heads, const yad_tensor* out, int dtype, void* stream);
/* -- mean over `s` token groups: x (n,1,s*T,c) -> y (n,1,T,c) */
int yad_group_mean(const yad_tensor* x, int s, const yad_tensor* y, int dtype, void* stream);
/* -- EDFFN 8x8 patch spectral filter (block.py:2398-2413) as a per-channel 64x64 real matrix m: fp32 [64 out][64 in][c];
 *    y = add + alpha * filter(reflect_pad(x)) cropped. */
int yad_patch_filter(const yad_tensor* x, const float* m, float alpha, const void* add, int add_ld, const yad_tensor* y, int dtype,
                     void* stream);

/* -- layout conversion: NCHW fp32 image (n,c,h,w) -> NHWC view (channels >= c zero-filled) */
int yad_nchw_to_nhwc(const float* src, int c_src, const yad_tensor* y, int dtype, void* stream);

/* -- a9: fused DFL softmax-expectation + make_anchors + dist2bbox(xywh) + x stride + class sigmoid
 *    (head.py:1181-1204,1236-1252; block.py:78-81; utils/tal.py:303-327).
 *    levels: nl raw head outputs; element (b, ch, anchor a of level l) at lvl_ptr[l][b*lvl_sb[l] + ch*lvl_sc[l] + a*lvl_sa[l]]
 *    (NHWC view: sc=1, sa=ld, sb=h*w*ld; (B,no,N) layout: sc=N, sa=1).  proj: fp32[reg_max] DFL conv weight.
 *    y: fp32 (B, 4+nc, N).  */
int yad_decode(const void* const* lvl_ptr_host, const int64_t* lvl_sb_host, const int64_t* lvl_sc_host, const int64_t* lvl_sa_host,
               const int32_t* lvl_h_host, const int32_t* lvl_w_host, const float* lvl_stride_host, int nl, int batch, int nc,
               int reg_max, const float* proj, float* y, int dtype, void* stream);

/* -- a10: batched NMS, replaces utils/ops.py:163-312 non_max_suppression (+ torchvision.ops.nms at :292).
 *    pred: fp32 (B, 4+nc, N) xywh + scores (NOT modified).  classes_mask: uint8[nc] or NULL.
 *    Outputs: out fp32 [B][max_det][6] (x1,y1,x2,y2,conf,cls), out_idx int32 [B][max_det][2] (anchor, class), out_count int32 [B].
 *    workspace: yad_nms_workspace_bytes(B, N, nc) bytes.  status int32[B]: 0 ok, 1 = candidate capacity exceeded. */
int64_t yad_nms_workspace_bytes(int batch, int n_anchors, int nc);
int yad_nms(const float* pred, int batch, int nc, int n_anchors, float conf_thres, float iou_thres, const uint8_t* classes_mask,
            int agnostic, int multi_label, int max_det, int max_nms, float max_wh, float* out, int32_t* out_idx, int32_t* out_count,
            int32_t* status, void* workspace, void* stream);

/* -- a12: TaskAlignedAssigner.forward (utils/tal.py:38-88), one CTA per image.
 *    pd_scores fp32 (B,N,nc) sigmoid scores, pd_bboxes fp32 (B,N,4) xyxy px, anc fp32 (N,2) px, gt_labels fp32 (B,M), gt_bboxes
 *    fp32 (B,M,4), mask_gt fp32 (B,M).  Outputs: target_labels int64 (B,N), target_bboxes fp32 (B,N,4), target_scores fp32 (B,N,nc),
 *    fg_mask uint8 (B,N), target_gt_idx int64 (B,N).  workspace: yad_tal_workspace_bytes(B,N,M). */
int64_t yad_tal_workspace_bytes(int batch, int n_anchors, int n_max_boxes);
int yad_tal_assign(const float* pd_scores, const float* pd_bboxes, const float* anc, const float* gt_labels, const float* gt_bboxes,
                   const float* mask_gt, int batch, int n_anchors, int nc, int n_max_boxes, int topk, float alpha, float beta, float eps,
                   int64_t* target_labels, float* target_bboxes, float* target_scores, uint8_t* fg_mask, int64_t* target_gt_idx,
                   void* workspace, void* stream);

/* -- a11/a13/a14: v8DetectionLoss pieces (utils/loss.py:410-417, 264-311, 18-42, 426-520).
 *    yad_loss_decode: pred_distri fp32 (B,N,4*reg_max) -> pred_bboxes (B,N,4) xyxy grid units and *stride (px) copies;
 *                     pred_scores logits (B,N,nc) -> sigmoid.
 *    yad_loss_bbox : over foreground anchors: CIoU + NWD + DFL terms, accumulates sums[0..3] = {sum (1-ciou) w, sum (1-nwd) w,
 *                     sum dfl w, sum ciou, n_fg} (double[5]) and writes d(loss)/d(pred_distri) into grad_distri given the final
 *                     scale factors (two-phase: call with grad_distri=NULL first to get sums; then with scales).
 *    yad_loss_cls  : SlideLoss-BCE sum over (B,N,nc) -> sums_cls double[1]; grad_scores optional. */
int yad_loss_decode(const float* pred_distri, const float* pred_logits, const float* anc, const float* stride_t, int batch,
                    int n_anchors, int nc, int reg_max, float* pred_bboxes, float* pred_bboxes_px, float* pred_scores_sig, void* stream);
int yad_loss_bbox(const float* pred_distri, const float* pred_bboxes, const float* anc, const float* stride_t,
                  const float* target_bboxes_px, const float* target_scores, const uint8_t* fg_mask, int batch, int n_anchors, int nc,
                  int reg_max, double* sums, float box_scale, float dfl_scale, float* grad_distri, void* stream);
int yad_loss_cls(const float* pred_logits, const float* target_scores, int64_t count, float auto_iou, double* sum_out,
                 float grad_scale, float* grad_logits, void* stream);

/* -- tcgen05 self-test: C[M][N] (fp32) = A[M][K] (bf16, row-major) x B[N][K]^T (bf16) through the UMMA/TMEM path.  Used by the GPU
 *    tests to validate descriptor encodings independently of the convolution loader. */
int yad_tc_gemm_selftest(const void* a, const void* b, float* c, int m, int n, int k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YAD_H_ */
