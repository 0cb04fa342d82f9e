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

/* y = act(acc * img_scale[n] * pix_scale[pixel] + bias[co]) * alpha ; y *= mul[pixel][co] ; y *= gate_h[n][oy][co] * gate_w[n][ox][co] ; y += add[pixel][co] */
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
  /* optional fused GroupNorm statistics of the conv output (after scale / bias / act / alpha, before mul / add): stats double
   * [n][gn_groups][2] += (sum, sum of squares); yad_conv2d zeroes the buffer itself.  Used by Conv_GN (nn/modules/head.py:1265-1279). */
  double* gn_stats;
  int32_t gn_groups; /* 0 with gn_stats set: per-channel statistics over the WHOLE batch, double [cout][2] (train-mode BatchNorm, conv.py:50) */
  /* optional separable gate: the output of ELA_HSFPN with flag=False (nn/modules/block.py:1408-1424: sigmoid(x_h) * sigmoid(x_w) broadcast over the
   * map) followed by the yaml's Multiply, without materialising the (n, h, w, c) gate map.  gate_h: (n, out_h, c) and gate_w: (n, out_w, c) in the
   * activation dtype, rows of gate_ld elements; both or neither; the product gate_h * gate_w is rounded to the activation dtype first (= the map
   * yad_rowcol_gate would have written).  NORMAL-mode 1x1 stride-1 convolutions only, not combined with mul / gn_stats. */
  const void* gate_h;
  const void* gate_w;
  int32_t gate_ld;
  int32_t gate_hm, gate_wm; /* set by yad_conv2d itself (output height / width); callers leave them 0 */
} yad_epilogue;

typedef struct {
  int32_t mode;        /* YAD_CONV_* */
  int32_t kh, kw;      /* kernel size: 1x1, 3x3, or 7x1 (ELA_HSFPN's Conv1d over a (n, L, 1, c) view) */
  int32_t stride;      /* 1 or 2.  TRANSPOSED: 3x3, stride=2, pad=1, output_padding=1 */
  int32_t pad_h, pad_w;
  const void* offmask; /* DEFORM: NHWC view (activation dtype) with >= 27 channels: 18 offsets (dy,dx per tap) + 9 mask logits */
  int32_t offmask_ld;
  int32_t impl;        /* 0 = auto, 1 = SIMT fp32-accumulate kernel, 2 = tcgen05/TMEM kernel (bf16 only; TMA-fed where eligible),
                          3 = tcgen05/TMEM kernel with thread-gathered operands only */
} yad_conv_desc;

const char* yad_last_error(void);
int yad_version(void);
/* Programmatic dependent launch (griddepcontrol) between consecutive kernels of a stream: OFF by default (environment YAD_PDL=1 or yad_set_pdl(1) turns it on);
 * returns the previous setting.  Takes effect for launches (and graph captures) made after the call. */
int yad_set_pdl(int enabled);
/* sizeof of the structs that cross this ABI: 0 yad_tensor, 1 yad_epilogue, 2 yad_conv_desc, 3 yad_image_desc, 4 yad_permute_entry (-1 otherwise).
 * A binding checks its own struct declarations against the loaded library with it (tests/test_library.py does so for the ctypes mirror). */
int yad_struct_size(int which);
/* 1 when the running device is sm_100 (tcgen05 path usable) */
int yad_device_is_sm100(void);

/* -- a1/a7/a8: dense convolution as implicit GEMM, fused epilogue.  Replaces Conv.forward_fuse (nn/modules/conv.py:52-54),
 *    nn.Conv2d / nn.ConvTranspose2d of the neck (z-yaml L12,13,15,20,22), Linear layers of layer 10, the per-image dynamic 1x1
 *    of TaskDecomposition (nn/modules/head.py:651-669) and mmcv ModulatedDeformConv2d (head.py:772-779).
 *    w: [cout][kh*kw][cin] in the activation dtype (taps row-major), cin = x->c, cout = y->c. */
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
/* row / column means -> views rowmean (n, h, 1, c), colmean (n, w, 1, c) in the activation dtype
 * (ELA_HSFPN block.py:1421-1422, CoordAtt head.py:691-692).  The 1-D conv / GroupNorm / 1x1 convs that follow in the reference run
 * through yad_conv2d (kh=7,kw=1 or 1x1) + yad_gn_* on these (n, L, 1, c) views. */
int yad_rowcol_mean(const yad_tensor* x, const yad_tensor* rowmean, const yad_tensor* colmean, int dtype, void* stream);
/* y = x * gh[n][y][ch] * gw[n][x][ch]  (x may be NULL: pure outer product, ELA flag=False); gh (n,h,1,c), gw (n,w,1,c) views */
int yad_rowcol_gate(const yad_tensor* x, const yad_tensor* gh, const yad_tensor* gw, const yad_tensor* y, int dtype, void* stream);
/* CoordAtt between its pooling and its gating (nn/modules/head.py:694-703): rows (n, h, 1, c) / cols (n, w, 1, c) from yad_rowcol_mean ->
 * gh (n, h, 1, co) = sigmoid(conv_h(hardswish(bn1(conv1(rows))))), gw likewise with conv_w.  w1 fp32 [mip][c] / b1 [mip]: conv1 with the
 * eval-mode bn1 folded in; wh, ww fp32 [co][mip], bh, bw [co]; mip = 8, 16 or 32.  One launch instead of four 1x1 yad_conv2d calls. */
int yad_coordatt_mlp(const yad_tensor* rows, const yad_tensor* cols, const float* w1, const float* b1, int mip, const float* wh, const float* bh,
                     const float* ww, const float* bw, const yad_tensor* gh, const yad_tensor* gw, int dtype, void* stream);
/* adaptive_avg_pool to (h/s, w/s) followed by bilinear upsample back, align_corners=False (block.py:2451-2457) */
int yad_pool_upsample(const yad_tensor* x, int s, const yad_tensor* y, int dtype, void* stream);

/* -- MLCA (block.py:1540-1584). local: fp32 [n][ls*ls][c] workspace, att: fp32 [n][ls*ls][c] (ls = local_size = 5), scratch: fp32 [n][c]
 *    (per-image global-branch gates; the reference pools them over the BATCH axis, block.py:1575-1579).
 *    yad_mlca_apply: y = x * adaptive_avg_pool(att -> (h,w)) (+ add) */
int yad_mlca_pool(const yad_tensor* x, float* local, int local_size, int dtype, void* stream);
int yad_mlca_att(const float* local, const float* w_global, const float* w_local, int ksize, float local_weight, int n, int c,
                 int local_size, float* att, float* scratch, void* stream);
int yad_mlca_apply(const yad_tensor* x, const float* att, int local_size, const void* add, int add_ld, const yad_tensor* y, int dtype,
                   void* stream);

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

/* -- CrossScaleAttentionTSSA token statistics (block.py:2459-2474) for one scale: qkv view (n, h, w, 3c) with T = h*w tokens ->
 *    tokens [out_token_offset, out_token_offset + T) of the out view (n, T_out, 1, c)  (torch.stack over scales, block.py:2476-2477).
 *    temps: fp32 [heads]. */
int yad_tssa(const yad_tensor* qkv, const float* temps, int heads, const yad_tensor* out, int out_token_offset, int dtype, void* stream);
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
/* uint8 NCHW image batch -> NHWC view, multiplied by `scale` (1/255): the device half of BasePredictor.preprocess
 * (engine/predictor.py:129-133: H2D of the uint8 batch, .float(), /= 255) */
int yad_u8_to_nhwc(const uint8_t* src, int c_src, const yad_tensor* y, float scale, int dtype, void* stream);

/* -- stem: layer 0 (Conv 3->16 k3 s2 p1, BN folded, activation) computed straight from the NCHW image (uint8 when img_is_u8, else fp32)
 *    into an NHWC view: fuses BasePredictor.preprocess' device half (engine/predictor.py:129-133; fold 1/255 into wgt), the layout change
 *    and Conv.forward_fuse (nn/modules/conv.py:52-54).  wgt: fp32 [16][9*cin] with K index = tap*cin + ci; bias fp32[16] or NULL. */
int yad_stem_conv(const void* img, int img_is_u8, int n, int h, int w, int cin, const float* wgt, const float* bias, int act,
                  const yad_tensor* y, int dtype, void* stream);

/* -- f3, Mona adapter (nn/modules/mona.py:36-64): y = LayerNorm_c(x) * gamma + x * gammax, LayerNorm2d (:5-10) over the
 *    channels of every pixel (biased variance, eps inside the square root), weight / bias / gamma / gammax fp32 [c], c <= 1024.  The rest of Mona
 *    runs on yad_conv2d (project1, projector + identity folded into its weights with GELU, project2 + residual) and yad_dwconv (the three depthwise
 *    convolutions of MonaOp :12-34 merged into one 7x7 kernel). */
int yad_ln_mix(const yad_tensor* x, const float* weight, const float* bias, const float* gamma, const float* gammax, float eps,
               const yad_tensor* y, int dtype, void* stream);

/*    AttentionTSSA.forward (nn/modules/block.py:1646-1683) between its two Linear layers (which run on yad_conv2d): w = the projected tokens
 *    (n, h, w, c) with c = heads * head_dim, temp fp32 [heads].  Token-axis L2 normalisation, softmax ACROSS THE HEADS of each token, the
 *    per-channel attention 1 / (1 + dots) and out = -w * Pi * attn, all statistics in fp32; one CTA per image. */
int yad_attention_tssa(const yad_tensor* w, const float* temp, int heads, const yad_tensor* out, int dtype, void* stream);

/* -- a9: fused DFL softmax-expectation + make_anchors + dist2bbox(xywh) + x stride + class sigmoid
 *    (head.py:1181-1204,1236-1252; block.py:78-81; utils/tal.py:303-327).
 *    levels: nl raw head outputs; element (b, ch, anchor a of level l) at lvl_ptr[l][b*lvl_sb[l] + ch*lvl_sc[l] + a*lvl_sa[l]]
 *    (NHWC view: sc=1, sa=ld, sb=h*w*ld; (B,no,N) layout: sc=N, sa=1).  proj: fp32[reg_max] DFL conv weight.
 *    y: fp32 (B, 4+nc, N).  */
int yad_decode(const void* const* lvl_ptr_host, const int64_t* lvl_sb_host, const int64_t* lvl_sc_host, const int64_t* lvl_sa_host,
               const int32_t* lvl_h_host, const int32_t* lvl_w_host, const float* lvl_stride_host, int nl, int batch, int nc,
               int reg_max, const float* proj, float* y, int dtype, void* stream);

/* -- a10: batched NMS, replaces utils/ops.py:163-312 non_max_suppression (+ torchvision.ops.nms at :292).
 *    pred: fp32 (B, 4+nc, N) xywh + scores (NOT modified).  classes_mask: uint8[nc] (1 = keep class) or NULL.
 *    Outputs: out fp32 [B][max_det][6] (x1,y1,x2,y2,conf,cls), out_idx int32 [B][max_det][2] (anchor, class), out_count int32 [B];
 *    rows >= out_count[b] are not written.  workspace: yad_nms_workspace_bytes(...) bytes.  The wall-clock early exit of the
 *    reference (ops.py:234,308-310) is not reproduced. */
int64_t yad_nms_workspace_bytes(int batch, int n_anchors, int nc, int multi_label, int max_nms);
int yad_nms(const float* pred, int batch, int nc, int n_anchors, float conf_thres, float iou_thres, const uint8_t* classes_mask,
            int agnostic, int multi_label, int max_det, int max_nms, float max_wh, float* out, int32_t* out_idx, int32_t* out_count,
            void* workspace, void* stream);

/* -- f1: image pre-processing and box post-scaling on the device.
 *    One descriptor per image; the HOST fills it with the reference's own scalar geometry (LetterBox.__call__, data/augment.py:1558-1586:
 *    ratio, new_unpad, top / left borders; scale_boxes, utils/ops.py:104-112: gain, pad) -- yolo_ad_refine_b200/preprocess.py does. */
typedef struct {
  const uint8_t* src;                 /* DEVICE pointer: source image, HWC uint8 (BGR, as cv2.imread delivers it), 3 channels */
  int32_t src_h, src_w, src_pitch;    /* pitch = bytes per source row */
  int32_t new_w, new_h;               /* LetterBox new_unpad: size of the resized image inside the padded output */
  int32_t top, left;                  /* LetterBox borders: output row / column of the resized image's first pixel */
  float gain, pad_x, pad_y;           /* scale_boxes: boxes = (boxes - pad) / gain, then clipped to [0, src_w] x [0, src_h] */
} yad_image_desc;

/*    yad_letterbox replaces LetterBox.__call__ (data/augment.py:1475-1600: cv2.resize INTER_LINEAR + cv2.copyMakeBorder value 114) and
 *    BasePredictor.preprocess' host half (engine/predictor.py:127-129: stack, BGR -> RGB when swap_rb, HWC -> CHW).  cv2's 8-bit INTER_LINEAR is
 *    fixed-point arithmetic (11-bit coefficients, two truncating shifts); the kernel reproduces it BIT FOR BIT, so the tensor the model sees equals
 *    the reference's.  desc: DEVICE array [batch].  out: uint8 (batch, 3, out_h, out_w) NCHW, out_w a multiple of 4 -- what yad_stem_conv /
 *    yad_u8_to_nhwc take (the /255 stays folded into the stem weights). */
int yad_letterbox(const yad_image_desc* desc, int batch, uint8_t* out, int out_h, int out_w, int pad_value, int swap_rb, void* stream);

/*    yad_scale_boxes replaces scale_boxes + clip_boxes (utils/ops.py:88-123, 315-334; called per image by DetectionPredictor.construct_result,
 *    models/yolo/detect/predict.py:36-41) for a whole batch of NMS outputs: det fp32 [batch][max_det][row_ld] rows starting with x1,y1,x2,y2,
 *    modified in place for rows < count[b] (count NULL: all max_det rows).  fp32 sub / div / clamp in the reference's order. */
int yad_scale_boxes(float* det, int row_ld, const int32_t* count, int batch, int max_det, const yad_image_desc* desc, void* stream);

/* -- f2: validator statistics on the device.
 *    yad_val_labels replaces DetectionValidator._prepare_batch (models/yolo/detect/val.py:104-116): labels of a whole batch, normalised xywh of the
 *    letterboxed image (m, 4) + batch_idx int32 (m) -> xyxy pixels of the native image (xywh2xyxy utils/ops.py:425-431, x imgsz, scale_boxes with
 *    ratio_pad :112-123 = desc[b].gain / pad_x / pad_y, clip_boxes :327-331 to desc[b].src_w / src_h).  fp32 in the reference's operation order.
 *    yad_val_match replaces DetectionValidator._process_batch (val.py:209-227) = box_iou (utils/metrics.py:52-71) + BaseValidator.match_predictions
 *    (engine/validator.py:221-261, numpy branch) for every image of a batched NMS output (det / count as yad_nms + yad_scale_boxes leave them):
 *    gt_xyxy fp32 (M, 4), gt_cls fp32 (M) grouped by image with gt_offset int32 [batch + 1]; iouv fp32 [niou] (niou <= 12);
 *    correct uint8 [batch][max_det][niou], rows < count[b] written (all zero for an image without labels).  Equal IoUs of one detection with two
 *    labels go to the lower label index (the reference's unstable argsort leaves them unspecified).  max_labels_per_image sizes the shared memory.
 *    stat_conf / stat_cls fp32 [batch][max_det] (both or neither NULL): the conf / pred_cls rows update_metrics appends to its statistics
 *    (val.py:152-153); padding rows >= count[b] get class -1 and an all-zero correct row, which yad_val_ap ignores. */
int yad_val_labels(const float* bboxes_xywhn, const int32_t* batch_idx, int m, int img_h, int img_w, const yad_image_desc* desc, float* out_xyxy,
                   void* stream);
int yad_val_match(const float* det, int row_ld, const int32_t* count, int batch, int max_det, const float* gt_xyxy, const float* gt_cls,
                  const int32_t* gt_offset, int max_labels_per_image, const float* iouv, int niou, uint8_t* correct, float* stat_conf,
                  float* stat_cls, void* stream);
/*    yad_val_ap replaces ap_per_class (utils/metrics.py:1144-1231; compute_ap :1112-1141; smooth :1054-1059) on the concatenated statistics of a
 *    validation run (DetectionValidator.get_stats, val.py:183-191): tp uint8 (n, niou), conf fp32 (n), pred_cls fp32 (n), target_cls fp32 (m).
 *    Outputs are indexed by CLASS ID 0..nc-1 (the reference compacts to the classes that have labels: nt[c] > 0): ap fp64 (nc, niou),
 *    p_curve / r_curve / f1_curve fp64 (nc, 1000), nt int32 (nc) labels per class, summary fp64 (nc, 5) = p, r, f1, tp, fp at the max-F1 index,
 *    f1_index int32.  fp64 arithmetic in numpy's order (np.interp's index rule, np.trapz's pairwise sum); equal confidences keep their input
 *    order (stable sort; np.argsort leaves them unspecified).  The ordering step is cub::DeviceRadixSort on (class, confidence) keys.
 *    workspace: yad_val_ap_workspace_bytes(n, nc, niou) bytes (-1 on error). */
int64_t yad_val_ap_workspace_bytes(int64_t n, int nc, int niou);
int yad_val_ap(const uint8_t* tp, const float* conf, const float* pred_cls, int64_t n, const float* target_cls, int64_t m, int nc, int niou,
               double eps, double* ap, double* p_curve, double* r_curve, double* f1_curve, int32_t* nt, double* summary, int32_t* f1_index,
               void* workspace, void* stream);

/* -- a12: TaskAlignedAssigner.forward (utils/tal.py:38-88): one CTA per (image, gt) for the top-k, one CTA per image for the rest.
 *    pd_scores fp32 (B,N,nc) sigmoid scores, pd_bboxes fp32 (B,N,4) xyxy px, anc fp32 (N,2) px, gt_labels fp32 (B,M), gt_bboxes
 *    fp32 (B,M,4), mask_gt fp32 (B,M).  Outputs: target_labels int64 (B,N), target_bboxes fp32 (B,N,4), target_scores fp32 (B,N,nc),
 *    fg_mask uint8 (B,N), target_gt_idx int64 (B,N).  sums: double[8] (may be NULL), [4] += n_fg, [5] += sum(target_scores);
 *    the caller zeroes it.  Top-k ties (equal align metric, in practice only metric 0) go to the lower anchor index; torch.topk
 *    leaves them unspecified.  workspace: yad_tal_workspace_bytes(B,N,M). */
int64_t yad_tal_workspace_bytes(int batch, int n_anchors, int n_max_boxes);
int yad_tal_assign(const float* pd_scores, const float* pd_bboxes, const float* anc, const float* gt_labels, const float* gt_bboxes,
                   const float* mask_gt, int batch, int n_anchors, int nc, int n_max_boxes, int topk, float alpha, float beta, float eps,
                   int64_t* target_labels, float* target_bboxes, float* target_scores, uint8_t* fg_mask, int64_t* target_gt_idx,
                   double* sums, void* workspace, void* stream);

/* -- a11/a13/a14: v8DetectionLoss pieces (utils/loss.py:410-417, 264-311, 18-42, 426-520).  All scalars stay on the device in
 *    sums double[8]: [0] sum (1-CIoU) w, [1] sum (1-NWD) w, [2] sum DFL w, [3] sum CIoU, [4] n_fg, [5] sum target_scores, [6] sum
 *    SlideLoss-BCE.  Call order: zero sums -> yad_loss_decode -> yad_tal_assign -> yad_loss_bbox -> yad_loss_cls -> yad_loss_finalize.
 *    yad_loss_decode: pred_distri fp32 (B,N,4*reg_max) -> pred_bboxes (B,N,4) xyxy in grid units and a *stride (px) copy;
 *                     pred_logits (B,N,nc) -> sigmoid (pred_scores_sig may be NULL).
 *    yad_loss_bbox : over foreground anchors; grad_distri (may be NULL) receives d(loss.sum()*B)/d(pred_distri).
 *    yad_loss_cls  : SlideLoss-BCE with auto_iou = sums[3]/sums[4]; grad_logits (may be NULL) receives d(loss.sum()*B)/d(pred_logits).
 *    yad_loss_finalize: out4 = {box*gain, cls*gain, dfl*gain, (sum of the three) * B}. */
int yad_loss_decode(const float* pred_distri, const float* pred_logits, const float* anc, const float* stride_t, int batch,
                    int n_anchors, int nc, int reg_max, float* pred_bboxes, float* pred_bboxes_px, float* pred_scores_sig, void* stream);
int yad_loss_bbox(const float* pred_distri, const float* pred_bboxes, const float* anc, const float* stride_t,
                  const float* target_bboxes_px, const float* target_scores, const uint8_t* fg_mask, int batch, int n_anchors, int nc,
                  int reg_max, double* sums, float box_gain, float dfl_gain, float* grad_distri, void* stream);
int yad_loss_cls(const float* pred_logits, const float* target_scores, int batch, int n_anchors, int nc, double* sums, float cls_gain,
                 float* grad_logits, void* stream);
int yad_loss_finalize(const double* sums, float box_gain, float cls_gain, float dfl_gain, int batch, float* out4, void* stream);

/* =====================================================================================================================
 * Training path: backward kernels of every operator above + optimizer.  The reference's
 * backward is PyTorch autograd over the same modules (engine/trainer.py:382-397 loss.backward(); optimizer_step :580-588).
 * Conventions: activation gradients are written (acc = 0) or accumulated (acc = 1: dx += ...) in the activation dtype;
 * parameter gradients are ACCUMULATED into fp32 buffers that the caller zeroes once per step.
 * dgrad of a convolution is yad_conv2d itself on the gradient with permuted weights (stride 1: flipped taps; stride 2:
 * the TRANSPOSED mode; ConvTranspose2d: a stride-2 NORMAL conv) -- see yad_permute_pack.
 * ===================================================================================================================== */

/* y = f(a, b, c3, d4) with coefficients that may live on the device (model parameters): *_dev non-NULL overrides the host value.
 * ops 0-3 as yad_eltwise; 4: a*b (+ c3); 5: alpha*a (+ c3); 6: alpha*a*b (+ c3); 7: gelu(a)*b (EDFFN gate, block.py:2396);
 * 8: a * s[pixel] + c3 (s = b, per-pixel scalar view).  y may alias c3 (in-place accumulation). */
int yad_eltwise_dev(int op, const yad_tensor* a, const void* b, int b_ld, const void* c3, int c3_ld, const void* d4, int d4_ld, float alpha,
                    float beta, float gamma, const float* alpha_dev, const float* beta_dev, const float* gamma_dev, const yad_tensor* y, int dtype,
                    void* stream);

/* conv weight gradient (autograd of F.conv2d w.r.t. weight): dw fp32 [cout][kh*kw][cin] += sum_m dy[m][co] * x[pix(m,tap)][ci].
 * NORMAL mode geometry only (kh, kw, stride, pad of `d`; d->impl 1 = SIMT twin); ConvTranspose2d / deformable weights go through it with
 * swapped operands / the column tensor.  bf16, d->impl 0: tcgen05 / TMEM with TMA-fed MN-major operands (stride 1, cin % 64 == 0), a single-pass
 * mma.sync kernel for <= 32 channels, mma.sync split-K otherwise; d->impl 3 forces the mma.sync kernels.  fp32 atomics into dw. */
int yad_conv_wgrad(const yad_tensor* x, const yad_tensor* dy, const yad_conv_desc* d, float* dw, int dtype, void* stream);
/* depthwise weight gradient: dw fp32 [k*k][c] += sum_p dy[p][c] * x[p + tap][c] */
int yad_dwconv_wgrad(const yad_tensor* x, const yad_tensor* dy, int k, float* dw, int dtype, void* stream);
/* out[c] += sum over pixels of a[p][c] (* b[p][c] when b != NULL): bias and per-channel parameter gradients */
int yad_colsum(const yad_tensor* a, const void* b, int b_ld, float* out, int dtype, void* stream);
/* out[0] += scale * sum(a*b)  (per_image = 0)   or   out[n] += scale * sum over image n of a*b / img_div[n]  (per_image = 1; img_div may be NULL) */
int yad_dot(const yad_tensor* a, const void* b, int b_ld, int per_image, float scale, const float* img_div, float* out, int dtype, void* stream);
/* y[p][0] = sum_c a[p][c]*b[p][c], y[p][1..7] = 0 (y: 8-channel view): gradient of the per-pixel cls_prob gate (head.py:1168-1175) */
int yad_dot_pixel(const yad_tensor* a, const void* b, int b_ld, const yad_tensor* y, int dtype, void* stream);

/* GroupNorm / batch-statistics BatchNorm backward (autograd of F.group_norm / F.batch_norm(training=True) fused with the activation that
 * follows): x = the normalised op's INPUT, stats = (sum, sumsq) as produced by yad_gn_stats, y = act(xhat*gamma+beta).  BatchNorm: pass the
 * batch as ONE image (n=1, h=N*H) with groups = c.  sums: double [n][groups][2] scratch.  dgamma/dbeta fp32 [c] accumulated (may be NULL). */
int yad_norm_bwd(const yad_tensor* x, const yad_tensor* dy, const double* stats, int groups, const float* gamma, const float* beta, float eps,
                 int act, double* sums, float* dgamma, float* dbeta, const yad_tensor* dx, int acc, int dtype, void* stream);
/* nn.BatchNorm2d running statistics in train(): r = (1-m) r + m * batch stat (unbiased variance); stats double [c][2], count = N*H*W */
int yad_bn_running_update(const double* stats, int c, double count, float momentum, float* running_mean, float* running_var, void* stream);
/* dx (+)= dy * f'(y) for activations that are functions of their output (sigmoid, relu) */
int yad_act_bwd(const yad_tensor* y, const yad_tensor* dy, int act, const yad_tensor* dx, int acc, int dtype, void* stream);

/* dx[n,y,x,:] (+)= img[n][:] * s_img + row[n,y,:] * s_row + col[n,x,:] * s_col: backward of yad_gap / yad_rowcol_mean (NULL = absent) */
int yad_bcast_add(const yad_tensor* dx, const float* img, float s_img, const yad_tensor* row, float s_row, const yad_tensor* col, float s_col,
                  int acc, int dtype, void* stream);
/* backward of yad_rowcol_gate: dgh (n,h,1,c) / dgw (n,w,1,c) overwritten; dx (+)= dy*gh*gw when x != NULL */
int yad_rowcol_gate_bwd(const yad_tensor* x, const yad_tensor* gh, const yad_tensor* gw, const yad_tensor* dy, const yad_tensor* dx, int acc,
                        const yad_tensor* dgh, const yad_tensor* dgw, int dtype, void* stream);
/* backward of yad_mlca_pool/att/apply (block.py:1540-1584, incl. the batch-axis pooling of the global branch).  local / att: the forward's
 * fp32 [n][ls*ls][c] buffers; datt, dlocal: fp32 scratch of the same size; dG: fp32 [ls][c] scratch; dw_*: fp32 [ksize] accumulated */
int yad_mlca_bwd(const yad_tensor* x, const yad_tensor* dy, const float* local, const float* att, const float* w_global, const float* w_local,
                 int ksize, float local_weight, int local_size, float* datt, float* dlocal, float* dG, float* dw_global, float* dw_local,
                 const yad_tensor* dx, int acc, int dtype, void* stream);
/* 5x5/s1/p2 max-pool backward (SPPF, block.py:190-196): output gradient = dy_t (view, may be NULL) + dy_f (fp32 dense, may be NULL);
 * dx_f fp32 dense (n,h,w,c) += ... (first maximum in row-major order, torch's index rule) */
int yad_maxpool5_bwd(const yad_tensor* x, const void* dy_t, int dy_ld, const float* dy_f, float* dx_f, int dtype, void* stream);
/* y (+)= scale * src (fp32 dense NHWC with y->c channels) */
int yad_cast_acc(const float* src, float scale, const yad_tensor* y, int acc, int dtype, void* stream);
/* backward of yad_pool_upsample; dpool: fp32 [n][h/s][w/s][c] scratch */
int yad_pool_upsample_bwd(const yad_tensor* dy, int s, float* dpool, const yad_tensor* dx, int acc, int dtype, void* stream);
/* backward of yad_gate_mlp: dout [n][nout] -> dg [n][c] overwritten; dw1, db1, dw2, db2 accumulated */
int yad_gate_mlp_bwd(const float* g, const float* w1, const float* b1, const float* w2, const float* b2, int n, int c, int hidden, int nout,
                     int kind, const float* dout, float* dg, float* dw1, float* db1, float* dw2, float* db2, void* stream);
/* backward of yad_adt_apply: dimp fp32 [n][3] (zeroed here); dalpha[3], dweight[c], dbias[c] accumulated */
int yad_adt_bwd(const yad_tensor* x, const yad_tensor* dy, const float* imp, const float* alphas, const float* weight, const yad_tensor* dx, int acc,
                float* dimp, float* dalpha, float* dweight, float* dbias, int dtype, void* stream);
/* backward of y = gelu(a) * b */
int yad_gelu_gate_bwd(const yad_tensor* a, const void* b, int b_ld, const yad_tensor* dy, const yad_tensor* da, void* db, int db_ld, int acc,
                      int dtype, void* stream);
/* y (+)= a * s[n] (fp32 per-image scale): TaskDecomposition's per-image dynamic weight (head.py:651-669) */
int yad_scale_img(const yad_tensor* a, const float* s, const yad_tensor* y, int acc, int dtype, void* stream);
int yad_group_mean_bwd(const yad_tensor* dy, int s, const yad_tensor* dx, int acc, int dtype, void* stream);
/* backward of yad_patch_filter: dx_f fp32 dense (n,h,w,c) (zeroed here); dm fp32 [64][64][c] accumulated (may be NULL) */
int yad_patch_filter_bwd(const yad_tensor* x, const yad_tensor* dy, const float* m, float alpha, float* dx_f, float* dm, int dtype, void* stream);
/* backward of yad_tssa for one scale; dqkv overwritten (dq = 0: the statistic sum_j normalize(q)_j^2 is identically 1); dtemps accumulated */
int yad_tssa_bwd(const yad_tensor* qkv, const float* temps, int heads, const yad_tensor* dout, int out_token_offset, const yad_tensor* dqkv,
                 float* dtemps, int dtype, void* stream);
/* backward of yad_mha; lse_d: fp32 [n*heads][T][2] scratch */
int yad_mha_bwd(const yad_tensor* qkv, int heads, const yad_tensor* out, const yad_tensor* dout, const yad_tensor* dqkv, float* lse_d, int dtype,
                void* stream);

/* DCNv2 as column tensor (training path): col (n,h,w,9c) = mask * bilinear sample, tap-major; the convolution itself, its dgrad and wgrad
 * are then 1x1 yad_conv2d / yad_conv_wgrad calls on col.  yad_deform_col_bwd: dx_f fp32 dense (n,h,w,c) (zeroed here), doffmask view
 * (>= 27 channels) overwritten with d(offsets), d(mask logits) */
int yad_deform_col(const yad_tensor* x, const yad_tensor* offmask, const yad_tensor* col, int dtype, void* stream);
int yad_deform_col_bwd(const yad_tensor* x, const yad_tensor* offmask, const yad_tensor* dcol, float* dx_f, const yad_tensor* doffmask, int dtype,
                       void* stream);

/* raw head outputs (n,h,w,4*reg_max+nc) of one level <-> the loss layout: distri fp32 (B,N,4*reg_max), logits fp32 (B,N,nc)
 * (utils/loss.py:430-436 cat / permute), and the scatter of the loss gradients back (x scale) */
int yad_head_pack(const yad_tensor* level, int anchor0, int n_anchors, int reg_ch, int nc, float* distri, float* logits, int dtype, void* stream);
int yad_head_unpack(const float* grad_distri, const float* grad_logits, float scale, int anchor0, int n_anchors, int reg_ch, int nc,
                    const yad_tensor* level, int dtype, void* stream);

/* Fusion('bifpn') weights (block.py:1532-1534): w = relu(p)/(sum relu(p) + 1e-4) (w may be NULL); dp += J^T dw when dp != NULL */
int yad_fusion_weights(const float* p, int k, float* w, const float* dw, float* dp, void* stream);
/* tiny parameter-sized matrices: C[m][n] (+)= sum_k A[m][k] B[n][k] (trans_a = 0) or sum_k A[k][m] B[k][n] (trans_a = 1) */
int yad_small_gemm(const float* a, const float* b, float* c, int m, int n, int k, int trans_a, int acc, void* stream);

/* Layout table entry: dst[(i*n1 + t')*p2 + j] = src[src_off + i*s0 + t*s1 + j*s2] for i < n0, j < n2 (zero padding up to p0, p2),
 * t' = flip ? n1-1-t : t.  Covers conv weights (OIHW -> [cout][tap][cin]), their dgrad twins, ConvTranspose2d, depthwise, padded biases. */
typedef struct {
  int64_t src_off, dst_off;
  int64_t n0, n1, n2, p0, p2;
  int64_t s0, s1, s2;
  int32_t flip, dst_f32;
} yad_permute_entry;
/* master fp32 parameters (torch layout) -> kernel layouts; entries with dst_f32 go to dst_f32, the others to dst_t (activation dtype) */
int yad_permute_pack(const yad_permute_entry* table_dev, int n_entries, int64_t max_elems, const float* src, void* dst_t, float* dst_f32, int dtype,
                     void* stream);
/* packed fp32 gradients -> += torch-layout gradient arena */
int yad_permute_unpack(const yad_permute_entry* table_dev, int n_entries, int64_t max_elems, const float* packed, float* grads, void* stream);

/* optimizer (engine/trainer.py:580-588, 784-808): *out += sum x^2;  clip_grad_norm_(max_norm) + SGD(nesterov) with three parameter groups
 * (group[i] in {0: decay, 1: norm weight, 2: bias, 255: frozen}); ModelEMA (utils/torch_utils.py:530-541) */
int yad_sqnorm(const float* x, int64_t n, double* out, void* stream);
int yad_sgd_step(float* params, const float* grads, float* momentum_buf, const uint8_t* group, int64_t n, const float* lr3_host,
                 const float* wd3_host, float momentum, float max_norm, const double* norm_sq, int first_step, void* stream);
/* clip_grad_norm_(max_norm) + torch.optim.AdamW(betas = (beta1, beta2), eps) with the same three groups: the optimizer the reference's
 * `optimizer=auto` picks for runs shorter than 10,000 iterations (engine/trainer.py:773-782, 805).  step counts from 1. */
int yad_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* group, int64_t n, const float* lr3_host,
                   const float* wd3_host, float beta1, float beta2, float eps, int step, float max_norm, const double* norm_sq, void* stream);
int yad_ema_update(float* ema, const float* params, int64_t n, float decay, void* stream);
/* The same optimizer / EMA steps with every per-step scalar read from DEVICE memory, so that ONE captured CUDA graph serves a whole run (the
 * reference changes lr every iteration during warm-up, engine/trainer.py:363-378; Adam's bias correction and ModelEMA's decay ramp,
 * utils/torch_utils.py:531-533, change every step).  hyper_dev: float[13] = lr[3] | weight_decay[3] | momentum or beta1 | beta2 | eps |
 * 1 - beta1^t | sqrt(1 - beta2^t) | max_norm | ema decay.  The momentum arena must start at zero (then the first SGD step equals torch's). */
int yad_sgd_step_dev(float* params, const float* grads, float* momentum_buf, const uint8_t* group, int64_t n, const float* hyper_dev,
                     const double* norm_sq, void* stream);
int yad_adamw_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* group, int64_t n,
                       const float* hyper_dev, const double* norm_sq, void* stream);
int yad_ema_update_dev(float* ema, const float* params, int64_t n, const float* decay_dev, void* stream);


/* -- row f4: training-time augmentations on HWC uint8 BGR device images (the reference runs them on the host: data/augment.py).
 *    yad_hsv_lut : RandomHSV.__call__ (:1345-1378) in place: BGR2HSV -> per-image LUTs (uint8 [n][3][256]: hue, saturation, value; built on the
 *                  host exactly as the reference builds them) -> HSV2BGR, OpenCV 4.x's 8-bit arithmetic bit for bit (vectorised path).
 *    yad_flip    : RandomFlip.__call__ (:1429-1472), out of place; flags uint8 [n]: bit 0 = np.flipud, bit 1 = np.fliplr.
 *    yad_mosaic4 : Mosaic._mosaic4 (:657-713): n_out canvases (2s x 2s, filled with 114), four paste rectangles each. */
typedef struct {
  const void* src;          /* HWC uint8 image, row pitch src_w pixels */
  int32_t src_w;
  int32_t x1a, y1a, x2a, y2a; /* destination rectangle in the canvas (exclusive max) */
  int32_t x1b, y1b;           /* top-left corner of the source rectangle */
  int32_t pad_;
} yad_mosaic_desc;
int yad_hsv_lut(void* img_u8_hwc_bgr, int n, int h, int w, const void* luts_dev, void* stream);
int yad_flip(const void* src, void* dst, int n, int h, int w, const void* flags_dev, void* stream);
int yad_mosaic4(void* canvas, int s2, int n_out, const yad_mosaic_desc* desc_dev, void* stream);

/* -- tcgen05 self-test: C[M][N] (fp32) = A[M][K] (bf16, row-major) x B[N][K]^T (bf16) through the UMMA/TMEM path.  Used by the GPU
 *    tests to validate descriptor encodings independently of the convolution loader. */
int yad_tc_gemm_selftest(const void* a, const void* b, float* c, int m, int n, int k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YAD_H_ */
