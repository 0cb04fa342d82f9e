"""CPU ORACLE (test infrastructure, NOT product code) -- the validator's per-image matching step restated in numpy:
`box_iou` (utils/metrics.py:52-71) and `BaseValidator.match_predictions` (engine/validator.py:221-261, the non-scipy branch) as called by
`DetectionValidator._process_batch` (models/yolo/detect/val.py:209-227).  Pinned against the live reference by oracle/gen_golden.py
(tests/golden/match_cases.npz).  Only tests/ may import this file.

The reference's three numpy steps (sort matches by IoU descending; `np.unique` on the detection column keeps each detection's FIRST = best label
and re-orders by detection index; `np.unique` on the label column keeps each label's FIRST = lowest-index detection) are restated in closed form:
  best(d)  = argmax over labels l of iou[l, d] * (cls[l] == pred_cls[d])           (independent of the threshold)
  correct[d, t] = iou[best(d), d] >= thr[t]  and  no d' < d with best(d') == best(d) and iou[best(d'), d'] >= thr[t]
Equal IoUs of one detection with two labels are ordered by an unstable argsort in the reference (unspecified); here the lower label index wins."""
import numpy as np

IOUV = np.linspace(0.5, 0.95, 10, dtype=np.float32)  # torch.linspace(0.5, 0.95, 10), models/yolo/detect/val.py:39


def box_iou(box1, box2, eps=1e-7):
    """(M, 4) x (N, 4) xyxy -> (M, N) fp32, operation order of utils/metrics.py:67-71"""
    a1, a2 = box1[:, None, :2].astype(np.float32), box1[:, None, 2:4].astype(np.float32)
    b1, b2 = box2[None, :, :2].astype(np.float32), box2[None, :, 2:4].astype(np.float32)
    wh = np.clip(np.minimum(a2, b2) - np.maximum(a1, b1), 0, None)
    inter = wh[..., 0] * wh[..., 1]
    area1 = (a2 - a1)[..., 0] * (a2 - a1)[..., 1]
    area2 = (b2 - b1)[..., 0] * (b2 - b1)[..., 1]
    return inter / (area1 + area2 - inter + np.float32(eps))


def match_predictions(pred_classes, true_classes, iou, iouv=IOUV):
    """-> correct (N, len(iouv)) bool"""
    n = pred_classes.shape[0]
    correct = np.zeros((n, len(iouv)), bool)
    if n == 0 or true_classes.shape[0] == 0:
        return correct
    iou = iou * (true_classes[:, None] == pred_classes[None, :])
    best = iou.argmax(0)                    # first maximum = lowest label index on ties
    m = iou[best, np.arange(n)]
    for t, thr in enumerate(iouv):
        taken = set()
        for d in range(n):
            if m[d] >= thr and best[d] not in taken:
                taken.add(best[d])
                correct[d, t] = True
    return correct


def process_batch(detections, gt_bboxes, gt_cls):
    """DetectionValidator._process_batch: detections (N, 6) [x1, y1, x2, y2, conf, cls], gt (M, 4) xyxy, (M,) -> (N, 10) bool"""
    return match_predictions(detections[:, 5], gt_cls, box_iou(gt_bboxes, detections[:, :4]))


def prepare_labels(bboxes_xywhn, img_hw, ori_hw, gain, pad):
    """DetectionValidator._prepare_batch (models/yolo/detect/val.py:104-116) for one image: normalised xywh labels -> xyxy pixels of the
    letterboxed image (xywh2xyxy, utils/ops.py:425-431, times imgsz[[1,0,1,0]]) -> native-image pixels (scale_boxes with ratio_pad,
    utils/ops.py:112-123: subtract pad, divide by gain, clip_boxes :327-331 to the native shape).  fp32 throughout."""
    b = np.asarray(bboxes_xywhn, np.float32).reshape(-1, 4)
    half = b[:, 2:] / np.float32(2)
    xyxy = np.concatenate([b[:, :2] - half, b[:, :2] + half], 1)
    xyxy = xyxy * np.array([img_hw[1], img_hw[0], img_hw[1], img_hw[0]], np.float32)
    xyxy[:, 0::2] -= np.float32(pad[0])
    xyxy[:, 1::2] -= np.float32(pad[1])
    return clip_boxes(xyxy / np.float32(gain), ori_hw)


def clip_boxes(xyxy, ori_hw):
    """utils/ops.py:327-331 (torch branch): x to [0, w], y to [0, h]"""
    out = xyxy.copy()
    out[:, 0::2] = np.clip(out[:, 0::2], 0, np.float32(ori_hw[1]))
    out[:, 1::2] = np.clip(out[:, 1::2], 0, np.float32(ori_hw[0]))
    return out


def _interp_right_continuous(x, xp, fp, left):
    """np.interp's rule (numpy/core/src/multiarray/compiled_base.c, arr_interp) for non-decreasing xp: j = last index with xp[j] <= x;
    below xp[0] -> left; at or beyond xp[-1] -> fp[-1]; x == xp[j] -> fp[j]; else fp[j] + slope * (x - xp[j])."""
    j = np.searchsorted(xp, x, side="right") - 1
    out = np.empty(len(x), np.float64)
    for i, (xv, jj) in enumerate(zip(x, j)):
        if jj < 0:
            out[i] = left
        elif jj >= len(xp) - 1 or xp[jj] == xv:
            out[i] = fp[min(jj, len(xp) - 1)]
        else:
            out[i] = (fp[jj + 1] - fp[jj]) / (xp[jj + 1] - xp[jj]) * (xv - xp[jj]) + fp[jj]
    return out


def average_precision(recall, precision):
    """compute_ap (utils/metrics.py:1112-1141), method 'interp': sentinels (0, 1) / (1, 0), precision envelope = running maximum from the right,
    101-point interpolation, trapezoid rule."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    x = np.linspace(0, 1, 101)
    y = _interp_right_continuous(x, mrec, mpre, left=mpre[0])
    return float(np.sum(np.diff(x) * (y[1:] + y[:-1]) / 2.0))


def ap_per_class(tp, conf, pred_cls, target_cls, eps=1e-16):
    """ap_per_class (utils/metrics.py:1144-1231) without the plots.  tp (n, niou) bool, conf (n,) fp32, pred_cls (n,), target_cls (m,).
    -> dict(tp, fp, p, r, f1, ap (nc, niou), unique_classes, p_curve, r_curve, f1_curve (nc, 1000), f1_index).
    Detections are ordered by a STABLE sort on -conf (the reference's np.argsort default leaves equal confidences unspecified)."""
    order = np.argsort(-conf, kind="stable")
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes, nt = np.unique(target_cls, return_counts=True)
    nc, niou = len(classes), tp.shape[1]
    px = np.linspace(0, 1, 1000)
    ap = np.zeros((nc, niou))
    p_curve, r_curve = np.zeros((nc, 1000)), np.zeros((nc, 1000))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        if sel.sum() == 0 or nt[ci] == 0:
            continue
        tpc = tp[sel].cumsum(0)
        fpc = (1 - tp[sel]).cumsum(0)
        recall = tpc / (nt[ci] + eps)
        precision = tpc / (tpc + fpc)
        neg_conf = -conf[sel].astype(np.float64)     # increasing
        r_curve[ci] = _interp_right_continuous(-px, neg_conf, recall[:, 0], left=0.0)
        p_curve[ci] = _interp_right_continuous(-px, neg_conf, precision[:, 0], left=1.0)
        for j in range(niou):
            ap[ci, j] = average_precision(recall[:, j], precision[:, j])
    f1_curve = 2 * p_curve * r_curve / (p_curve + r_curve + eps)
    # smooth(f1_curve.mean(0), 0.1) (utils/metrics.py:1054-1059): box filter of nf = 101 taps over the edge-padded mean curve
    y = f1_curve.mean(0) if nc else np.zeros(1000)
    nf = round(len(y) * 0.1 * 2) // 2 + 1
    yp = np.concatenate((np.full(nf // 2, y[0]), y, np.full(nf // 2, y[-1])))
    i = int(np.convolve(yp, np.ones(nf) / nf, mode="valid").argmax())
    p, r, f1 = p_curve[:, i], r_curve[:, i], f1_curve[:, i]
    tpn = (r * nt).round()
    fpn = (tpn / (p + eps) - tpn).round()
    return dict(tp=tpn, fp=fpn, p=p, r=r, f1=f1, ap=ap, unique_classes=classes.astype(int), p_curve=p_curve, r_curve=r_curve, f1_curve=f1_curve,
                f1_index=i)


def mean_results(res):
    """Metric.mean_results (utils/metrics.py: mp, mr, map50, map): means over the classes that have data"""
    ap = res["ap"]
    if len(ap) == 0:
        return 0.0, 0.0, 0.0, 0.0
    return float(res["p"].mean()), float(res["r"].mean()), float(ap[:, 0].mean()), float(ap.mean())
