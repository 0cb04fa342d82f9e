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
