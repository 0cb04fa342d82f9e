"""CPU ORACLE (test infrastructure, NOT product code) -- non_max_suppression restated in numpy fp32.

Follows /root/reference/ultralytics/utils/ops.py:163-312 (non_max_suppression), :412-431 (xywh2xyxy) and the
third-party greedy NMS it calls at ops.py:292 -- torchvision.ops.nms (torchvision 0.26.0 in this image, absent from
/root/reference; published algorithm: stable sort by score descending, box j is suppressed by an earlier kept box i
when inter/(area_i+area_j-inter) > iou_thres, all arithmetic fp32, no epsilon).
Pinned against the live reference + torchvision by oracle/gen_golden.py (fixtures tests/golden/nms_*.npz).
"""
import numpy as np

f32 = np.float32


def xywh2xyxy(x):
    """ops.py:412-431 (fp32: xy -/+ wh/2)."""
    x = np.asarray(x, dtype=f32)
    y = np.empty_like(x)
    half = x[..., 2:4] / f32(2)
    y[..., :2] = x[..., :2] - half
    y[..., 2:4] = x[..., :2] + half
    return y


def greedy_nms(boxes, scores, iou_thres):
    """torchvision.ops.nms CPU semantics (call site ops.py:292). boxes (n,4) xyxy fp32 -> kept indices, score-descending."""
    boxes = np.asarray(boxes, dtype=f32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-np.asarray(scores, dtype=f32), kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = f32(iou_thres)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(f32(0), xx2 - xx1)
        h = np.maximum(f32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        max_det=300, nc=0, max_nms=30000, max_wh=7680, return_indices=False):
    """ops.py:163-312 without the wall-clock break (F7), labels, masks and rotated paths.
    prediction: (B, 4+nc, N) fp32, xywh + class scores.  Returns list of (k,6) fp32 [x1,y1,x2,y2,conf,cls];
    with return_indices also the list of (anchor index, class) int64 pairs of the kept rows."""
    pred = np.asarray(prediction, dtype=f32)
    bs = pred.shape[0]
    nc = nc or pred.shape[1] - 4
    conf = f32(conf_thres)
    xc = pred[:, 4:4 + nc].max(1) > conf  # ops.py:230
    multi_label = multi_label and nc > 1
    pred = pred.transpose(0, 2, 1).copy()  # (B,N,4+nc)
    pred[..., :4] = xywh2xyxy(pred[..., :4])
    out, out_idx = [], []
    for xi in range(bs):
        anchor_ids = np.nonzero(xc[xi])[0]
        x = pred[xi][anchor_ids]
        if x.shape[0] == 0:
            out.append(np.zeros((0, 6), f32)); out_idx.append(np.zeros((0, 2), np.int64)); continue
        box, cls = x[:, :4], x[:, 4:4 + nc]
        if multi_label:
            i, j = np.nonzero(cls > conf)  # row-major: anchor then class (ops.py:267)
            det = np.concatenate((box[i], cls[i, j][:, None], j[:, None].astype(f32)), 1)
            aid = anchor_ids[i]
        else:
            j = cls.argmax(1)  # first max on ties (ops.py:270)
            cf = cls[np.arange(len(j)), j]
            det = np.concatenate((box, cf[:, None], j[:, None].astype(f32)), 1)
            m = cf > conf
            det, aid = det[m], anchor_ids[m]
        if classes is not None:
            m = np.isin(det[:, 5], np.asarray(classes, dtype=f32))
            det, aid = det[m], aid[m]
        n = det.shape[0]
        if n == 0:
            out.append(np.zeros((0, 6), f32)); out_idx.append(np.zeros((0, 2), np.int64)); continue
        if n > max_nms:  # ops.py:281-282 (argsort descending; ties unspecified in torch -> fixtures are tie-free)
            o = np.argsort(-det[:, 4], kind="stable")[:max_nms]
            det, aid = det[o], aid[o]
        c = det[:, 5:6] * f32(0 if agnostic else max_wh)  # ops.py:285 (fp32 class offset, F7)
        keep = greedy_nms(det[:, :4] + c, det[:, 4], iou_thres)[:max_det]
        out.append(det[keep])
        out_idx.append(np.stack((aid[keep], det[keep, 5].astype(np.int64)), 1))
    return (out, out_idx) if return_indices else out
