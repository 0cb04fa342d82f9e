"""Validator statistics on the GPU: host-side mirror of the reference's validation bookkeeping.

  * `DeviceDetectionStats` follows DetectionValidator (models/yolo/detect/val.py): `init_metrics` (:67-85), `update_metrics(preds, batch)` (:126-176),
    `get_stats()` (:183-191), with the per-image python loop replaced by three launches per BATCH (`yad_val_labels`, `yad_scale_boxes`,
    `yad_val_match`) that append straight into device-resident statistics, and ap_per_class replaced by `yad_val_ap` (one pass per validation run).
    Nothing is read back until `get_stats()`.
  * `process_batch(detections, gt_bboxes, gt_cls)` = DetectionValidator._process_batch (:209-227) for one image, same signature / return value.
  * `ap_per_class(tp, conf, pred_cls, target_cls, ...)` = utils/metrics.py:1144-1231, same signature and 12-tuple (plots are not on this path).

There is no CPU fallback: tensors that arrive on the host are uploaded, the arithmetic runs in libyad.so.
"""
import numpy as np
import torch

from . import ops
from .preprocess import DESC_DTYPE, scale_boxes_params

NPX = 1000  # confidence abscissae of the curves (utils/metrics.py:1183)


def _dev(x, dtype, device):
    t = torch.as_tensor(x)
    return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def _desc_table(ori_shapes, ratio_pads, img_hw, device):
    """yad_image_desc per image carrying only what scale_boxes / clip_boxes need: native (h, w), gain, pad (utils/ops.py:104-112)"""
    desc = np.zeros(len(ori_shapes), DESC_DTYPE)
    for i, shp in enumerate(ori_shapes):
        h0, w0 = int(shp[0]), int(shp[1])
        gain, pad = scale_boxes_params(img_hw, (h0, w0), None if ratio_pads is None else ratio_pads[i])
        desc[i] = (0, h0, w0, 0, 0, 0, 0, 0, gain, pad[0], pad[1])
    return torch.from_numpy(desc.view(np.uint8)).to(device, non_blocking=True)


def process_batch(detections, gt_bboxes, gt_cls, iouv=None):
    """-> correct (N, niou) bool device tensor.  detections (N, 6) [x1, y1, x2, y2, conf, cls], gt_bboxes (M, 4) xyxy, gt_cls (M,)."""
    assert detections.is_cuda, "the validator matching runs on the GPU only (no CPU fallback)"
    dev = detections.device
    iouv = torch.linspace(0.5, 0.95, 10, device=dev) if iouv is None else _dev(iouv, torch.float32, dev)
    n, m = detections.shape[0], gt_cls.shape[0]
    correct = torch.zeros((1, max(n, 1), iouv.numel()), dtype=torch.uint8, device=dev)
    if n and m:
        det = detections.float().contiguous().unsqueeze(0)
        off = torch.tensor([0, m], dtype=torch.int32, device=dev)
        ops.val_match(det, None, _dev(gt_bboxes, torch.float32, dev).reshape(-1, 4), _dev(gt_cls, torch.float32, dev).reshape(-1), off, m, iouv, correct)
    return correct[0, :n].bool()


class ApResult:
    """device outputs of yad_val_ap, indexed by class id"""

    def __init__(self, nc, niou, device):
        f64 = dict(dtype=torch.float64, device=device)
        self.ap = torch.empty((nc, niou), **f64)
        self.p_curve, self.r_curve, self.f1_curve = (torch.empty((nc, NPX), **f64) for _ in range(3))
        self.summary = torch.empty((nc, 5), **f64)
        self.nt = torch.empty(nc, dtype=torch.int32, device=device)
        self.f1_index = torch.empty(1, dtype=torch.int32, device=device)


def ap_per_class_device(tp, conf, pred_cls, target_cls, nc, eps=1e-16):
    """tp uint8 (n, niou), conf / pred_cls fp32 (n), target_cls fp32 (m) on the device -> ApResult (device, nothing read back)"""
    n, niou = tp.shape
    res = ApResult(nc, niou, tp.device)
    ws = torch.empty(ops.val_ap_workspace_bytes(n, nc, niou), dtype=torch.uint8, device=tp.device)
    ops.val_ap(tp, conf, pred_cls, target_cls, nc, eps, res.ap, res.p_curve, res.r_curve, res.f1_curve, res.nt, res.summary, res.f1_index, ws)
    return res


def ap_per_class(tp, conf, pred_cls, target_cls, plot=False, on_plot=None, save_dir=None, names={}, eps=1e-16, prefix=""):
    """Same signature and return tuple as ultralytics.utils.metrics.ap_per_class (numpy in, numpy out); the computation runs on the current GPU."""
    if plot:
        raise NotImplementedError("plots are outside the YOLO-AD-Refine hot path")
    dev = torch.device("cuda", torch.cuda.current_device())
    target_cls = np.asarray(target_cls)
    nc = int(max(np.max(target_cls, initial=-1), np.max(np.asarray(pred_cls), initial=-1))) + 1
    tp_d = _dev(np.ascontiguousarray(tp).astype(np.uint8), torch.uint8, dev).reshape(len(conf), -1)
    res = ap_per_class_device(tp_d, _dev(conf, torch.float32, dev), _dev(pred_cls, torch.float32, dev), _dev(target_cls, torch.float32, dev), max(nc, 1), eps)
    nt = res.nt.cpu().numpy()
    u = np.nonzero(nt)[0]
    s = res.summary.cpu().numpy()[u]
    return (s[:, 3], s[:, 4], s[:, 0], s[:, 1], s[:, 2], res.ap.cpu().numpy()[u], u.astype(int), res.p_curve.cpu().numpy()[u],
            res.r_curve.cpu().numpy()[u], res.f1_curve.cpu().numpy()[u], np.linspace(0, 1, NPX), np.array([]))


class DeviceDetectionStats:
    """The statistics half of DetectionValidator with device-resident state.

    update_metrics(preds, batch):
      preds : (det (B, max_det, 6) fp32, count (B,) int32) device tensors as `postprocess.nms_raw` returns them (boxes in letterboxed-image pixels),
              or the reference's list of (k_i, 6) tensors.
      batch : the collated batch dict of the reference: 'batch_idx' (M,), 'cls' (M, 1) or (M,), 'bboxes' (M, 4) normalised xywh, 'ori_shape' list of
              (h, w), 'ratio_pad' list of ((gain, gain), (left, top)) or absent, 'img' (only its shape is read) or 'imgsz' (h, w).
    """

    keys = ["metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)"]

    def __init__(self, nc, max_det=300, iouv=None, device="cuda", single_cls=False, capacity_images=1024):
        self.nc, self.max_det, self.single_cls = int(nc), int(max_det), bool(single_cls)
        self.device = torch.device(device)
        self.iouv = torch.linspace(0.5, 0.95, 10) if iouv is None else torch.as_tensor(iouv, dtype=torch.float32)  # val.py:39
        self.iouv = self.iouv.to(self.device)
        self.niou = self.iouv.numel()
        self._cap = int(capacity_images)
        self.init_metrics()

    # -- state ------------------------------------------------------------------------------------------------------------------------------
    def init_metrics(self):
        self.seen = 0
        self._tp = torch.empty((self._cap, self.max_det, self.niou), dtype=torch.uint8, device=self.device)
        self._conf = torch.empty((self._cap, self.max_det), dtype=torch.float32, device=self.device)
        self._cls = torch.empty((self._cap, self.max_det), dtype=torch.float32, device=self.device)
        self._target = []  # device tensors of label classes, one per batch
        self.nt_per_image = np.zeros(self.nc, np.int64)
        self.result = None

    def _grow(self, need):
        if need <= self._cap:
            return
        cap = max(need, 2 * self._cap)
        for name in ("_tp", "_conf", "_cls"):
            old = getattr(self, name)
            new = torch.empty((cap,) + tuple(old.shape[1:]), dtype=old.dtype, device=self.device)
            new[:self.seen].copy_(old[:self.seen])
            setattr(self, name, new)
        self._cap = cap

    @staticmethod
    def _pack(preds, max_det, device):
        """list of (k_i, 6) tensors -> (det (B, max_det, 6), count (B,))"""
        det = torch.zeros((len(preds), max_det, 6), dtype=torch.float32, device=device)
        cnt = torch.tensor([min(len(p), max_det) for p in preds], dtype=torch.int32)
        for i, p in enumerate(preds):
            det[i, :int(cnt[i])].copy_(torch.as_tensor(p)[:max_det, :6])
        return det, cnt.to(device)

    # -- per batch ----------------------------------------------------------------------------------------------------------------------------
    def update_metrics(self, preds, batch):
        if isinstance(preds, (list,)):
            det, count = self._pack(preds, self.max_det, self.device)
        else:
            det, count = preds
            det = det.clone()  # _prepare_pred clones (val.py:120); the caller's NMS output stays in letterbox space
        B, max_det = det.shape[0], det.shape[1]
        assert max_det == self.max_det, f"NMS output holds {max_det} rows per image, the statistics were sized for {self.max_det}"
        img_hw = tuple(batch["img"].shape[2:]) if "img" in batch else tuple(batch["imgsz"])
        desc = _desc_table(batch["ori_shape"], batch.get("ratio_pad"), img_hw, self.device)
        # labels: group by image on the host (the collate function already emits them in image order; a stable sort keeps that order)
        bidx = torch.as_tensor(batch["batch_idx"]).reshape(-1).cpu().numpy().astype(np.int64)
        order = np.argsort(bidx, kind="stable")
        counts = np.bincount(bidx, minlength=B)[:B]
        offset = np.zeros(B + 1, np.int32)
        np.cumsum(counts, out=offset[1:])
        m = int(offset[-1])
        cls_np = torch.as_tensor(batch["cls"]).reshape(-1).cpu().numpy().astype(np.float32)[order]
        for si in range(B):  # stat["target_img"] = cls.unique() (val.py:138) -> nt_per_image (val.py:186)
            self.nt_per_image[np.unique(cls_np[offset[si]:offset[si + 1]]).astype(np.int64)] += 1
        gt_cls = _dev(cls_np, torch.float32, self.device)
        boxes = torch.as_tensor(batch["bboxes"]).reshape(-1, 4)
        boxes_n = _dev(boxes[torch.as_tensor(order, device=boxes.device)], torch.float32, self.device)
        gt = torch.empty((m, 4), dtype=torch.float32, device=self.device)
        ops.val_labels(boxes_n, _dev(bidx[order].astype(np.int32), torch.int32, self.device), img_hw, desc, gt)
        if self.single_cls:
            det[:, :, 5] = 0  # val.py:149-150
        ops.scale_boxes(det, count, desc)  # _prepare_pred (val.py:118-124)
        self._grow(self.seen + B)
        s = slice(self.seen, self.seen + B)
        ops.val_match(det, count, gt, gt_cls, _dev(offset, torch.int32, self.device), int(counts.max(initial=0)), self.iouv, self._tp[s],
                      self._conf[s], self._cls[s])
        self._target.append(gt_cls)
        self.seen += B
        return det, count, gt

    # -- per run --------------------------------------------------------------------------------------------------------------------------------
    def get_stats(self):
        n = self.seen * self.max_det
        target = torch.cat(self._target) if self._target else torch.empty(0, dtype=torch.float32, device=self.device)
        res = ap_per_class_device(self._tp[:self.seen].reshape(n, self.niou), self._conf[:self.seen].reshape(n), self._cls[:self.seen].reshape(n),
                                  target, self.nc)
        nt = res.nt.cpu().numpy()
        self.nt_per_class = nt.astype(np.int64)  # np.bincount(target_cls, minlength=nc) (val.py:185)
        ap = res.ap.cpu().numpy()
        u = np.nonzero(nt)[0]
        summ = res.summary.cpu().numpy()
        if not ap.any():  # `if len(stats) and stats["tp"].any()` (val.py:188): without a single true positive the metrics keep their empty defaults
            u = u[:0]
        self.ap_class_index = u.astype(int)
        self.p, self.r, self.f1, self.all_ap = summ[u, 0], summ[u, 1], summ[u, 2], ap[u]
        self.result = res
        return self.results_dict

    # Metric.mean_results / fitness (utils/metrics.py:1291-1358; this fork weights mAP50 0.9 and mAP50-95 0.1)
    def mean_results(self):
        if len(self.ap_class_index) == 0:
            return [0.0, 0.0, 0.0, 0.0]
        return [float(self.p.mean()), float(self.r.mean()), float(self.all_ap[:, 0].mean()), float(self.all_ap.mean())]

    def class_result(self, i):
        return self.p[i], self.r[i], self.all_ap[i, 0], self.all_ap[i].mean()

    @property
    def fitness(self):
        return float((np.array(self.mean_results()) * [0.0, 0.0, 0.9, 0.1]).sum())

    @property
    def results_dict(self):
        return dict(zip(self.keys + ["fitness"], self.mean_results() + [self.fitness]))
