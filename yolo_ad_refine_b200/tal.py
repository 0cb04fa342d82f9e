"""Drop-in for ultralytics.utils.tal.TaskAlignedAssigner (utils/tal.py:13-265) and make_anchors (:303-315), backed by libyad.so."""
import ctypes as C

import torch

from . import ops
from ._lib import check


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """utils/tal.py:303-315.  feats: list of tensors (.., h, w) or (h, w) tuples.  Host-side constant generation (cached by callers)."""
    pts, sts = [], []
    dev = feats[0].device if hasattr(feats[0], "device") else None
    for f, s in zip(feats, strides):
        h, w = (f.shape[2], f.shape[3]) if hasattr(f, "shape") else f
        sx = torch.arange(w, dtype=torch.float32, device=dev) + grid_cell_offset
        sy = torch.arange(h, dtype=torch.float32, device=dev) + grid_cell_offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        sts.append(torch.full((h * w, 1), float(s), dtype=torch.float32, device=dev))
    return torch.cat(pts), torch.cat(sts)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t):
    return t.detach().float().contiguous()


def tal_assign(pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt, topk=10, alpha=0.5, beta=6.0, eps=1e-9, sums=None):
    bs, n, nc = pd_scores.shape
    m = gt_bboxes.shape[1]
    dev = pd_scores.device
    assert pd_scores.is_cuda, "TaskAlignedAssigner runs on the GPU only (no CPU fallback)"
    pd_scores, pd_bboxes, anc_points = _f32c(pd_scores), _f32c(pd_bboxes), _f32c(anc_points)
    gt_labels, gt_bboxes, mask_gt = _f32c(gt_labels).view(bs, m), _f32c(gt_bboxes), _f32c(mask_gt).view(bs, m)
    target_labels = torch.empty((bs, n), dtype=torch.int64, device=dev)
    target_bboxes = torch.empty((bs, n, 4), dtype=torch.float32, device=dev)
    target_scores = torch.empty((bs, n, nc), dtype=torch.float32, device=dev)
    fg_mask = torch.empty((bs, n), dtype=torch.uint8, device=dev)
    target_gt_idx = torch.empty((bs, n), dtype=torch.int64, device=dev)
    L = ops.lib()
    ws = torch.empty(max(int(L.yad_tal_workspace_bytes(bs, n, m)), 256), dtype=torch.uint8, device=dev)
    ops._count("yad_tal_assign")
    check(L.yad_tal_assign(_p(pd_scores), _p(pd_bboxes), _p(anc_points), _p(gt_labels), _p(gt_bboxes), _p(mask_gt), bs, n, nc, m, topk,
                           alpha, beta, eps, _p(target_labels), _p(target_bboxes), _p(target_scores), _p(fg_mask), _p(target_gt_idx),
                           _p(sums), _p(ws), ops.stream_ptr()), "yad_tal_assign")
    return target_labels, target_bboxes, target_scores, fg_mask.bool(), target_gt_idx


class TaskAlignedAssigner(torch.nn.Module):
    """Same constructor and forward signature as utils/tal.py:13-88."""

    def __init__(self, topk=13, num_classes=80, alpha=1.0, beta=6.0, eps=1e-9):
        super().__init__()
        self.topk, self.num_classes, self.bg_idx, self.alpha, self.beta, self.eps = topk, num_classes, num_classes, alpha, beta, eps

    @torch.no_grad()
    def forward(self, pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt):
        self.bs, self.n_max_boxes = pd_scores.shape[0], gt_bboxes.shape[1]
        return tal_assign(pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt, self.topk, self.alpha, self.beta, self.eps)
