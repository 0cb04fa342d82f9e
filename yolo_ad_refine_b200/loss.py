"""Drop-in for ultralytics.utils.loss.v8DetectionLoss (utils/loss.py:355-520) as this fork defines it: box = 0.5 (1-CIoU) + 0.5 (1-NWD),
DFL, SlideLoss-wrapped BCE, TaskAlignedAssigner(topk=10, alpha=0.5, beta=6.0).  Forward AND the gradients w.r.t. the raw head outputs are
computed by libyad.so kernels; all reductions stay on the device (no host synchronisation inside the loss)."""
import ctypes as C

import torch

from . import ops
from ._lib import check
from .tal import make_anchors, tal_assign


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def preprocess_targets(batch_idx, cls, bboxes, batch_size, imgsz_hw, device):
    """utils/loss.py:392-408 + :443-446 -- host-side ragged -> padded packing of the ground truth (input plumbing, as in the reference)."""
    n_max_host = None
    if batch_idx.numel() and not batch_idx.is_cuda:  # the dataloader delivers labels on the host: the padded width costs no device sync
        n_max_host = int(torch.bincount(batch_idx.reshape(-1).long(), minlength=batch_size).max())
    batch_idx = batch_idx.reshape(-1).to(device)
    cls, bboxes = cls.reshape(-1).to(device).float(), bboxes.reshape(-1, 4).to(device).float()
    if batch_idx.numel() == 0:
        out = torch.zeros(batch_size, 0, 5, device=device)
    else:
        bi = batch_idx.long()
        counts = torch.bincount(bi, minlength=batch_size)
        n_max = n_max_host if n_max_host is not None else int(counts.max())
        order = torch.argsort(bi, stable=True)
        start = torch.cumsum(counts, 0) - counts
        rank = torch.arange(bi.numel(), device=device) - start[bi[order]]
        out = torch.zeros(batch_size, n_max, 5, device=device)
        h, w = imgsz_hw
        xywh = bboxes[order] * torch.tensor([w, h, w, h], dtype=torch.float32, device=device)
        out[bi[order], rank, 0] = cls[order]
        out[bi[order], rank, 1:3] = xywh[:, :2] - xywh[:, 2:] / 2
        out[bi[order], rank, 3:5] = xywh[:, :2] + xywh[:, 2:] / 2
    gt_labels, gt_bboxes = out[..., :1].contiguous(), out[..., 1:5].contiguous()
    mask_gt = (gt_bboxes.sum(2, keepdim=True) > 0).float()
    return gt_labels, gt_bboxes, mask_gt


def pack_targets_static(batch_idx, cls, bboxes, batch_size, imgsz_hw, n_max):
    """preprocess_targets for a CUDA-graph step: device tensors of FIXED length (padding rows carry batch_idx < 0), fixed n_max, no host
    synchronisation and no data-dependent shapes.  batch_idx must be non-decreasing over the valid rows, which is how the reference's collate_fn
    builds it (data/dataset.py:230-246).  Targets beyond n_max per image are dropped (choose n_max from the dataset's label statistics)."""
    dev = batch_idx.device
    bi = batch_idx.reshape(-1).long()
    M = bi.numel()
    valid = (bi >= 0) & (bi < batch_size)
    bic = torch.where(valid, bi, torch.full_like(bi, batch_size))
    counts = torch.zeros(batch_size + 1, dtype=torch.long, device=dev).scatter_add_(0, bic, torch.ones_like(bic))
    start = torch.cumsum(counts, 0) - counts
    rank = torch.arange(M, device=dev) - start[bic]
    ok = valid & (rank < n_max)
    row = torch.where(ok, bic, torch.full_like(bic, batch_size))       # trash row for padding / overflow
    col = torch.where(ok, rank, torch.zeros_like(rank))
    out = torch.zeros(batch_size + 1, n_max, 5, device=dev)
    h, w = imgsz_hw
    b = bboxes.reshape(-1, 4).float()
    cx, cy, bw, bh = b[:, 0] * float(w), b[:, 1] * float(h), b[:, 2] * float(w), b[:, 3] * float(h)  # python scalars: no host tensor during capture
    vals = torch.stack([cls.reshape(-1).float(), cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    out[row, col] = vals
    out = out[:batch_size]
    gt_labels, gt_bboxes = out[..., :1].contiguous(), out[..., 1:5].contiguous()
    mask_gt = (gt_bboxes.sum(2, keepdim=True) > 0).float()
    return gt_labels, gt_bboxes, mask_gt


def detection_loss_raw(pred_distri, pred_scores, anchor_points, stride_tensor, gt_labels, gt_bboxes, mask_gt, gains=(7.5, 0.5, 1.5), topk=10,
                       reg_max=16, want_grad=True, assign=None):
    """pred_distri (B,N,4*reg_max), pred_scores (B,N,nc) fp32 contiguous logits.  Returns (out4 = [box, cls, dfl, total*B], grad_distri,
    grad_scores, aux).  assign: the `aux` of another call -- its (no-grad) TaskAlignedAssigner outputs are used instead of running the assigner,
    which lets a lower-precision step be compared with an fp32 step on the SAME discrete targets (tests)."""
    B, N, nc = pred_scores.shape
    dev = pred_scores.device
    L = ops.lib()
    st = ops.stream_ptr()
    sums = torch.zeros(8, dtype=torch.float64, device=dev)
    boxes = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    boxes_px = torch.empty_like(boxes)
    sig = torch.empty_like(pred_scores)
    anc = anchor_points.float().contiguous()
    stv = stride_tensor.float().reshape(-1).contiguous()
    ops._count("yad_loss_decode")
    check(L.yad_loss_decode(_p(pred_distri), _p(pred_scores), _p(anc), _p(stv), B, N, nc, reg_max, _p(boxes), _p(boxes_px), _p(sig), st),
          "yad_loss_decode")
    anc_px = (anc * stv[:, None]).contiguous()
    if assign is None:
        tl, tb, ts, fg, tgi = tal_assign(sig, boxes_px, anc_px, gt_labels, gt_bboxes, mask_gt, topk, 0.5, 6.0, 1e-9, sums=sums)
    else:
        tl, tb, ts, fg, tgi = (assign[k] for k in ("target_labels", "target_bboxes", "target_scores", "fg_mask", "target_gt_idx"))
        sums[4], sums[5] = fg.sum().double(), ts.sum().double()  # what yad_tal_assign leaves there: positives, target_scores_sum
    fg_u8 = fg.to(torch.uint8)
    gd = torch.empty_like(pred_distri) if want_grad else None
    gs = torch.empty_like(pred_scores) if want_grad else None
    ops._count("yad_loss_bbox")
    check(L.yad_loss_bbox(_p(pred_distri), _p(boxes), _p(anc), _p(stv), _p(tb), _p(ts), _p(fg_u8), B, N, nc, reg_max, _p(sums), gains[0], gains[2],
                          _p(gd), st), "yad_loss_bbox")
    ops._count("yad_loss_cls")
    check(L.yad_loss_cls(_p(pred_scores), _p(ts), B, N, nc, _p(sums), gains[1], _p(gs), st), "yad_loss_cls")
    out4 = torch.empty(4, dtype=torch.float32, device=dev)
    ops._count("yad_loss_finalize")
    check(L.yad_loss_finalize(_p(sums), gains[0], gains[1], gains[2], B, _p(out4), st), "yad_loss_finalize")
    aux = dict(target_labels=tl, target_bboxes=tb, target_scores=ts, fg_mask=fg, target_gt_idx=tgi, pred_bboxes=boxes, sums=sums)
    return out4, gd, gs, aux


class _DetLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_distri, pred_scores, anchor_points, stride_tensor, gt_labels, gt_bboxes, mask_gt, gains, topk, reg_max):
        out4, gd, gs, aux = detection_loss_raw(pred_distri.detach().contiguous(), pred_scores.detach().contiguous(), anchor_points, stride_tensor,
                                               gt_labels, gt_bboxes, mask_gt, gains, topk, reg_max, True)
        ctx.save_for_backward(gd, gs)
        ctx.aux = aux
        items = out4[:3].clone()
        ctx.mark_non_differentiable(items)
        return out4[3].clone(), items

    @staticmethod
    def backward(ctx, g_total, _g_items):
        gd, gs = ctx.saved_tensors
        return gd * g_total, gs * g_total, None, None, None, None, None, None, None, None


class v8DetectionLoss:
    """Criterion with the reference's constructor and call signature (utils/loss.py:355-424)."""

    def __init__(self, model, tal_topk=10):
        m = model.model[-1]
        self.hyp = model.args
        self.stride = torch.as_tensor(m.stride).float()
        self.nc, self.reg_max = m.nc, m.reg_max
        self.no = m.nc + m.reg_max * 4
        self.topk = tal_topk
        self.device = next(model.parameters()).device
        self._anchor_cache = {}
        self.last_aux = None

    def _hyp(self, k):
        return float(self.hyp[k] if isinstance(self.hyp, dict) else getattr(self.hyp, k))

    def __call__(self, preds, batch):
        feats = preds[1] if isinstance(preds, tuple) else preds
        feats = feats[: self.stride.numel()]
        B = feats[0].shape[0]
        dev = feats[0].device
        # (B, no, H, W) x 3 -> (B, N, 64) / (B, N, nc): layout plumbing, exactly the reference's cat / permute (loss.py:430-436)
        cat = torch.cat([f.reshape(B, self.no, -1) for f in feats], 2).float()
        pred_distri = cat[:, : self.reg_max * 4].permute(0, 2, 1).contiguous()
        pred_scores = cat[:, self.reg_max * 4:].permute(0, 2, 1).contiguous()
        shapes = tuple((f.shape[2], f.shape[3]) for f in feats)
        if shapes not in self._anchor_cache:
            a, s = make_anchors(list(shapes), self.stride.tolist())
            self._anchor_cache[shapes] = (a.to(dev), s.to(dev))
        anchor_points, stride_tensor = self._anchor_cache[shapes]
        imgsz = (shapes[0][0] * float(self.stride[0]), shapes[0][1] * float(self.stride[0]))
        gt_labels, gt_bboxes, mask_gt = preprocess_targets(batch["batch_idx"], batch["cls"], batch["bboxes"], B, imgsz, dev)
        gains = (self._hyp("box"), self._hyp("cls"), self._hyp("dfl"))
        total, items = _DetLossFn.apply(pred_distri, pred_scores, anchor_points, stride_tensor, gt_labels, gt_bboxes, mask_gt, gains, self.topk,
                                        self.reg_max)
        return total, items
