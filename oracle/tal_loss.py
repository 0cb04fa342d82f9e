"""CPU ORACLE (test infrastructure, NOT product code) -- TaskAlignedAssigner + v8DetectionLoss restated in fp32 torch-CPU,
image by image, without the (B, n_max, N) broadcast tensors of the reference.

Follows /root/reference/ultralytics/utils/tal.py:13-265 (TaskAlignedAssigner), utils/loss.py:18-42 (SlideLoss),
:238-261 (DFLoss), :264-311 (BboxLoss with NWD), :355-520 (v8DetectionLoss.compute_loss),
utils/metrics.py:74-125 (bbox_iou CIoU), :539-564 (wasserstein_loss).
Pinned against the live reference by oracle/gen_golden.py (fixtures tests/golden/tal_*.npz, loss_*.npz).
"""
import math

import torch
import torch.nn.functional as F


def ciou(b1, b2, eps=1e-7):
    """utils/metrics.py:94-125 with xywh=False, CIoU=True.  b1, b2: (...,4) xyxy -> (...,)"""
    b1_x1, b1_y1, b1_x2, b1_y2 = b1.unbind(-1)
    b2_x1, b2_y1, b2_x2, b2_y2 = b2.unbind(-1)
    w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1 + eps
    w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1 + eps
    inter = (torch.minimum(b1_x2, b2_x2) - torch.maximum(b1_x1, b2_x1)).clamp(0) * \
            (torch.minimum(b1_y2, b2_y2) - torch.maximum(b1_y1, b2_y1)).clamp(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.maximum(b1_x2, b2_x2) - torch.minimum(b1_x1, b2_x1)
    ch = torch.maximum(b1_y2, b2_y2) - torch.minimum(b1_y1, b2_y1)
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = ((b2_x1 + b2_x2 - b1_x1 - b1_x2) ** 2 + (b2_y1 + b2_y2 - b1_y1 - b1_y2) ** 2) / 4
    v = (4 / math.pi ** 2) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)).pow(2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


def nwd(pred, target, eps=1e-7, constant=12.8):
    """utils/metrics.py:539-564 wasserstein_loss (returns the similarity exp(-sqrt(W2)/C))."""
    b1_x1, b1_y1, b1_x2, b1_y2 = pred.unbind(-1)
    b2_x1, b2_y1, b2_x2, b2_y2 = target.unbind(-1)
    w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1 + eps
    w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1 + eps
    cx1, cy1 = b1_x1 + w1 / 2, b1_y1 + h1 / 2
    cx2, cy2 = b2_x1 + w2 / 2, b2_y1 + h2 / 2
    cd = (cx1 - cx2) ** 2 + (cy1 - cy2) ** 2 + eps
    wh = ((w1 - w2) ** 2 + (h1 - h2) ** 2) / 4
    return torch.exp(-torch.sqrt(cd + wh) / constant)


def assign_image(pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt, topk=10, alpha=0.5, beta=6.0, eps=1e-9):
    """tal.py:38-88 for ONE image. pd_scores (N,nc) sigmoid, pd_bboxes (N,4) xyxy pixels, anc_points (N,2) pixels,
    gt_labels (M,) , gt_bboxes (M,4), mask_gt (M,) bool/0-1 (padded rows are 0).
    Returns target_labels (N) int64, target_bboxes (N,4), target_scores (N,nc), fg_mask (N) bool, target_gt_idx (N) int64."""
    N, nc = pd_scores.shape
    M = gt_bboxes.shape[0]
    if M == 0:  # tal.py:62-70
        return (torch.full((N,), nc, dtype=torch.int64), torch.zeros(N, 4), torch.zeros(N, nc),
                torch.zeros(N, dtype=torch.bool), torch.zeros(N, dtype=torch.int64))
    mask_gt = mask_gt.reshape(M).float()
    lt, rb = gt_bboxes[:, None, :2], gt_bboxes[:, None, 2:]
    deltas = torch.cat((anc_points[None] - lt, rb - anc_points[None]), 2)  # (M,N,4) tal.py:228-232
    mask_in = (deltas.amin(2) > eps).float()
    m = (mask_in * mask_gt[:, None]).bool()  # tal.py:94
    lab = gt_labels.reshape(M).long()
    overlaps = torch.zeros(M, N)
    scores = torch.zeros(M, N)
    scores[m] = pd_scores[:, lab].t()[m]  # tal.py:113
    iou = ciou(gt_bboxes[:, None, :].expand(M, N, 4), pd_bboxes[None].expand(M, N, 4)).clamp(0)  # tal.py:123-125
    overlaps[m] = iou[m]
    align = scores.pow(alpha) * overlaps.pow(beta)
    # select_topk_candidates tal.py:127-160
    k = min(topk, N)
    _, idx = torch.topk(align, k, dim=-1, largest=True)
    idx = idx.masked_fill(~mask_gt.bool()[:, None].expand(M, k), 0)
    count = torch.zeros(M, N, dtype=torch.int32)
    count.scatter_add_(1, idx, torch.ones_like(idx, dtype=torch.int32))
    count[count > 1] = 0
    mask_pos = count.float() * mask_in * mask_gt[:, None]
    # select_highest_overlaps tal.py:234-265
    fg = mask_pos.sum(0)
    if fg.max() > 1:
        multi = (fg[None] > 1).expand(M, N)
        is_max = torch.zeros(M, N)
        is_max.scatter_(0, overlaps.argmax(0)[None], 1.0)
        mask_pos = torch.where(multi, is_max, mask_pos)
        fg = mask_pos.sum(0)
    tgt_idx = mask_pos.argmax(0)
    # get_targets tal.py:162-208
    t_labels = lab[tgt_idx].clamp(0)
    t_boxes = gt_bboxes[tgt_idx]
    t_scores = F.one_hot(t_labels, nc).float() * (fg > 0)[:, None]
    # normalise tal.py:80-86
    align = align * mask_pos
    pos_align = align.amax(1, keepdim=True)
    pos_ov = (overlaps * mask_pos).amax(1, keepdim=True)
    norm = (align * pos_ov / (pos_align + eps)).amax(0)
    return t_labels, t_boxes, t_scores * norm[:, None], fg > 0, tgt_idx


def assign(pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt, **kw):
    """Batched wrapper with the reference's argument shapes (tal.py:39-57)."""
    outs = [assign_image(pd_scores[b], pd_bboxes[b], anc_points, gt_labels[b].reshape(-1), gt_bboxes[b], mask_gt[b].reshape(-1), **kw)
            for b in range(pd_scores.shape[0])]
    return tuple(torch.stack(o) for o in zip(*outs))


def preprocess_targets(batch_idx, cls, bboxes, batch_size, imgsz_hw):
    """loss.py:392-408 + :443-446: (M,) (M,) (M,4 normalised xywh) -> gt_labels (B,n_max,1), gt_bboxes (B,n_max,4) xyxy px, mask_gt."""
    batch_idx = batch_idx.reshape(-1)
    if batch_idx.numel() == 0:
        out = torch.zeros(batch_size, 0, 5)
    else:
        counts = torch.stack([(batch_idx == j).sum() for j in range(batch_size)])
        out = torch.zeros(batch_size, int(counts.max()), 5)
        for j in range(batch_size):
            sel = batch_idx == j
            n = int(sel.sum())
            if n:
                out[j, :n, 0] = cls.reshape(-1)[sel]
                out[j, :n, 1:] = bboxes[sel]
        h, w = imgsz_hw
        xywh = out[..., 1:5] * torch.tensor([w, h, w, h], dtype=torch.float32)
        out[..., 1:3] = xywh[..., :2] - xywh[..., 2:] / 2
        out[..., 3:5] = xywh[..., :2] + xywh[..., 2:] / 2
    gt_labels, gt_bboxes = out[..., :1], out[..., 1:5]
    mask_gt = (gt_bboxes.sum(2, keepdim=True) > 0).float()
    return gt_labels, gt_bboxes, mask_gt


def detection_loss(feats, batch_idx, cls, bboxes, strides=(8, 16, 32), nc=80, reg_max=16, gains=(7.5, 0.5, 1.5), topk=10):
    """loss.py:426-520 compute_loss + :419-424 __call__.  feats: list of (B, 4*reg_max+nc, H, W) (requires_grad allowed).
    Returns (loss_sum*B, loss_items[3], aux dict)."""
    from .model import make_anchors
    B = feats[0].shape[0]
    no = nc + 4 * reg_max
    cat = torch.cat([f.reshape(B, no, -1) for f in feats], 2)
    pred_distri = cat[:, :4 * reg_max].permute(0, 2, 1).contiguous()
    pred_scores = cat[:, 4 * reg_max:].permute(0, 2, 1).contiguous()
    imgsz = (feats[0].shape[2] * strides[0], feats[0].shape[3] * strides[0])
    anchor_points, stride_tensor = make_anchors([f.shape[2:] for f in feats], strides)
    gt_labels, gt_bboxes, mask_gt = preprocess_targets(batch_idx, cls, bboxes, B, imgsz)
    proj = torch.arange(reg_max, dtype=torch.float32)
    dist = pred_distri.view(B, -1, 4, reg_max).softmax(3).matmul(proj)  # loss.py:410-417
    pred_bboxes = torch.cat((anchor_points - dist[..., :2], anchor_points + dist[..., 2:]), -1)
    t_labels, t_boxes, t_scores, fg, t_idx = assign(pred_scores.detach().sigmoid(), (pred_bboxes.detach() * stride_tensor),
                                                    anchor_points * stride_tensor, gt_labels, gt_bboxes, mask_gt, topk=topk)
    tss = max(t_scores.sum(), 1)
    loss = torch.zeros(3)
    auto_iou = -1.0
    if fg.sum():
        t_boxes = t_boxes / stride_tensor
        w = t_scores.sum(-1)[fg].unsqueeze(-1)
        pb, tb = pred_bboxes[fg], t_boxes[fg]
        iou = ciou(pb, tb).unsqueeze(-1)
        l_iou = ((1.0 - iou) * w).sum() / tss
        l_nwd = ((1.0 - nwd(pb, tb).unsqueeze(-1)) * w).sum() / tss
        loss[0] = 0.5 * l_iou + 0.5 * l_nwd  # loss.py:298-301
        tgt = torch.cat((anchor_points - t_boxes[..., :2], t_boxes[..., 2:] - anchor_points), -1).clamp(0, reg_max - 1 - 0.01)[fg]
        pd = pred_distri[fg].view(-1, reg_max)
        tl = tgt.long()
        wl = (tl + 1) - tgt
        dfl = (F.cross_entropy(pd, tl.view(-1), reduction="none").view(tl.shape) * wl
               + F.cross_entropy(pd, (tl + 1).view(-1), reduction="none").view(tl.shape) * (1 - wl)).mean(-1, keepdim=True)
        loss[2] = (dfl * w).sum() / tss
        auto_iou = float(iou.detach().mean())
    # SlideLoss loss.py:25-42, 510-515
    bce = F.binary_cross_entropy_with_logits(pred_scores, t_scores, reduction="none")
    a = max(auto_iou, 0.2)
    mod = 1.0 * (t_scores <= a - 0.1) + math.exp(1.0 - a) * ((t_scores > a - 0.1) & (t_scores < a)) + torch.exp(-(t_scores - 1.0)) * (t_scores >= a)
    loss[1] = (bce * mod).sum() / tss
    loss = loss * torch.tensor(gains)
    aux = dict(target_labels=t_labels, target_bboxes=t_boxes, target_scores=t_scores, fg_mask=fg, target_gt_idx=t_idx,
               pred_bboxes=pred_bboxes, auto_iou=auto_iou, target_scores_sum=float(tss))
    return loss.sum() * B, loss.detach(), aux
