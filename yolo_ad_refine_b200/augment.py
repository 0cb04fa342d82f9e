"""Device-side training augmentations (row f4 of SURVEY.md section 8): host mirrors of the reference's RandomHSV, RandomFlip and Mosaic._mosaic4
(data/augment.py:1301-1378, 1380-1472, 657-713) over libyad.so kernels.  Images are HWC uint8 BGR tensors ON THE DEVICE (what the reference's
dataset hands to its transforms, uploaded once); random draws are taken exactly where and how the reference takes them (np.random.uniform(-1, 1, 3)
for the HSV gains, random.random() for a flip, random.uniform for the mosaic centre), or passed in explicitly.  Labels (a few floats per image) are
updated on the host with the reference's arithmetic.  No CPU path: CPU tensors raise."""
import ctypes as C
import random

import numpy as np
import torch

from . import ops
from ._lib import YadMosaicDesc, check


def _dev(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()):
        raise RuntimeError("yolo_ad_refine_b200.augment works on contiguous uint8 CUDA tensors only (no CPU fallback)")
    return t


def hsv_luts(r):
    """augment.py:1369-1374"""
    x = np.arange(0, 256, dtype=np.asarray(r).dtype)
    return np.stack([((x * r[0]) % 180).astype(np.uint8), np.clip(x * r[1], 0, 255).astype(np.uint8), np.clip(x * r[2], 0, 255).astype(np.uint8)])


def random_hsv_(batch, gains):
    """in place on batch (n, h, w, 3) uint8 BGR; gains (n, 3) float64 = the reference's `r` per image"""
    batch = _dev(batch)
    n, h, w, c = batch.shape
    assert c == 3
    luts = torch.from_numpy(np.stack([hsv_luts(g) for g in np.asarray(gains, np.float64).reshape(n, 3)])).to(batch.device)
    ops._count("yad_hsv_lut")
    check(ops.lib().yad_hsv_lut(C.c_void_p(batch.data_ptr()), n, h, w, C.c_void_p(luts.data_ptr()), ops.stream_ptr()), "yad_hsv_lut")
    return batch


def flip(batch, ud, lr):
    """out of place; ud / lr: per-image booleans"""
    batch = _dev(batch)
    n, h, w, c = batch.shape
    flags = torch.tensor([int(bool(a)) | (int(bool(b)) << 1) for a, b in zip(ud, lr)], dtype=torch.uint8).to(batch.device)
    out = torch.empty_like(batch)
    ops._count("yad_flip")
    check(ops.lib().yad_flip(C.c_void_p(batch.data_ptr()), C.c_void_p(out.data_ptr()), n, h, w, C.c_void_p(flags.data_ptr()), ops.stream_ptr()), "yad_flip")
    return out


def mosaic4(groups, centers, s):
    """groups: list of 4-tuples of (h, w, 3) uint8 device images; centers: list of (yc, xc); returns (canvases (n, 2s, 2s, 3), rects per group) with
    rects as oracle-compatible tuples (x1a, y1a, x2a, y2a, x1b, y1b, x2b, y2b, padw, padh) for the label update"""
    n = len(groups)
    dev = groups[0][0].device
    arr = (YadMosaicDesc * (4 * n))()
    all_rects = []
    for gi, (imgs, (yc, xc)) in enumerate(zip(groups, centers)):
        rects = []
        for i, im in enumerate(imgs):
            im = _dev(im)
            h, w = im.shape[:2]
            if i == 0:
                x1a, y1a, x2a, y2a = max(xc - w, 0), max(yc - h, 0), xc, yc
                x1b, y1b, x2b, y2b = w - (x2a - x1a), h - (y2a - y1a), w, h
            elif i == 1:
                x1a, y1a, x2a, y2a = xc, max(yc - h, 0), min(xc + w, s * 2), yc
                x1b, y1b, x2b, y2b = 0, h - (y2a - y1a), min(w, x2a - x1a), h
            elif i == 2:
                x1a, y1a, x2a, y2a = max(xc - w, 0), yc, xc, min(s * 2, yc + h)
                x1b, y1b, x2b, y2b = w - (x2a - x1a), 0, w, min(y2a - y1a, h)
            else:
                x1a, y1a, x2a, y2a = xc, yc, min(xc + w, s * 2), min(s * 2, yc + h)
                x1b, y1b, x2b, y2b = 0, 0, min(w, x2a - x1a), min(y2a - y1a, h)
            arr[4 * gi + i] = YadMosaicDesc(im.data_ptr(), w, x1a, y1a, x2a, y2a, x1b, y1b, 0)
            rects.append((x1a, y1a, x2a, y2a, x1b, y1b, x2b, y2b, x1a - x1b, y1a - y1b))
        all_rects.append(rects)
    desc = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
    out = torch.empty((n, 2 * s, 2 * s, 3), dtype=torch.uint8, device=dev)
    ops._count("yad_mosaic4")
    check(ops.lib().yad_mosaic4(C.c_void_p(out.data_ptr()), 2 * s, n, C.c_void_p(desc.data_ptr()), ops.stream_ptr()), "yad_mosaic4")
    return out, all_rects


class RandomHSV:
    """data/augment.py:1301-1378 for a device batch: the gains are drawn per image with np.random.uniform(-1, 1, 3), in batch order"""

    def __init__(self, hgain=0.5, sgain=0.5, vgain=0.5):
        self.hgain, self.sgain, self.vgain = hgain, sgain, vgain

    def __call__(self, batch):
        if self.hgain or self.sgain or self.vgain:
            r = np.stack([np.random.uniform(-1, 1, 3) * [self.hgain, self.sgain, self.vgain] + 1 for _ in range(batch.shape[0])])
            random_hsv_(batch, r)
        return batch


class RandomFlip:
    """data/augment.py:1380-1472: one random.random() draw per image; boxes normalised xywh (n_i, 4) arrays per image are flipped on the host"""

    def __init__(self, p=0.5, direction="horizontal"):
        assert direction in {"horizontal", "vertical"} and 0 <= p <= 1.0
        self.p, self.direction = p, direction

    def __call__(self, batch, boxes=None):
        n = batch.shape[0]
        hit = [random.random() < self.p for _ in range(n)]
        ud = hit if self.direction == "vertical" else [False] * n
        lr = hit if self.direction == "horizontal" else [False] * n
        out = flip(batch, ud, lr)
        if boxes is not None:
            boxes = [np.array(b, np.float32, copy=True) for b in boxes]
            for b, u, l in zip(boxes, ud, lr):
                if u:
                    b[:, 1] = 1 - b[:, 1]
                if l:
                    b[:, 0] = 1 - b[:, 0]
            return out, boxes
        return out


def mosaic4_boxes(boxes_xywhn, shapes, rects, s):
    """labels of a mosaic (Mosaic._update_labels + _cat_labels, data/augment.py:786-857), on the host: per-image normalised xywh boxes -> xyxy pixels
    of the 2s canvas, clipped, zero-area boxes dropped.  Returns (boxes (k, 4) float32, keep mask over the concatenated inputs)."""
    out = []
    for b, (h, w), r in zip(boxes_xywhn, shapes, rects):
        b = np.asarray(b, np.float32)
        xyxy = np.stack([b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2], 1).astype(np.float32)
        xyxy[:, [0, 2]] = xyxy[:, [0, 2]] * w + r[8]
        xyxy[:, [1, 3]] = xyxy[:, [1, 3]] * h + r[9]
        out.append(xyxy)
    allb = np.concatenate(out, 0)
    allb[:, [0, 2]] = allb[:, [0, 2]].clip(0, 2 * s)
    allb[:, [1, 3]] = allb[:, [1, 3]].clip(0, 2 * s)
    good = ((allb[:, 2] - allb[:, 0]) * (allb[:, 3] - allb[:, 1])) > 0
    return allb[good], good


class Mosaic4:
    """Mosaic._mosaic4 (data/augment.py:657-713) for device images: the centre is drawn with random.uniform exactly as the reference draws it"""

    def __init__(self, imgsz=640):
        self.imgsz, self.border = imgsz, (-imgsz // 2, -imgsz // 2)

    def __call__(self, groups, boxes=None):
        s = self.imgsz
        centers = [tuple(int(random.uniform(-x, 2 * s + x)) for x in self.border) for _ in groups]
        canv, rects = mosaic4(groups, centers, s)
        if boxes is None:
            return canv
        labels = [mosaic4_boxes(bx, [tuple(im.shape[:2]) for im in imgs], r, s) for bx, imgs, r in zip(boxes, groups, rects)]
        return canv, labels
