"""Device-side image pre-processing and box post-scaling behind the reference's names.

Replaces, for lists of HWC uint8 BGR images (what `cv2.imread` / the reference's loaders deliver):
  * `LetterBox.__call__(image=...)` (data/augment.py:1475-1600) + `BasePredictor.pre_transform` / `preprocess` up to the uint8 NCHW RGB tensor
    (engine/predictor.py:115-133, 144-156)  ->  `DevicePreprocessor.__call__`: ONE host->device copy of the raw images and their descriptor
    table, ONE kernel (`yad_letterbox`) for the whole ragged batch.  The bytes equal the reference's (cv2's fixed-point INTER_LINEAR is reproduced
    bit for bit); the `/255` stays folded into the stem convolution's weights.
  * `scale_boxes` / `clip_boxes` (utils/ops.py:88-123, 315-334; called per image from DetectionPredictor.postprocess,
    models/yolo/detect/predict.py:36-41)  ->  `scale_boxes_batched` on the batched NMS output, and `scale_boxes` with the reference's signature.

Only the scalar geometry (ratio, rounded sizes, borders, gain, pad) is computed on the host, with the reference's own Python arithmetic.
GPU only: there is no CPU fallback (libyad.so must be built)."""
import ctypes as C

import numpy as np
import torch

from . import ops
from ._lib import YadImageDesc

DESC_DTYPE = np.dtype([("src", "<u8"), ("src_h", "<i4"), ("src_w", "<i4"), ("src_pitch", "<i4"), ("new_w", "<i4"), ("new_h", "<i4"),
                       ("top", "<i4"), ("left", "<i4"), ("gain", "<f4"), ("pad_x", "<f4"), ("pad_y", "<f4")])
assert DESC_DTYPE.itemsize == C.sizeof(YadImageDesc) == 48
_ALIGN = 256


def letterbox_params(shape, new_shape=(640, 640), auto=False, scale_fill=False, scaleup=True, center=True, stride=32):
    """LetterBox geometry (data/augment.py:1558-1586) for a source of `shape` (h, w): (new_w, new_h, top, bottom, left, right)."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    h, w = int(shape[0]), int(shape[1])
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    elif scale_fill:
        dw, dh, new_w, new_h = 0.0, 0.0, new_shape[1], new_shape[0]
    if center:
        dw, dh = dw / 2, dh / 2
    top, bottom = (int(round(dh - 0.1)) if center else 0), int(round(dh + 0.1))
    left, right = (int(round(dw - 0.1)) if center else 0), int(round(dw + 0.1))
    return new_w, new_h, top, bottom, left, right


def scale_boxes_params(img1_shape, img0_shape, ratio_pad=None):
    """gain and (pad_x, pad_y) of scale_boxes (utils/ops.py:104-112)"""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    return gain, pad


class PreprocessedBatch:
    """result of DevicePreprocessor.__call__: `im` (B, 3, H, W) uint8 RGB on the device, `desc` the device descriptor table (one 48-byte
    yad_image_desc per image, also the argument of scale_boxes_batched), `orig_shapes` the (h, w) of every source image"""
    __slots__ = ("im", "desc", "orig_shapes")

    def __init__(self, im, desc, orig_shapes):
        self.im, self.desc, self.orig_shapes = im, desc, orig_shapes


class DevicePreprocessor:
    """`LetterBox(imgsz, auto=same_shapes and auto, stride=stride)` over a list of images + stack + BGR->RGB + HWC->CHW on the device.

    The raw images travel to the device in ONE copy from a pinned staging buffer laid out as [descriptor table | image 0 | image 1 | ...];
    `staging_views(shapes)` hands out numpy views of that buffer so a decoder can write into pinned memory directly (then `__call__` copies
    nothing on the host)."""

    def __init__(self, imgsz=640, stride=32, auto=False, scale_fill=False, scaleup=True, center=True, device="cuda", pad_value=114):
        if not torch.cuda.is_available():
            raise RuntimeError("DevicePreprocessor needs a CUDA device: the YOLO-AD-Refine path has no CPU fallback")
        ops.lib()
        self.new_shape = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        self.stride, self.auto, self.scale_fill, self.scaleup, self.center = stride, auto, scale_fill, scaleup, center
        self.pad_value = pad_value
        self.device = torch.device(device)
        # two staging sets used alternately: the host may fill set k + 1 while the copy / kernel of set k is still in flight
        self._sets = [dict(host=None, dev=None, done=None, views=()) for _ in range(2)]
        self._k = 0

    # -- staging ------------------------------------------------------------------------------------------------------------------------
    def _layout(self, shapes):
        offs, off = [], (len(shapes) * DESC_DTYPE.itemsize + _ALIGN - 1) // _ALIGN * _ALIGN
        for h, w in shapes:
            offs.append(off)
            off += (h * w * 3 + _ALIGN - 1) // _ALIGN * _ALIGN
        return offs, off

    def _reserve(self, nbytes):
        """the staging set of the next call, at least nbytes large and no longer in use by the device"""
        st = self._sets[self._k]
        if st["done"] is not None:
            st["done"].synchronize()
        if st["host"] is None or st["host"].numel() < nbytes:
            nbytes = max(nbytes, 1 << 20)
            st["host"] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            st["dev"] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            st["views"] = ()
        return st

    def staging_views(self, shapes):
        """pinned (h, w, 3) uint8 numpy views for images of the given (h, w) shapes, valid for the NEXT call: pass exactly these arrays to
        __call__ to skip the host-side copy (a decoder can write into pinned memory directly)"""
        shapes = [(int(s[0]), int(s[1])) for s in shapes]
        offs, total = self._layout(shapes)
        st = self._reserve(total)
        host = st["host"].numpy()
        st["views"] = tuple(host[o:o + h * w * 3].reshape(h, w, 3) for o, (h, w) in zip(offs, shapes))
        return list(st["views"])

    # -- the call -------------------------------------------------------------------------------------------------------------------------
    def __call__(self, images, out=None):
        """images: list of (h, w, 3) uint8 arrays (BGR).  out: optional (>= B, 3, H, W) uint8 device tensor to write into (e.g. an engine's
        static input); returns a PreprocessedBatch."""
        if len(images) == 0:
            raise ValueError("empty image list")
        for im in images:
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise TypeError(f"expected (h, w, 3) uint8 images, got {im.dtype} {im.shape}")
        shapes = [(im.shape[0], im.shape[1]) for im in images]
        same = len(set(shapes)) == 1  # engine/predictor.py:154
        geo = [letterbox_params(s, self.new_shape, self.auto and same, self.scale_fill, self.scaleup, self.center, self.stride) for s in shapes]
        out_hw = {(nh + t + b, nw + l + r) for nw, nh, t, b, l, r in geo}
        if len(out_hw) != 1:  # np.stack raises in the reference (engine/predictor.py:127)
            raise ValueError(f"all input arrays must have the same shape after LetterBox, got {sorted(out_hw)}")
        H, W = out_hw.pop()
        offs, total = self._layout(shapes)
        views = self._sets[self._k]["views"]
        reuse = len(views) == len(images) and all(a is b for a, b in zip(views, images))
        st = self._sets[self._k] if reuse else self._reserve(total)
        st["views"] = ()
        host = st["host"].numpy()
        desc = host[:len(images) * DESC_DTYPE.itemsize].view(DESC_DTYPE)
        base = st["dev"].data_ptr()
        for i, (im, (h, w), (nw, nh, t, b, l, r), o) in enumerate(zip(images, shapes, geo, offs)):
            if not reuse:
                host[o:o + h * w * 3].reshape(h, w, 3)[...] = im
            gain, pad = scale_boxes_params((H, W), (h, w))
            desc[i] = (base + o, h, w, w * 3, nw, nh, t, l, gain, pad[0], pad[1])
        st["dev"][:total].copy_(st["host"][:total], non_blocking=True)
        B = len(images)
        if out is None:
            out = torch.empty((B, 3, H, W), dtype=torch.uint8, device=self.device)
        else:
            if out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous() or out.shape[0] < B or tuple(out.shape[1:]) != (3, H, W):
                raise ValueError(f"out must be a contiguous uint8 device tensor (>= {B}, 3, {H}, {W}), got {out.dtype} {tuple(out.shape)}")
        desc_dev = st["dev"][:B * DESC_DTYPE.itemsize]
        ops.letterbox(desc_dev, B, out, H, W, self.pad_value, swap_rb=True)
        desc_keep = desc_dev.clone()  # the staging set is recycled two calls later; the (3 KB) table outlives it for scale_boxes_batched
        if st["done"] is None:
            st["done"] = torch.cuda.Event()
        st["done"].record()
        self._k ^= 1
        return PreprocessedBatch(out[:B], desc_keep, shapes)


def scale_boxes_batched(det, count, desc):
    """scale_boxes + clip_boxes for every image of a batched NMS output, in place: det (B, max_det, >= 4) fp32 device tensor whose rows start with
    x1, y1, x2, y2 (e.g. `nms_raw`'s `out`), count (B,) int32 device tensor or None, desc = PreprocessedBatch.desc."""
    ops.scale_boxes(det, count, desc)
    return det


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """Same signature as ultralytics.utils.ops.scale_boxes (utils/ops.py:88-123) for one image's (k, >= 4) fp32 device tensor; in place, returns
    `boxes`.  The batched entry point (scale_boxes_batched) is the fast path; this mirror uploads a one-entry descriptor per call."""
    if not padding or xywh:
        raise NotImplementedError("padding=False / xywh=True are outside the YOLO-AD-Refine detection path")
    assert boxes.is_cuda and boxes.dtype == torch.float32 and boxes.dim() == 2 and boxes.stride(1) == 1, "fp32 (k, >= 4) device rows expected"
    if boxes.shape[0] == 0:
        return boxes
    gain, pad = scale_boxes_params(img1_shape, img0_shape, ratio_pad)
    d = np.zeros(1, DESC_DTYPE)
    d[0] = (0, img0_shape[0], img0_shape[1], 0, 0, 0, 0, 0, gain, pad[0], pad[1])
    desc = torch.from_numpy(d.view(np.uint8)).to(boxes.device)
    ops.scale_boxes(boxes.unsqueeze(0), None, desc, row_ld=boxes.stride(0))
    return boxes
