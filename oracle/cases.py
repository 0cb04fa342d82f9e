"""TEST INFRASTRUCTURE -- the seeded input cases shared by oracle/gen_golden.py (which runs the live reference on them) and the
tests (which run the oracle and the CUDA path on them).  No reference imports here: this module also runs on the GPU box."""
import numpy as np

from . import synth

SAMPLES = 96


def sample_positions(numel, k=SAMPLES, seed=7):
    """deterministic positions of the sub-sampled activations kept in the fixtures"""
    rs = np.random.RandomState(seed + numel % 9973)
    return np.sort(rs.randint(0, numel, k))


NMS_CASES = {
    # name: (make_predictions kwargs, nms kwargs)
    "predict": (dict(batch=4, n_anchors=8400, seed=10), dict(conf_thres=0.25, iou_thres=0.7, max_det=300)),
    "predict_iou45": (dict(batch=2, n_anchors=8400, seed=11), dict(conf_thres=0.25, iou_thres=0.45, max_det=300)),
    "agnostic": (dict(batch=2, n_anchors=8400, seed=12), dict(conf_thres=0.25, iou_thres=0.6, agnostic=True, max_det=300)),
    "classes": (dict(batch=2, n_anchors=8400, seed=13), dict(conf_thres=0.25, iou_thres=0.7, classes=[0, 3, 17, 79], max_det=300)),
    "maxdet": (dict(batch=2, n_anchors=8400, seed=14, p_hot=0.5, cluster=False), dict(conf_thres=0.25, iou_thres=0.7, max_det=50)),
    "val_multilabel": (dict(batch=2, n_anchors=2100, seed=15, p_hot=0.2), dict(conf_thres=0.001, iou_thres=0.7, multi_label=True, max_det=300)),
    "val_truncate": (dict(batch=1, n_anchors=2100, seed=16, p_hot=0.2),
                     dict(conf_thres=0.001, iou_thres=0.7, multi_label=True, max_det=300, max_nms=30000)),
    "none_pass": (dict(batch=2, n_anchors=525, seed=17, p_hot=0.0), dict(conf_thres=0.25, iou_thres=0.7, max_det=300)),
    "odd_sizes": (dict(batch=3, n_anchors=1333, seed=18, p_hot=0.3, nc=5), dict(conf_thres=0.3, iou_thres=0.5, max_det=300)),
}


def train_inputs(batch, seed, hw=(20, 10, 5), empty_images=(0,), no_targets=False):
    """Raw train-mode head outputs (3 levels) + targets, from seeds. Logits are shaped so that predicted boxes are a few cells wide."""
    rs = np.random.RandomState(seed)
    feats = []
    for s in hw:
        f = rs.standard_normal((batch, 144, s, s)).astype(np.float32)
        f[:, :64] *= np.float32(1.5)
        f[:, 64:] = np.float32(-3.0) + np.float32(1.5) * f[:, 64:]
        feats.append(f)
    bi, cl, bb = synth.make_targets(batch, seed=seed + 1, max_per_img=12, empty_images=empty_images)
    if no_targets:
        bi, cl, bb = bi[:0], cl[:0], bb[:0]
    return feats, bi, cl, bb


TRAIN_CASES = {
    "small": dict(batch=4, seed=30, hw=(20, 10, 5)),
    "b2_640": dict(batch=2, seed=31, hw=(80, 40, 20), empty_images=()),
    "no_targets": dict(batch=2, seed=32, hw=(20, 10, 5), no_targets=True),
}



TRAIN_STEP_CASES = {
    "b2_160": dict(batch=2, size=160, seed=50),
    "b3_192x256": dict(batch=3, size=(192, 256), seed=52, empty_images=(1,)),
}


def train_step_inputs(batch, size, seed, empty_images=()):
    """images (B,3,H,W) in [0,1] + targets for one full training step (forward + loss + backward)"""
    h, w = (size, size) if isinstance(size, int) else size
    img = synth.make_images(batch, h, w, seed=seed)
    bi, cl, bb = synth.make_targets(batch, seed=seed + 1, max_per_img=6, empty_images=empty_images)
    return img, bi, cl, bb


def tal_inputs():
    """Stand-alone TaskAlignedAssigner call with padded gts (tal.py:38-88): B=3 images with 9, 4 and 0 valid gts."""
    import torch  # noqa: F401
    from .model import make_anchors as mk
    rs = np.random.RandomState(40)
    B, N, nc, M = 3, 2100, 80, 9
    from oracle.model import make_anchors as mk
    anc, st = mk([(40, 40), (20, 20), (10, 10)], (8, 16, 32))
    anc_px = (anc * st).numpy()
    pd_scores = rs.uniform(0, 1, (B, N, nc)).astype(np.float32) ** 4
    ctr = anc_px[None] + rs.standard_normal((B, N, 2)).astype(np.float32) * 4
    wh = rs.uniform(10, 120, (B, N, 2)).astype(np.float32)
    pd_bboxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(np.float32)
    gt_bboxes = np.zeros((B, M, 4), np.float32)
    gt_labels = np.zeros((B, M, 1), np.float32)
    mask_gt = np.zeros((B, M, 1), np.float32)
    for b, n in enumerate((9, 4, 0)):
        c = rs.uniform(60, 260, (n, 2)); s = rs.uniform(20, 150, (n, 2))
        gt_bboxes[b, :n] = np.concatenate([c - s / 2, c + s / 2], -1)
        gt_labels[b, :n, 0] = rs.randint(0, nc, n)
        mask_gt[b, :n] = 1
    return pd_scores, pd_bboxes, anc_px, gt_labels, gt_bboxes, mask_gt


# pre-processing cases: (source h, source w, target size, auto / minimum-rectangle, seed)
PREPROCESS_CASES = {
    "vga_to_160": (480, 640, 160, False, 60),
    "hd_to_160_auto": (720, 1280, 160, True, 61),
    "odd_to_192": (333, 500, 192, False, 62),
    "upscale_to_128": (50, 40, 128, False, 63),
    "same_size": (160, 160, 160, False, 64),
    "portrait_auto": (641, 480, 160, True, 65),
    "tiny": (7, 9, 64, False, 66),
    "half": (256, 320, 160, False, 67),          # exact 2x: cv2 switches INTER_LINEAR to its INTER_AREA fast path, same bytes
    "wide_to_160": (90, 400, 160, False, 68),
    "tall_to_160": (301, 77, 160, False, 69),
}
# ragged batch: images of different sizes that letterbox to the same 160 x 160 tensor (one yad_letterbox launch)
PREPROCESS_BATCH = ["vga_to_160", "same_size", "half", "wide_to_160", "tall_to_160"]
# random (source h, source w) -> 96 x 96 sweep: the fixture stores only a CRC32 of the reference's bytes per case
PREPROCESS_SWEEP = [(int(h), int(w), 200 + i) for i, (h, w) in enumerate(np.random.RandomState(7).randint(5, 700, (48, 2)))]


def preprocess_image(h, w, seed):
    """uint8 HWC 'BGR' image with structure (gradients + noise), so that interpolation errors are visible"""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 255 // max(h + w - 2, 1))], -1)
    return np.clip(base + rs.randint(-40, 41, (h, w, 3)), 0, 255).astype(np.uint8)


def scale_boxes_inputs(seed=70, n=64):
    rs = np.random.RandomState(seed)
    b = rs.uniform(-20, 700, (n, 4)).astype(np.float32)
    return np.concatenate([np.minimum(b[:, :2], b[:, 2:]), np.maximum(b[:, :2], b[:, 2:])], 1)


# validator matching cases: name -> (n detections, m labels, nc, jitter px, seed)
MATCH_CASES = {
    "typical": (120, 18, 5, 6.0, 80),
    "crowded": (300, 64, 3, 10.0, 81),
    "no_labels": (40, 0, 5, 4.0, 82),
    "no_detections": (0, 12, 5, 4.0, 83),
    "single_pair": (1, 1, 1, 2.0, 84),
    "many_classes": (200, 40, 80, 5.0, 85),
    "one_label_many_dets": (60, 1, 1, 8.0, 86),
}


def match_inputs(n, m, nc, jitter, seed):
    """detections (n, 6) sorted by confidence descending (as NMS returns them) scattered around m ground-truth boxes; continuous fp32 -> no IoU ties"""
    rs = np.random.RandomState(seed)
    c = rs.uniform(60, 580, (m, 2))
    wh = rs.uniform(20, 160, (m, 2))
    gt = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    gt_cls = rs.randint(0, nc, m).astype(np.float32)
    det = np.zeros((n, 6), np.float32)
    if n:
        if m:
            src = rs.randint(0, m, n)
            det[:, :4] = gt[src] + rs.normal(0, jitter, (n, 4))
            det[:, 5] = np.where(rs.rand(n) < 0.8, gt_cls[src], rs.randint(0, nc, n))
        else:
            c = rs.uniform(60, 580, (n, 2))
            det[:, :4] = np.concatenate([c - 30, c + 30], 1)
            det[:, 5] = rs.randint(0, nc, n)
        det[:, 4] = np.sort(rs.uniform(0.001, 1, n))[::-1]
    return det.astype(np.float32), gt, gt_cls


# validation-set cases (update_metrics -> get_stats -> ap_per_class): name -> (images, nc, max detections / image,
# max labels / image, jitter px, seed)
VAL_CASES = {
    "small": (6, 3, 40, 6, 5.0, 90),
    "coco_like": (48, 80, 300, 24, 6.0, 91),
    "few_classes_many_dets": (32, 4, 300, 40, 9.0, 92),
    "sparse": (12, 20, 12, 3, 4.0, 93),
}


def val_inputs(n_img, nc, max_det, max_lab, jitter, seed, imgsz=640):
    """A seeded validation set in the layout DetectionValidator.update_metrics sees (models/yolo/detect/val.py:126-176): per image the NMS output
    in letterboxed-image pixels (k_i, 6) sorted by confidence, the labels as normalised xywh of the letterboxed image + batch_idx + cls (the
    collated batch dict), ori_shape and ratio_pad = ((gain, gain), (left, top)) as data/base.py:295 + LetterBox (data/augment.py:1590-1591) set them.
    Some images have no labels, some no detections.  Continuous fp32 coordinates and confidences: no ties."""
    rs = np.random.RandomState(seed)
    dets, ori, ratio_pad, bboxes, cls, bidx = [], [], [], [], [], []
    for i in range(n_img):
        h0, w0 = int(rs.randint(240, 1100)), int(rs.randint(240, 1100))
        gain = min(imgsz / h0, imgsz / w0)
        nw, nh = int(round(w0 * gain)), int(round(h0 * gain))
        left, top = int(round((imgsz - nw) / 2 - 0.1)), int(round((imgsz - nh) / 2 - 0.1))
        m = 0 if i % 7 == 3 else int(rs.randint(1, max_lab + 1))
        n = 0 if i % 11 == 5 else int(rs.randint(1, max_det + 1))
        c = np.stack([rs.uniform(left + 20, left + nw - 20, m), rs.uniform(top + 20, top + nh - 20, m)], 1)
        wh = rs.uniform(12, 0.45 * min(nw, nh), (m, 2))
        gt = np.concatenate([c - wh / 2, c + wh / 2], 1)
        gt_cls = rs.randint(0, nc, m)
        det = np.zeros((n, 6), np.float32)
        if n:
            if m:
                src = rs.randint(0, m, n)
                det[:, :4] = gt[src] + rs.normal(0, jitter, (n, 4))
                det[:, 5] = np.where(rs.rand(n) < 0.75, gt_cls[src], rs.randint(0, nc, n))
            else:
                cc = rs.uniform(60, 580, (n, 2))
                det[:, :4] = np.concatenate([cc - 30, cc + 30], 1)
                det[:, 5] = rs.randint(0, nc, n)
            det[:, 4] = np.sort(rs.uniform(0.001, 1, n))[::-1]
        dets.append(det.astype(np.float32))
        ori.append((h0, w0))
        ratio_pad.append(((gain, gain), (left, top)))
        bboxes.append(np.concatenate([(gt[:, :2] + gt[:, 2:]) / 2, gt[:, 2:] - gt[:, :2]], 1) / imgsz)
        cls.append(gt_cls)
        bidx.append(np.full(m, i))
    return dict(dets=dets, ori_shape=ori, ratio_pad=ratio_pad, bboxes=np.concatenate(bboxes).astype(np.float32),
                cls=np.concatenate(cls).astype(np.float32), batch_idx=np.concatenate(bidx).astype(np.int64), imgsz=imgsz)


# ---- row f4: augmentation cases (oracle/augment.py) --------------------------------------------------------------------------------------
# name -> (h, w, seed).  Widths are multiples of 64 so that every pixel of a row takes cv2's vectorised HSV2BGR path (see oracle/augment.py).
AUG_HSV_CASES = {"a": (96, 128, 1), "b": (64, 192, 2), "c": (50, 64, 3), "d": (160, 640, 4)}
# name -> (s, four (h, w) source shapes, seed): letterboxed-to-s images as the reference's dataset delivers them (long side = s)
AUG_MOSAIC_CASES = {"s96": (96, [(96, 72), (64, 96), (96, 96), (80, 96)], 11), "s160": (160, [(160, 120), (160, 160), (90, 160), (160, 100)], 12)}


def aug_image(h, w, seed):
    return np.random.RandomState(1000 + seed).randint(0, 256, (h, w, 3)).astype(np.uint8)
