"""TEST INFRASTRUCTURE -- numpy restatement of the training-time augmentations of the reference (data/augment.py) that row f4 moves to the device:
RandomHSV (:1301-1378), RandomFlip (:1380-1472) and Mosaic._mosaic4 (:657-713).  cv2 is third-party arithmetic (opencv-python 4.13.0 in this
image; the reference pins no version): its 8-bit BGR<->HSV conversions are restated here and PINNED exhaustively against the live cv2
(oracle/gen_golden.py augment: all 2^24 colours for BGR2HSV, all 180 x 256 x 256 triples for HSV2BGR) and against the reference's own classes on
seeded cases (tests/golden/augment.npz).

cv2.cvtColor(HSV2BGR) is NOT self-consistent: the vectorised body of a row (first 16 * (w // 16) pixels on an SSE/NEON baseline build) fuses
1 - s * f into one rounding and TRUNCATES after * 255; the scalar tail multiplies separately and ROUNDS.  The two differ by at most 1 LSB.
hsv2bgr(..., path="simd") is what the device kernel implements (every pixel of a row whose width is a multiple of the vector step -- 640-wide
training images -- takes it); path="scalar" restates the tail; path="cv2" mixes them per row as cv2 does on this machine."""
import numpy as np


def bgr2hsv(img):
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    sdiv, hdiv = np.zeros(256, np.int64), np.zeros(256, np.int64)
    i = np.arange(1, 256)
    sdiv[1:] = np.rint((255 << 12) / (1.0 * i)).astype(np.int64)
    hdiv[1:] = np.rint((180 << 12) / (6.0 * i)).astype(np.int64)
    v = np.maximum(np.maximum(b, g), r)
    diff = v - np.minimum(np.minimum(b, g), r)
    vr, vg = np.where(v == r, -1, 0), np.where(v == g, -1, 0)
    s = (diff * sdiv[v] + (1 << 11)) >> 12
    h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))))
    h = (h * hdiv[diff] + (1 << 11)) >> 12
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def hsv2bgr(hsv, path="simd", vec_step=16):
    f32, f64 = np.float32, np.float64
    if path == "cv2":
        w = hsv.shape[-2]
        body = (w // vec_step) * vec_step
        out = np.empty_like(hsv)
        out[..., :body, :] = hsv2bgr(hsv[..., :body, :], "simd")
        out[..., body:, :] = hsv2bgr(hsv[..., body:, :], "scalar")
        return out
    h = hsv[..., 0].astype(f32) * f32(6.0 / 180.0)
    s, v = hsv[..., 1].astype(f32) * f32(1.0 / 255.0), hsv[..., 2].astype(f32) * f32(1.0 / 255.0)
    h = np.where(h >= 6, h - f32(6), h)
    sector = np.floor(h).astype(np.int64)
    f = h - sector.astype(f32)
    one = f32(1)
    if path == "simd":  # fma(-s, f, 1): one rounding
        a2 = (f64(1) - s.astype(f64) * f.astype(f64)).astype(f32)
        a3 = (f64(1) - s.astype(f64) * (one - f).astype(f64)).astype(f32)
    else:
        a2, a3 = one - s * f, one - s * (one - f)
    tab = np.stack([v, v * (one - s), v * a2, v * a3], -1)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    idx = sd[np.clip(sector, 0, 5)]
    out = np.stack([np.take_along_axis(tab, idx[..., k:k + 1], -1)[..., 0] for k in range(3)], -1)
    out = np.where((hsv[..., 1] == 0)[..., None], v[..., None], out) * f32(255)
    out = np.floor(out) if path == "simd" else np.rint(out)
    return np.clip(out, 0, 255).astype(np.uint8)


def hsv_luts(r):
    """augment.py:1369-1374 -- r = the three random gains (np.random.uniform(-1, 1, 3) * [hgain, sgain, vgain] + 1)"""
    x = np.arange(0, 256, dtype=np.asarray(r).dtype)
    return (((x * r[0]) % 180).astype(np.uint8), np.clip(x * r[1], 0, 255).astype(np.uint8), np.clip(x * r[2], 0, 255).astype(np.uint8))


def random_hsv(img, r, path="simd"):
    lh, ls, lv = hsv_luts(r)
    hsv = bgr2hsv(img)
    return hsv2bgr(np.stack([lh[hsv[..., 0]], ls[hsv[..., 1]], lv[hsv[..., 2]]], -1), path)


def flip(img, ud=False, lr=False):
    if ud:
        img = np.flipud(img)
    if lr:
        img = np.fliplr(img)
    return np.ascontiguousarray(img)


def flip_boxes_xywhn(b, ud=False, lr=False):
    """Instances.flipud / fliplr on normalised xywh boxes (utils/instance.py): y -> 1 - y, x -> 1 - x"""
    b = np.array(b, np.float32, copy=True)
    if ud:
        b[:, 1] = 1 - b[:, 1]
    if lr:
        b[:, 0] = 1 - b[:, 0]
    return b


def mosaic4_rects(shapes, yc, xc, s):
    """augment.py:684-707: per image (x1a, y1a, x2a, y2a, x1b, y1b, x2b, y2b, padw, padh); shapes = four (h, w)"""
    out = []
    for i, (h, w) in enumerate(shapes):
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
        out.append((x1a, y1a, x2a, y2a, x1b, y1b, x2b, y2b, x1a - x1b, y1a - y1b))
    return out


def mosaic4(imgs, yc, xc, s):
    img4 = np.full((s * 2, s * 2, 3), 114, np.uint8)
    rects = mosaic4_rects([im.shape[:2] for im in imgs], yc, xc, s)
    for im, (x1a, y1a, x2a, y2a, x1b, y1b, x2b, y2b, _, _) in zip(imgs, rects):
        img4[y1a:y2a, x1a:x2a] = im[y1b:y2b, x1b:x2b]
    return img4, rects


def mosaic4_boxes(boxes_xywhn, shapes, rects, s):
    """Mosaic._update_labels + _cat_labels on normalised xywh boxes (augment.py:786-857): per image xywhn -> xyxy pixels of its (h, w) -> + (padw, padh)
    -> concatenated, clipped to the 2s canvas, zero-area boxes removed.  Returns (boxes xyxy (k, 4), keep mask over the concatenation)."""
    out = []
    for b, (h, w), r in zip(boxes_xywhn, shapes, rects):
        b = np.asarray(b, np.float32)
        xyxy = np.stack([b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2], 1).astype(np.float32)
        xyxy[:, [0, 2]] *= w
        xyxy[:, [1, 3]] *= h
        xyxy[:, [0, 2]] += r[8]
        xyxy[:, [1, 3]] += r[9]
        out.append(xyxy)
    allb = np.concatenate(out, 0)
    allb[:, [0, 2]] = allb[:, [0, 2]].clip(0, 2 * s)
    allb[:, [1, 3]] = allb[:, [1, 3]].clip(0, 2 * s)
    good = ((allb[:, 2] - allb[:, 0]) * (allb[:, 3] - allb[:, 1])) > 0
    return allb[good], good
