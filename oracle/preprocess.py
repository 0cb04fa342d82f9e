"""CPU ORACLE (test infrastructure, NOT product code) -- the reference's image pre-processing and box post-scaling restated in numpy:
LetterBox (data/augment.py:1475-1600: ratio, rounding of the unpadded size and the borders, cv2.resize INTER_LINEAR, copyMakeBorder 114),
BasePredictor.preprocess (engine/predictor.py:115-133: BGR -> RGB, HWC -> CHW; the /255 stays with the model's first kernel) and
scale_boxes / clip_boxes (utils/ops.py:88-123, 315-334).

cv2.resize(INTER_LINEAR) on uint8 is a THIRD-PARTY algorithm (OpenCV, not under /root/reference; the container has opencv 4.13.0, no version is
pinned by the reference): restated here from its published fixed-point scheme -- 11-bit horizontal / vertical coefficients
(INTER_RESIZE_COEF_BITS), (b0 * (S0 >> 4) >> 16) + (b1 * (S1 >> 4) >> 16) + 2 >> 2 in the vertical pass -- and PINNED bit for bit against the
live cv2 build by oracle/gen_golden.py (tests/golden/preprocess.npz).  Only tests/ may import this file."""
import numpy as np


def _coeffs(src, dst, horizontal):
    """per destination index: the two source indices and their int16 coefficients scaled by 2048 -- cv::resize, INTER_LINEAR, 8U.
    Horizontally OpenCV snaps out-of-range taps to the border AND zeroes the fraction; vertically it only clips the two row indices and keeps
    the fraction (so a border row is the sum of two separately truncated products of the same source row)."""
    scale = src / dst  # double, as cv::resize computes inv_scale from the sizes
    d = np.arange(dst)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(np.float32)
    if horizontal:
        lo = s < 0
        f[lo], s[lo] = 0.0, 0
        hi = s >= src - 1
        f[hi], s[hi] = 0.0, src - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int32)
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int32)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), a0, a1


def resize_linear_u8(img, new_w, new_h):
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC images"""
    h, w = img.shape[:2]
    sx, x1, ax0, ax1 = _coeffs(w, new_w, True)
    sy, y1, by0, by1 = _coeffs(h, new_h, False)
    src = img.astype(np.int32)
    rows = src[:, sx] * ax0[None, :, None] + src[:, x1] * ax1[None, :, None]  # (h, new_w, c), scaled by 2048
    r0, r1 = rows[sy] >> 4, rows[y1] >> 4
    out = (((by0[:, None, None] * r0) >> 16) + ((by1[:, None, None] * r1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_geometry(shape_hw, new_shape=(640, 640), auto=False, scaleup=True, center=True, stride=32):
    """data/augment.py:1558-1586 -> (new_unpad (w, h), top, bottom, left, right)"""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape_hw[0], new_shape[1] / shape_hw[1])
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(shape_hw[1] * r)), int(round(shape_hw[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    if center:
        dw /= 2
        dh /= 2
    top, bottom = (int(round(dh - 0.1)) if center else 0), int(round(dh + 0.1))
    left, right = (int(round(dw - 0.1)) if center else 0), int(round(dw + 0.1))
    return new_unpad, top, bottom, left, right


def letterbox(img, new_shape=(640, 640), auto=False, scaleup=True, center=True, stride=32):
    """LetterBox.__call__(image=img) -> padded uint8 HWC (BGR order untouched)"""
    (nw, nh), top, bottom, left, right = letterbox_geometry(img.shape[:2], new_shape, auto, scaleup, center, stride)
    if (img.shape[1], img.shape[0]) != (nw, nh):
        img = resize_linear_u8(img, nw, nh)
    out = np.full((nh + top + bottom, nw + left + right, img.shape[2]), 114, np.uint8)
    out[top:top + nh, left:left + nw] = img
    return out


def preprocess(images, new_shape=(640, 640), auto=False, stride=32):
    """engine/predictor.py:115-133 + :144-156 up to the uint8 tensor: list of HWC BGR uint8 -> (B, 3, H, W) RGB uint8"""
    same = len({im.shape for im in images}) == 1
    lb = [letterbox(im, new_shape, auto=auto and same, stride=stride) for im in images]
    return np.ascontiguousarray(np.stack(lb)[..., ::-1].transpose(0, 3, 1, 2))


def scale_boxes(img1_shape, boxes, img0_shape):
    """utils/ops.py:88-123 (ratio_pad None, padding True, xyxy) + clip_boxes :315-334, fp32 like the torch reference"""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    b = boxes.astype(np.float32).copy()
    b[..., 0] -= np.float32(pad[0])
    b[..., 1] -= np.float32(pad[1])
    b[..., 2] -= np.float32(pad[0])
    b[..., 3] -= np.float32(pad[1])
    b[..., :4] = b[..., :4] / np.float32(gain)
    b[..., 0] = b[..., 0].clip(0, img0_shape[1])
    b[..., 1] = b[..., 1].clip(0, img0_shape[0])
    b[..., 2] = b[..., 2].clip(0, img0_shape[1])
    b[..., 3] = b[..., 3].clip(0, img0_shape[0])
    return b
