"""Row f4, CPU: oracle/augment.py (numpy restatement of RandomHSV / RandomFlip / Mosaic._mosaic4 and of cv2's 8-bit BGR<->HSV arithmetic) against the
fixtures the live reference + cv2 wrote (tests/golden/augment.npz, oracle/gen_golden.py augment)."""
import zlib

import numpy as np

from oracle import augment as oa
from oracle.cases import AUG_HSV_CASES, AUG_MOSAIC_CASES, aug_image


def test_cv2_arithmetic_was_pinned_exhaustively(gold):
    g = gold("augment.npz")
    assert int(g["bgr2hsv_exhaustive_mismatches"]) == 0 and int(g["hsv2bgr_exhaustive_mismatches"]) == 0


def test_random_hsv_matches_reference(gold):
    g = gold("augment.npz")
    for name, (h, w, seed) in AUG_HSV_CASES.items():
        out = oa.random_hsv(aug_image(h, w, seed), g[f"hsv_{name}_r"], path="simd")
        assert zlib.crc32(out.tobytes()) == int(g[f"hsv_{name}_crc"]), name
        np.testing.assert_array_equal(out[::7, ::5], g[f"hsv_{name}_sample"])


def test_hsv_scalar_tail_is_within_one_lsb():
    img = aug_image(40, 64, 9)
    a, b = oa.random_hsv(img, np.array([1.01, 1.4, 0.8]), "simd"), oa.random_hsv(img, np.array([1.01, 1.4, 0.8]), "scalar")
    assert int(np.abs(a.astype(int) - b.astype(int)).max()) <= 1


def test_flip_matches_reference(gold):
    g = gold("augment.npz")
    for name, (h, w, seed) in AUG_HSV_CASES.items():
        boxes = np.random.RandomState(seed).uniform(0.1, 0.9, (5, 4)).astype(np.float32)
        for direction in ("vertical", "horizontal"):
            hit = bool(g[f"flip_{name}_{direction}_hit"])
            ud, lr = hit and direction == "vertical", hit and direction == "horizontal"
            out = oa.flip(aug_image(h, w, seed), ud, lr)
            assert zlib.crc32(out.tobytes()) == int(g[f"flip_{name}_{direction}_crc"])
            np.testing.assert_allclose(oa.flip_boxes_xywhn(boxes, ud, lr), g[f"flip_{name}_{direction}_boxes"], rtol=0, atol=1e-7)


def test_mosaic4_matches_reference(gold):
    g = gold("augment.npz")
    for name, (s, shapes, seed) in AUG_MOSAIC_CASES.items():
        imgs = [aug_image(h, w, seed + i) for i, (h, w) in enumerate(shapes)]
        yc, xc = (int(v) for v in g[f"mosaic_{name}_center"])
        out, rects = oa.mosaic4(imgs, yc, xc, s)
        assert zlib.crc32(out.tobytes()) == int(g[f"mosaic_{name}_crc"])
        np.testing.assert_array_equal(out[::9, ::9], g[f"mosaic_{name}_sample"])
        boxes = []
        for i in range(4):
            bx = np.random.RandomState(seed + 10 + i).uniform(0.2, 0.8, (3, 4)).astype(np.float32)
            bx[:, 2:] *= 0.3
            boxes.append(bx)
        got, good = oa.mosaic4_boxes(boxes, shapes, rects, s)
        np.testing.assert_allclose(got, g[f"mosaic_{name}_boxes_xyxy"], rtol=0, atol=1e-4)
        cls = np.concatenate([np.full((3, 1), float(i), np.float32) for i in range(4)])[good]
        np.testing.assert_array_equal(cls, g[f"mosaic_{name}_cls"])
