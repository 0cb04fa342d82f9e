"""Row f4, GPU: yad_hsv_lut / yad_flip / yad_mosaic4 through the host mirror yolo_ad_refine_b200/augment.py -- byte-exact against the fixtures of the
live reference (tests/golden/augment.npz) and against the oracle on larger seeded batches (640-wide training images)."""
import random
import zlib

import numpy as np
import pytest
import torch

from oracle import augment as oa
from oracle.cases import AUG_HSV_CASES, AUG_MOSAIC_CASES, aug_image
from yolo_ad_refine_b200 import augment as ya

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_hsv_matches_reference_fixtures(gold):
    g = gold("augment.npz")
    for name, (h, w, seed) in AUG_HSV_CASES.items():
        out = ya.random_hsv_(dev(aug_image(h, w, seed)[None]), g[f"hsv_{name}_r"][None]).cpu().numpy()[0]
        assert zlib.crc32(out.tobytes()) == int(g[f"hsv_{name}_crc"]), name


def test_hsv_batch_of_training_images_matches_oracle_and_the_reference_draw_order():
    imgs = np.stack([aug_image(96, 640, 20 + i) for i in range(6)])
    np.random.seed(5)
    want_r = np.stack([np.random.uniform(-1, 1, 3) * [0.015, 0.7, 0.4] + 1 for _ in range(6)])
    np.random.seed(5)
    out = ya.RandomHSV(0.015, 0.7, 0.4)(dev(imgs)).cpu().numpy()
    for i in range(6):
        np.testing.assert_array_equal(out[i], oa.random_hsv(imgs[i], want_r[i], "simd"))
    odd = aug_image(33, 37, 3)  # not a multiple of 4 pixels / unaligned rows: the kernel's tail path
    np.testing.assert_array_equal(ya.random_hsv_(dev(odd[None]), want_r[:1]).cpu().numpy()[0], oa.random_hsv(odd, want_r[0], "simd"))


def test_flip_matches_reference_fixtures(gold):
    g = gold("augment.npz")
    for name, (h, w, seed) in AUG_HSV_CASES.items():
        boxes = np.random.RandomState(seed).uniform(0.1, 0.9, (5, 4)).astype(np.float32)
        for direction in ("vertical", "horizontal"):
            random.seed(seed)
            out, bx = ya.RandomFlip(0.5, direction)(dev(aug_image(h, w, seed)[None]), [boxes])
            assert zlib.crc32(out.cpu().numpy()[0].tobytes()) == int(g[f"flip_{name}_{direction}_crc"])
            np.testing.assert_allclose(bx[0], g[f"flip_{name}_{direction}_boxes"], rtol=0, atol=1e-7)
    imgs = np.stack([aug_image(48, 64, 40 + i) for i in range(4)])
    out = ya.flip(dev(imgs), [True, False, True, False], [False, True, True, False]).cpu().numpy()
    for i, (u, l) in enumerate([(True, False), (False, True), (True, True), (False, False)]):
        np.testing.assert_array_equal(out[i], oa.flip(imgs[i], u, l))


def test_mosaic4_matches_reference_fixtures(gold):
    g = gold("augment.npz")
    for name, (s, shapes, seed) in AUG_MOSAIC_CASES.items():
        imgs = [aug_image(h, w, seed + i) for i, (h, w) in enumerate(shapes)]
        boxes = []
        for i in range(4):
            bx = np.random.RandomState(seed + 10 + i).uniform(0.2, 0.8, (3, 4)).astype(np.float32)
            bx[:, 2:] *= 0.3
            boxes.append(bx)
        random.seed(seed)
        canv, labels = ya.Mosaic4(s)([tuple(dev(im) for im in imgs)], [boxes])
        out = canv.cpu().numpy()[0]
        assert zlib.crc32(out.tobytes()) == int(g[f"mosaic_{name}_crc"])
        np.testing.assert_allclose(labels[0][0], g[f"mosaic_{name}_boxes_xyxy"], rtol=0, atol=1e-4)


def test_mosaic4_batch_at_640_matches_oracle():
    s = 640
    rs = np.random.RandomState(8)
    groups, centers, want = [], [], []
    for k in range(3):
        shapes = [(640, int(rs.randint(320, 641))) if rs.rand() < 0.5 else (int(rs.randint(320, 641)), 640) for _ in range(4)]
        imgs = [aug_image(h, w, 60 + 4 * k + i) for i, (h, w) in enumerate(shapes)]
        yc, xc = int(rs.randint(320, 960)), int(rs.randint(320, 960))
        groups.append(tuple(dev(im) for im in imgs))
        centers.append((yc, xc))
        want.append(oa.mosaic4(imgs, yc, xc, s)[0])
    out, _ = ya.mosaic4(groups, centers, s)
    out = out.cpu().numpy()
    for k in range(3):
        np.testing.assert_array_equal(out[k], want[k])


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError):
        ya.random_hsv_(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), np.ones((1, 3)))
