"""GPU parity (through the C ABI: yad_letterbox, yad_scale_boxes): device LetterBox + BGR->RGB + CHW and scale_boxes against the fixtures written
by the live reference (cv2 + ultralytics, tests/golden/preprocess.npz) and against the oracle.  Bytes and fp32 boxes must be IDENTICAL."""
import zlib

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import preprocess as op
from util_gpu import DEV
from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.preprocess import DevicePreprocessor, scale_boxes, scale_boxes_batched

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(cases.PREPROCESS_CASES))
def test_letterbox_bit_exact_vs_reference_golden(gold, name):
    g = gold("preprocess.npz")
    h, w, size, auto, seed = cases.PREPROCESS_CASES[name]
    img = cases.preprocess_image(h, w, seed)
    pre = DevicePreprocessor(size, stride=32, auto=auto, device=DEV)
    pb = pre([img])
    got = pb.im.cpu().numpy()[0]
    assert got.shape == g[name].shape
    np.testing.assert_array_equal(got, g[name])
    # scale_boxes: batched entry point (descriptor of this image) and the reference-signature mirror
    boxes = cases.scale_boxes_inputs(seed + 100) * np.float32(size / 640.0)
    det = torch.zeros(1, boxes.shape[0], 6, device=DEV)
    det[0, :, :4] = torch.from_numpy(boxes).to(DEV)
    det[0, :, 4] = 0.5
    cnt = torch.tensor([boxes.shape[0] - 3], dtype=torch.int32, device=DEV)
    scale_boxes_batched(det, cnt, pb.desc)
    out = det.cpu().numpy()[0]
    np.testing.assert_array_equal(out[:-3, :4], g[name + "_boxes"][:-3])
    np.testing.assert_array_equal(out[-3:, :4], boxes[-3:])  # rows >= count untouched
    assert (out[:, 4] == 0.5).all()
    b2 = torch.from_numpy(boxes).to(DEV)
    r = scale_boxes(got.shape[1:], b2, (h, w))
    assert r is b2
    np.testing.assert_array_equal(b2.cpu().numpy(), g[name + "_boxes"])


def test_ragged_batch_one_launch(gold):
    g = gold("preprocess.npz")
    imgs = [cases.preprocess_image(*cases.PREPROCESS_CASES[n][:2], cases.PREPROCESS_CASES[n][4]) for n in cases.PREPROCESS_BATCH]
    pre = DevicePreprocessor(160, device=DEV)
    before = ops.LAUNCHES
    out = torch.full((len(imgs) + 2, 3, 160, 160), 7, dtype=torch.uint8, device=DEV)
    pb = pre(imgs, out=out)
    assert ops.LAUNCHES - before == 1
    got = pb.im.cpu().numpy()
    for i, n in enumerate(cases.PREPROCESS_BATCH):
        np.testing.assert_array_equal(got[i], g[n])
    assert (out[len(imgs):] == 7).all()  # rows beyond the batch are not written
    # pinned staging views: the decoder-writes-into-pinned-memory path gives the same bytes; the two staging sets alternate
    for _ in range(3):
        views = pre.staging_views([im.shape[:2] for im in imgs])
        for v, im in zip(views, imgs):
            v[...] = im
        np.testing.assert_array_equal(pre(views).im.cpu().numpy(), got)
    with pytest.raises(TypeError):
        pre([imgs[0].astype(np.float32)])


def test_random_size_sweep_crc(gold):
    g = gold("preprocess.npz")
    pre = DevicePreprocessor(96, device=DEV)
    imgs = [cases.preprocess_image(h, w, seed) for h, w, seed in cases.PREPROCESS_SWEEP]
    got = pre(imgs).im.cpu().numpy()
    for i, crc in enumerate(g["sweep_crc"]):
        assert zlib.crc32(got[i].tobytes()) == int(crc), cases.PREPROCESS_SWEEP[i]


def test_full_size_batch_properties():
    """BASELINE size: 64 images -> 640 x 640.  (1) a 640 x 640 source is only re-ordered (BGR -> RGB, HWC -> CHW): identity property; (2) 720p and
    odd-sized sources equal the oracle byte for byte; (3) every pixel outside the resized region is 114."""
    rs = np.random.RandomState(5)
    shapes = [(640, 640), (720, 1280), (1080, 1920), (480, 640), (333, 517), (1280, 720)]
    imgs = [cases.preprocess_image(*shapes[i % len(shapes)], 300 + i) for i in range(64)]
    pre = DevicePreprocessor(640, device=DEV)
    pb = pre(imgs)
    got = pb.im.cpu().numpy()
    assert got.shape == (64, 3, 640, 640)
    for i in range(0, 64, len(shapes)):
        np.testing.assert_array_equal(got[i], imgs[i][..., ::-1].transpose(2, 0, 1))
    for i in (1, 2, 3, 4, 5, 61):
        np.testing.assert_array_equal(got[i], op.preprocess([imgs[i]], (640, 640))[0])
    for i in range(64):
        (nw, nh), top, bottom, left, right = op.letterbox_geometry(imgs[i].shape[:2], (640, 640))
        m = np.ones((640, 640), bool)
        m[top:top + nh, left:left + nw] = False
        assert (got[i][:, m] == 114).all()
    del rs


def test_detect_images_end_to_end(state_dict):
    """RefineEngine.detect_images (device LetterBox -> forward -> decode -> NMS -> scale_boxes) equals the engine run on the ORACLE's letterboxed
    tensor followed by the oracle's scale_boxes: same rows, same classes; boxes and scores to 1e-4 (two forward passes of the engine are not
    bit-identical run to run: the GroupNorm statistics are accumulated with floating-point atomics)"""
    from yolo_ad_refine_b200.engine import RefineEngine
    nms = dict(conf_thres=0.001, iou_thres=0.7, max_det=50)
    shapes = [(120, 200), (160, 160), (333, 250)]
    imgs = [cases.preprocess_image(h, w, 400 + i) for i, (h, w) in enumerate(shapes)]
    eng = RefineEngine(state_dict, batch=3, imgsz=160, dtype=torch.float32, device=DEV, input_u8=True, nms_args=nms)
    got = [d.clone() for d in eng.detect_images(imgs)]
    ref_in = op.preprocess(imgs, (160, 160))
    with pytest.raises(ValueError):
        eng.detect_images(imgs[:2])
    ref = eng.detect(torch.from_numpy(ref_in))
    assert len(got) == 3 and sum(d.shape[0] for d in got) > 0
    for d, r, (h, w) in zip(got, ref, shapes):
        r = r.cpu().numpy().copy()
        r[:, :4] = op.scale_boxes((160, 160), r[:, :4], (h, w))
        d = d.cpu().numpy()
        assert d.shape == r.shape
        np.testing.assert_array_equal(d[:, 5], r[:, 5])
        np.testing.assert_allclose(d[:, :5], r[:, :5], rtol=1e-4, atol=2e-3)
        assert (d[:, 0] >= 0).all() and (d[:, 2] <= w).all() and (d[:, 1] >= 0).all() and (d[:, 3] <= h).all()
