"""CPU: oracle/preprocess.py (LetterBox incl. the restated cv2.resize INTER_LINEAR fixed-point arithmetic, BGR->RGB / HWC->CHW, scale_boxes)
against the fixtures written by the live reference (oracle/gen_golden.py preprocess): bit-exact bytes, boxes exact in fp32."""
import numpy as np
import pytest

from oracle import cases
from oracle import preprocess as op


@pytest.mark.parametrize("name", list(cases.PREPROCESS_CASES))
def test_letterbox_matches_reference(gold, name):
    g = gold("preprocess.npz")
    h, w, size, auto, seed = cases.PREPROCESS_CASES[name]
    img = cases.preprocess_image(h, w, seed)
    got = op.preprocess([img], (size, size), auto=auto, stride=32)[0]
    assert got.shape == g[name].shape
    np.testing.assert_array_equal(got, g[name])
    boxes = cases.scale_boxes_inputs(seed + 100) * np.float32(size / 640.0)
    np.testing.assert_array_equal(op.scale_boxes(got.shape[1:], boxes, (h, w)), g[name + "_boxes"])


def test_random_size_sweep_crc(gold):
    """48 random source sizes -> 96 x 96: CRC32 of the oracle's bytes equals the CRC32 of the live reference's (cv2) bytes"""
    import zlib
    g = gold("preprocess.npz")
    for (h, w, seed), crc in zip(cases.PREPROCESS_SWEEP, g["sweep_crc"]):
        got = op.preprocess([cases.preprocess_image(h, w, seed)], (96, 96))[0]
        assert zlib.crc32(got.tobytes()) == int(crc), (h, w)


def test_host_geometry_matches_oracle():
    """the product's host-side LetterBox / scale_boxes scalar geometry (yolo_ad_refine_b200/preprocess.py) against the oracle's"""
    from yolo_ad_refine_b200.preprocess import letterbox_params, scale_boxes_params
    rs = np.random.RandomState(0)
    for _ in range(500):
        h, w = rs.randint(8, 2000, 2)
        size = int(rs.choice([64, 160, 320, 640, 1280]))
        auto = bool(rs.randint(2))
        (nw, nh), top, bottom, left, right = op.letterbox_geometry((h, w), (size, size), auto=auto, stride=32)
        assert letterbox_params((h, w), size, auto=auto, stride=32) == (nw, nh, top, bottom, left, right)
        H, W = nh + top + bottom, nw + left + right
        gain, pad = scale_boxes_params((H, W), (h, w))
        b = np.array([[10.0, 20.0, 30.0, 40.0]], np.float32)
        want = op.scale_boxes((H, W), b, (h, w))
        got = np.stack([np.clip((b[:, 0] - np.float32(pad[0])) / np.float32(gain), 0, w), np.clip((b[:, 1] - np.float32(pad[1])) / np.float32(gain), 0, h),
                        np.clip((b[:, 2] - np.float32(pad[0])) / np.float32(gain), 0, w), np.clip((b[:, 3] - np.float32(pad[1])) / np.float32(gain), 0, h)], 1)
        np.testing.assert_array_equal(got.astype(np.float32), want)
