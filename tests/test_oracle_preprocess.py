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
