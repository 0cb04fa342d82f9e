"""CPU: oracle/metrics.py (box_iou + the closed-form restatement of BaseValidator.match_predictions) against the fixtures written by the live
reference (oracle/gen_golden.py match): IoU matrices equal in fp32, correct-matrices bit-exact."""
import numpy as np
import pytest

from oracle import cases
from oracle import metrics as om


@pytest.mark.parametrize("name", list(cases.MATCH_CASES))
def test_match_predictions_matches_reference(gold, name):
    g = gold("match_cases.npz")
    det, gt, gt_cls = cases.match_inputs(*cases.MATCH_CASES[name])
    iou = om.box_iou(gt, det[:, :4])
    assert iou.shape == g[name + "_iou"].shape
    np.testing.assert_array_equal(iou, g[name + "_iou"])
    got = om.process_batch(det, gt, gt_cls)
    assert got.shape == g[name].shape
    np.testing.assert_array_equal(got, g[name])
