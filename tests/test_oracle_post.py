"""CPU: oracle NMS / decode restatement (oracle/postprocess.py, oracle/model.decode) against fixtures produced by the reference's
ops.non_max_suppression + torchvision.ops.nms (oracle/gen_golden.py).  Integer/ordering outputs must match exactly."""
import numpy as np
import pytest
import torch

from oracle import cases, synth
from oracle import model as om
from oracle import postprocess as op


def _split(rows, counts):
    out, o = [], 0
    for c in counts:
        out.append(rows[o:o + c]); o += c
    return out


@pytest.mark.parametrize("name", list(cases.NMS_CASES))
def test_nms_matches_reference(gold, name):
    g = gold("nms_cases.npz")
    pk, nk = cases.NMS_CASES[name]
    pred = synth.make_predictions(**pk)
    out = op.non_max_suppression(pred, **nk)
    counts = g[f"{name}_count"]
    assert [o.shape[0] for o in out] == list(counts)
    for got, ref in zip(out, _split(g[f"{name}_rows"], counts)):
        np.testing.assert_array_equal(got, ref)  # bit-exact rows in the same order => identical keep indices


def test_decode_then_nms_matches_reference(gold):
    g = gold("nms_cases.npz")
    raw = synth.make_head_logits(4, 8400, seed=0)
    y = om.decode(torch.from_numpy(raw), [(80, 80), (40, 40), (20, 20)]).numpy()
    np.testing.assert_allclose(y[:, :, ::11], g["decode_y_sub"], rtol=1e-5, atol=1e-5)
    out = op.non_max_suppression(y, conf_thres=0.25, iou_thres=0.7, max_det=300)
    counts = g["decode_count"]
    assert [o.shape[0] for o in out] == list(counts)
    for got, ref in zip(out, _split(g["decode_rows"], counts)):
        # the class/score columns and order must agree exactly; boxes to fp32 rounding of the decode
        np.testing.assert_array_equal(got[:, 5], ref[:, 5])
        np.testing.assert_allclose(got[:, :5], ref[:, :5], rtol=1e-5, atol=1e-4)
