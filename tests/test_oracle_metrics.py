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


def _run_val_oracle(v):
    """update_metrics -> get_stats -> ap_per_class through the oracle"""
    tp, conf, pcls, tcls, labels = [], [], [], [], []
    for si, det in enumerate(v["dets"]):
        sel = v["batch_idx"] == si
        (gain, _), pad = v["ratio_pad"][si]
        bbox = om.prepare_labels(v["bboxes"][sel], (v["imgsz"], v["imgsz"]), v["ori_shape"][si], gain, pad)
        cls = v["cls"][sel]
        labels.append(bbox)
        if len(det) == 0:
            tcls.append(cls)
            continue
        predn = det.copy()
        predn[:, :4] = om.clip_boxes((det[:, :4] - np.array([pad[0], pad[1], pad[0], pad[1]], np.float32)) / np.float32(gain), v["ori_shape"][si])
        tp.append(om.process_batch(predn, bbox, cls)), conf.append(predn[:, 4]), pcls.append(predn[:, 5]), tcls.append(cls)
    return np.concatenate(labels), np.concatenate(tp), np.concatenate(conf), np.concatenate(pcls), np.concatenate(tcls)


@pytest.mark.parametrize("name", list(cases.VAL_CASES))
def test_validation_statistics_match_reference(gold, name):
    g = gold("val_cases.npz")
    v = cases.val_inputs(*cases.VAL_CASES[name])
    labels, tp, conf, pcls, tcls = _run_val_oracle(v)
    np.testing.assert_array_equal(labels, g[name + "_labels"])
    n = int(g[name + "_n"])
    assert tp.shape == (n, 10)
    np.testing.assert_array_equal(tp, np.unpackbits(g[name + "_tp"], axis=0)[:n].astype(bool))
    r = om.ap_per_class(tp, conf, pcls, tcls)
    np.testing.assert_array_equal(r["unique_classes"], g[name + "_classes"])
    assert r["f1_index"] == int(g[name + "_f1_index"])
    for k, gk in (("ap", "ap"), ("p", "p"), ("r", "r"), ("f1", "f1"), ("tp", "tpn"), ("fp", "fpn")):
        np.testing.assert_allclose(r[k], g[f"{name}_{gk}"], rtol=1e-12, atol=1e-15, err_msg=k)
    for k in ("p_curve", "r_curve", "f1_curve"):
        np.testing.assert_allclose(r[k][:, ::8], g[f"{name}_{k}"], rtol=1e-12, atol=1e-15, err_msg=k)
