"""CPU: oracle TaskAlignedAssigner + detection loss (oracle/tal_loss.py) against fixtures produced by the reference's
utils/tal.py + utils/loss.py (oracle/gen_golden.py).  Index outputs exact; floats 1e-4 relative."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import tal_loss as ot


def test_assigner_matches_reference(gold):
    g = gold("tal_case.npz")
    args = [torch.from_numpy(a) for a in cases.tal_inputs()]
    tl, tb, ts, fg, tgi = ot.assign(*args)
    np.testing.assert_array_equal(fg.numpy().astype(np.uint8), g["fg_mask"])
    np.testing.assert_array_equal(tgi.numpy().astype(np.int32), g["target_gt_idx"])
    np.testing.assert_array_equal(tl.numpy().astype(np.int32), g["target_labels"])
    fgm = fg.numpy().astype(bool)
    np.testing.assert_allclose(tb.numpy()[fgm], g["target_bboxes_fg"], rtol=1e-6)
    np.testing.assert_allclose(ts.numpy().max(-1)[fgm], g["target_scores_fgmax"], rtol=1e-4, atol=1e-7)
    assert abs(float(ts.sum()) - float(g["target_scores_sum"])) < 1e-4 * float(g["target_scores_sum"])


@pytest.mark.parametrize("name", list(cases.TRAIN_CASES))
def test_loss_matches_reference(gold, name):
    g = gold("train_cases.npz")
    feats_np, bi, cl, bb = cases.train_inputs(**cases.TRAIN_CASES[name])
    feats = [torch.from_numpy(f).requires_grad_(True) for f in feats_np]
    loss, items, aux = ot.detection_loss(feats, torch.from_numpy(bi), torch.from_numpy(cl), torch.from_numpy(bb))
    np.testing.assert_array_equal(aux["fg_mask"].numpy().astype(np.uint8), g[f"{name}_fg_mask"])
    np.testing.assert_array_equal(aux["target_gt_idx"].numpy().astype(np.int32), g[f"{name}_target_gt_idx"])
    np.testing.assert_array_equal(aux["target_labels"].numpy().astype(np.int32), g[f"{name}_target_labels"])
    np.testing.assert_allclose(items.numpy(), g[f"{name}_items"], rtol=1e-4, atol=1e-6)
    assert abs(loss.item() - float(g[f"{name}_loss"])) < 1e-4 * abs(float(g[f"{name}_loss"]))
    loss.backward()
    for i, f in enumerate(feats):
        gr = f.grad.numpy().reshape(-1)
        ref = g[f"{name}_grad{i}_samples"]
        np.testing.assert_allclose(gr[cases.sample_positions(gr.size, 256)], ref, rtol=1e-3, atol=1e-6 * (np.abs(ref).max() + 1))
        assert abs(np.abs(gr).sum() - float(g[f"{name}_grad{i}_abssum"])) < 1e-3 * float(g[f"{name}_grad{i}_abssum"])
