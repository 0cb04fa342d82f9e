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


@pytest.mark.parametrize("name", list(cases.TRAIN_STEP_CASES))
def test_train_step_matches_reference(gold, state_dict, name):
    """oracle.model.train_step_grads (train-mode forward with batch-statistics BatchNorm + loss + autograd) against the fixture written
    by the live reference's BaseModel.loss(...).backward(): loss, every parameter gradient (norm + samples), BatchNorm buffer updates."""
    from oracle import model as om
    g = gold("train_step.npz")
    img, bi, cl, bb = cases.train_step_inputs(**cases.TRAIN_STEP_CASES[name])
    loss, items, grads, bn_upd, _ = om.train_step_grads(state_dict, torch.from_numpy(img), torch.from_numpy(bi), torch.from_numpy(cl),
                                                        torch.from_numpy(bb))
    assert abs(loss.item() - float(g[f"{name}_loss"])) < 1e-4 * abs(float(g[f"{name}_loss"]))
    np.testing.assert_allclose(items.numpy(), g[f"{name}_items"], rtol=1e-4)
    checked = 0
    for k, gr in grads.items():
        if f"{name}|{k}|none" in g.files:
            assert gr is None or float(gr.abs().max()) == 0.0, k
            continue
        ref_norm = float(g[f"{name}|{k}|norm"])
        v = gr.numpy().reshape(-1)
        if k.endswith((".conv.bias", ".conv1.bias")) and ref_norm < 1e-3:
            continue  # a bias in front of a batch-statistics BatchNorm: the exact gradient is 0, both sides hold rounding noise
        norm = float(np.sqrt((v.astype(np.float64) ** 2).sum()))
        assert abs(norm - ref_norm) <= 1e-2 * ref_norm + 1e-6, (k, norm, ref_norm)
        ref = g[f"{name}|{k}|samples"]
        np.testing.assert_allclose(v[cases.sample_positions(v.size, 16)], ref, rtol=2e-2, atol=1e-2 * ref_norm / np.sqrt(v.size) + 1e-7, err_msg=k)
        checked += 1
    assert checked > 300
    for k, v in bn_upd.items():
        np.testing.assert_allclose(v.numpy().reshape(-1)[:8], g[f"{name}|{k}|buf"], rtol=1e-4, atol=1e-6, err_msg=k)
    assert len(bn_upd) > 50


def test_train_step_with_three_classes_matches_reference(gold):
    """a custom-dataset class count (nc = 3, not a multiple of 8): oracle.model.train_step_grads against the live reference's
    DetectionModel(yaml, nc=3).loss(batch).backward() (tests/golden/train_step_nc3.npz, oracle/gen_golden.py train_step_nc3): loss, loss items and
    every parameter gradient.  The CUDA training head is compared with this oracle in tests/test_gpu_train_step.py."""
    from oracle import model as om
    from oracle import synth
    g = gold("train_step_nc3.npz")
    nc = 3
    spec = [[k, ([nc] + list(sh[1:]) if k in ("model.33.cv3.weight", "model.33.cv3.bias") else sh), dt] for k, sh, dt in synth.load_spec()]
    sd = synth.make_state_dict(seed=5, spec=spec)
    img, bi, cl, bb = cases.train_step_inputs(**cases.TRAIN_STEP_CASES["b2_160"])
    cl = (cl % nc).astype(cl.dtype)
    loss, items, grads, _, feats = om.train_step_grads(sd, torch.from_numpy(img), torch.from_numpy(bi), torch.from_numpy(cl), torch.from_numpy(bb))
    assert feats[0].shape[1] == 64 + nc
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    np.testing.assert_allclose(items.numpy(), g["items"], rtol=1e-4)
    checked = 0
    for k, gr in grads.items():
        if f"{k}|none" in g.files:
            assert gr is None or float(gr.abs().max()) == 0.0, k
            continue
        ref_norm = float(g[f"{k}|norm"])
        v = gr.numpy().reshape(-1)
        if k.endswith((".conv.bias", ".conv1.bias")) and ref_norm < 1e-3:
            continue  # a bias in front of a batch-statistics BatchNorm: the exact gradient is 0, both sides hold rounding noise
        norm = float(np.sqrt((v.astype(np.float64) ** 2).sum()))
        assert abs(norm - ref_norm) <= 1e-2 * ref_norm + 1e-6, (k, norm, ref_norm)
        ref = g[f"{k}|samples"]
        np.testing.assert_allclose(v[cases.sample_positions(v.size, 16)], ref, rtol=2e-2, atol=1e-2 * ref_norm / np.sqrt(v.size) + 1e-7, err_msg=k)
        checked += 1
    assert checked > 300
    assert tuple(grads["model.33.cv3.weight"].shape) == (nc, 64, 1, 1)


def test_optimizer_step_matches_reference(gold, state_dict):
    """oracle/optim.py (clip + SGD nesterov + EMA) driven by oracle.model.train_step_grads for two steps, against the live reference's two
    steps (tests/golden/opt_step.npz): sampled parameter / EMA deltas."""
    import json
    import os
    from conftest import GOLD
    from oracle import model as om
    from oracle import optim as oo
    g = gold("opt_step.npz")
    groups = json.load(open(os.path.join(GOLD, "optimizer_groups.json")))
    img, bi, cl, bb = [torch.from_numpy(a) for a in cases.train_step_inputs(**cases.TRAIN_STEP_CASES["b2_160"])]
    sd = {k: v.clone() for k, v in state_dict.items()}
    ema = {k: v.clone() for k, v in state_dict.items()}
    bufs = {}
    for step in range(2):
        loss, _, grads, bn_upd, _ = om.train_step_grads(sd, img, bi, cl, bb)
        assert abs(loss.item() - float(g[f"loss{step}"])) < 2e-4 * abs(float(g[f"loss{step}"]))
        params = {k: sd[k] for k in grads}
        norm = oo.sgd_step(params, grads, bufs, groups)
        assert abs(norm - float(g[f"gradnorm{step}"])) < 1e-3 * float(g[f"gradnorm{step}"])
        sd.update(bn_upd)
        oo.ema_update(ema, sd, step + 1)
    checked = 0
    for k, v in sd.items():
        if not v.dtype.is_floating_point:
            continue
        pos = cases.sample_positions(v.numel(), 16)
        dn = float(g[f"{k}|delta_norm"])
        atol = 2e-2 * dn / np.sqrt(v.numel()) + 2e-7 * (float(v.abs().max()) + 1e-3)  # deltas are quantised by the fp32 ulp of the parameter
        np.testing.assert_allclose((v.double() - state_dict[k].double()).numpy().reshape(-1)[pos], g[f"{k}|delta"], rtol=2e-2, atol=atol, err_msg=k)
        np.testing.assert_allclose((ema[k].double() - state_dict[k].double()).numpy().reshape(-1)[pos], g[f"{k}|ema_delta"], rtol=2e-2,
                                   atol=atol, err_msg=k)
        checked += 1
    assert checked > 400


def test_adamw_step_matches_reference(gold, state_dict):
    """oracle/optim.py adamw_step (clip + AdamW with the reference's groups) for two steps against the live reference
    (tests/golden/opt_step_adamw.npz: BaseTrainer.build_optimizer(name="AdamW", lr=1e-3))"""
    import json
    import os
    from conftest import GOLD
    from oracle import model as om
    from oracle import optim as oo
    g = gold("opt_step_adamw.npz")
    groups = json.load(open(os.path.join(GOLD, "optimizer_groups.json")))
    img, bi, cl, bb = [torch.from_numpy(a) for a in cases.train_step_inputs(**cases.TRAIN_STEP_CASES["b2_160"])]
    sd = {k: v.clone() for k, v in state_dict.items()}
    state = {}
    for step in range(2):
        loss, _, grads, bn_upd, _ = om.train_step_grads(sd, img, bi, cl, bb)
        assert abs(loss.item() - float(g[f"loss{step}"])) < 5e-4 * abs(float(g[f"loss{step}"]))
        oo.adamw_step({k: sd[k] for k in grads}, grads, state, groups, step + 1, lr=1e-3)
        sd.update(bn_upd)
    checked = 0
    for k, v in sd.items():
        if not v.dtype.is_floating_point or k.endswith(("running_mean", "running_var")):
            continue
        if k.endswith((".conv.bias", ".conv1.bias", ".temps")):
            continue  # exact gradient 0 (a bias in front of a batch-statistics BatchNorm; the TSSA temperatures): Adam amplifies rounding noise
        pos = cases.sample_positions(v.numel(), 16)
        dn = float(g[f"{k}|delta_norm"])
        atol = 5e-2 * dn / np.sqrt(v.numel()) + 2e-7 * (float(v.abs().max()) + 1e-3)
        np.testing.assert_allclose((v.double() - state_dict[k].double()).numpy().reshape(-1)[pos], g[f"{k}|delta"], rtol=5e-2, atol=atol, err_msg=k)
        checked += 1
    assert checked > 300
