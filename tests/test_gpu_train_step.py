"""GPU parity of the full training step through libyad.so: train()-mode forward, detection loss, complete
backward, BatchNorm buffer updates and the optimizer, against (i) the fixtures written by the live reference (tests/golden/train_step.npz,
opt_step.npz: BaseModel.loss(...).backward(), clip_grad_norm_ + SGD + ModelEMA) and (ii) the CPU oracle on the same seeded inputs.

Tolerances: fp32 kernels -- loss 1e-4 relative, every parameter gradient within 1e-2 of its norm / 2e-2 per sampled element (the fixture's own
bar in tests/test_oracle_train.py); bf16 -- the storage type changes the TaskAlignedAssigner's discrete choices on random-init weights, so the
step is checked for consistency (loss within 10 %, gradient direction) and the bf16 kernels individually in tests/test_gpu_backward_ops.py."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import cases
from yolo_ad_refine_b200.trainer import TrainEngine

pytestmark = pytest.mark.gpu


def _inputs(name):
    img, bi, cl, bb = cases.train_step_inputs(**cases.TRAIN_STEP_CASES[name])
    return torch.from_numpy(img).cuda(), torch.from_numpy(bi), torch.from_numpy(cl), torch.from_numpy(bb)


@pytest.mark.parametrize("name", list(cases.TRAIN_STEP_CASES))
def test_train_step_fp32_matches_reference(gold, state_dict, name):
    g = gold("train_step.npz")
    eng = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1)
    out4 = eng.forward_backward(*_inputs(name)).cpu().numpy()
    assert abs(out4[3] - float(g[f"{name}_loss"])) < 1e-4 * abs(float(g[f"{name}_loss"]))
    np.testing.assert_allclose(out4[:3], g[f"{name}_items"], rtol=1e-4)
    tp = eng.tp
    checked = 0
    for k in tp.keys:
        v = tp.g(k).cpu().numpy().reshape(-1)
        if f"{name}|{k}|none" in g.files:
            assert float(np.abs(v).max()) == 0.0, k
            continue
        ref_norm = float(g[f"{name}|{k}|norm"])
        if k.endswith((".conv.bias", ".conv1.bias")) and ref_norm < 1e-3:
            continue  # a bias in front of a batch-statistics BatchNorm: the exact gradient is 0, both sides hold rounding noise
        norm = float(np.sqrt((v.astype(np.float64) ** 2).sum()))
        assert abs(norm - ref_norm) <= 1e-2 * ref_norm + 1e-4, (k, norm, ref_norm)  # absolute term: see tests/test_gpu_reference_api.py
        ref = g[f"{name}|{k}|samples"]
        np.testing.assert_allclose(v[cases.sample_positions(v.size, 16)], ref, rtol=2e-2, atol=1e-2 * ref_norm / np.sqrt(v.size) + 1e-7, err_msg=k)
        checked += 1
    assert checked > 300
    nb = 0
    for k in tp.buf_keys:
        np.testing.assert_allclose(tp.buf(k).cpu().numpy()[:8], g[f"{name}|{k}|buf"], rtol=1e-4, atol=1e-6, err_msg=k)
        nb += 1
    assert nb > 50


def test_two_optimizer_steps_fp32_match_reference(gold, state_dict):
    """forward + backward + clip_grad_norm_(10) + SGD(nesterov, 3 groups) + EMA, twice, against the live reference's parameter / EMA deltas"""
    g = gold("opt_step.npz")
    eng = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1, lr=0.01, momentum=0.937, weight_decay=5e-4)
    inp = _inputs("b2_160")
    for step in range(2):
        out4 = eng.step(*inp).cpu().numpy()
        assert abs(out4[3] - float(g[f"loss{step}"])) < 2e-4 * abs(float(g[f"loss{step}"]))
        assert abs(float(eng.tp.norm_sq.sqrt()) - float(g[f"gradnorm{step}"])) < 1e-3 * float(g[f"gradnorm{step}"])
    sd, ema = eng.tp.state_dict(), eng.tp.state_dict(ema=True)
    checked = 0
    for k, v0 in state_dict.items():
        if not v0.dtype.is_floating_point:
            continue
        pos = cases.sample_positions(v0.numel(), 16)
        dn = float(g[f"{k}|delta_norm"])
        atol = 2e-2 * dn / np.sqrt(v0.numel()) + 2e-7 * (float(v0.abs().max()) + 1e-3)  # deltas are quantised by the fp32 ulp of the parameter
        np.testing.assert_allclose((sd[k].cpu().double() - v0.double()).numpy().reshape(-1)[pos], g[f"{k}|delta"], rtol=2e-2, atol=atol, err_msg=k)
        np.testing.assert_allclose((ema[k].cpu().double() - v0.double()).numpy().reshape(-1)[pos], g[f"{k}|ema_delta"], rtol=2e-2, atol=atol,
                                   err_msg=k)
        checked += 1
    assert checked > 400


def test_train_step_bf16_consistent_with_fp32(state_dict):
    from oracle import synth
    img = torch.from_numpy(synth.make_images(4, 320, 320, seed=5)).cuda()
    bi, cl, bb = [torch.from_numpy(a) for a in synth.make_targets(4, seed=6, max_per_img=8)]
    res = {}
    for dtype, impl in ((torch.float32, 1), (torch.bfloat16, 0)):
        eng = TrainEngine(state_dict, dtype=dtype, conv_impl=impl)
        out4 = eng.forward_backward(img, bi, cl, bb).cpu().numpy()
        res[dtype] = (out4, eng.tp.grad.clone())
    (a4, ga), (b4, gb) = res[torch.float32], res[torch.bfloat16]
    assert np.isfinite(b4).all() and bool(torch.isfinite(gb).all())
    assert abs(b4[3] - a4[3]) < 0.1 * abs(a4[3])
    cos = float((ga * gb).sum() / (ga.norm() * gb.norm()))
    assert cos > 0.5, cos


def _grad_report(ta, tb):
    """whole-arena and per-tensor gradient cosines of two TrainParams (tensors that carry 99 % of the squared norm = "heavy")"""
    ga, gb = ta.grad, tb.grad
    cos_all = float((ga * gb).sum() / (ga.norm() * gb.norm()))
    per = []
    for k in ta.keys:
        x, y = ta.g(k).flatten().double(), tb.g(k).flatten().double()
        nx, ny = float(x.norm()), float(y.norm())
        if nx > 0 and ny > 0:
            per.append((k, float((x * y).sum() / (nx * ny)), nx))
    per.sort(key=lambda t: -t[2])
    tot = sum(n * n for _, _, n in per)
    acc, heavy = 0.0, []
    for k, c, n in per:
        if acc >= 0.99 * tot:
            break
        heavy.append((k, c))
        acc += n * n
    by_layer = {}
    for k, c, n in per:
        by_layer.setdefault(int(k.split(".")[1]), []).append((c, n * n))
    layer_cos = {L: sum(c * w for c, w in v) / sum(w for _, w in v) for L, v in sorted(by_layer.items())}
    return dict(cos_all=cos_all, tensors=len(per), heavy_tensors=len(heavy), heavy_min_cos=min(c for _, c in heavy),
                median_cos=float(np.median([c for _, c, _ in per])), frac_ge_099=float(np.mean([c >= 0.99 for _, c, _ in per])),
                layer_cos={str(k): round(v, 4) for k, v in layer_cos.items()})


def test_train_step_bf16_frozen_assignment(state_dict):
    """The bf16 step bench.py times (tcgen05 forward / dgrad / wgrad kernels), with the discrete part of the loss frozen: every step below uses the
    fp32 step's TaskAlignedAssigner outputs (fg_mask, target_gt_idx, target boxes / scores -- no-grad quantities in the reference, utils/tal.py:38).
      (i)  KERNEL check: against the bf16-storage SIMT step of the same library (conv_impl=1: the same storage type, fp32 FMA instead of tensor
           cores) -- loss within 0.1 %; the head's parameters (layer 33, where the loss gradient has passed through the kernels but not yet
           through a normalisation backward) at cosine >= 0.999; whole gradient >= 0.95, every heavy tensor >= 0.9;
      (ii) PRECISION statement: against the fp32 step that is pinned to the live reference -- loss within 0.1 %, whole-gradient cosine >= 0.85.
    Measured (gpurun_out/train_bf16_parity.json): the two bf16 builds differ from EACH OTHER (0.974) almost as much as from fp32 (0.93): below the
    head the gradient is a small difference of large terms (GroupNorm / BatchNorm backward subtract two means), so the 2^-9 rounding of a bf16
    activation gradient is amplified -- storage-type noise on random-init weights, not a kernel defect; the fp32 build is the one compared
    element-wise with the reference."""
    from oracle import synth
    img = torch.from_numpy(synth.make_images(4, 320, 320, seed=5)).cuda()
    bi, cl, bb = [torch.from_numpy(a) for a in synth.make_targets(4, seed=6, max_per_img=8)]
    e32 = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1)
    a4 = e32.forward_backward(img, bi, cl, bb, keep=True).cpu().numpy()
    aux = e32.last["aux"]
    e16 = TrainEngine(state_dict, dtype=torch.bfloat16, conv_impl=0)
    b4 = e16.forward_backward(img, bi, cl, bb, assign=aux).cpu().numpy()
    e16s = TrainEngine(state_dict, dtype=torch.bfloat16, conv_impl=1)
    c4 = e16s.forward_backward(img, bi, cl, bb, assign=aux).cpu().numpy()
    assert np.isfinite(b4).all() and bool(torch.isfinite(e16.tp.grad).all())
    rep = {"loss_fp32": [float(v) for v in a4], "loss_bf16_tc": [float(v) for v in b4], "loss_bf16_simt": [float(v) for v in c4],
           "tc_vs_simt_bf16": _grad_report(e16s.tp, e16.tp), "tc_bf16_vs_fp32": _grad_report(e32.tp, e16.tp),
           "simt_bf16_vs_fp32": _grad_report(e32.tp, e16s.tp)}
    out = os.path.join(os.path.dirname(GOLD), "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(rep, open(os.path.join(out, "train_bf16_parity.json"), "w"), indent=1)
    assert abs(b4[3] - c4[3]) < 1e-3 * abs(c4[3]) and abs(b4[3] - a4[3]) < 1e-3 * abs(a4[3]), rep
    k = rep["tc_vs_simt_bf16"]
    assert k["layer_cos"]["33"] >= 0.999 and k["cos_all"] >= 0.95 and k["heavy_min_cos"] >= 0.9, rep
    assert rep["tc_bf16_vs_fp32"]["cos_all"] >= 0.85 and rep["tc_bf16_vs_fp32"]["layer_cos"]["33"] >= 0.999, rep


def test_training_is_deterministic_in_shape_and_repeatable(state_dict):
    """two identical steps from identical state give the same loss (no stale gradient / buffer state leaks between steps)"""
    inp = _inputs("b2_160")
    eng = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1)
    a = eng.forward_backward(*inp, update_bn=False).cpu().numpy()
    ga = eng.tp.grad.clone()
    b = eng.forward_backward(*inp, update_bn=False).cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-4)  # statistics and loss sums use floating-point atomics: the summation order varies run to run
    assert float((eng.tp.grad - ga).norm() / ga.norm()) < 1e-3  # atomics reorder fp32 sums, nothing else may differ


def test_uint8_images_match_float_images(state_dict):
    """uint8 batches (what the reference's dataloader delivers; /255 on the device, models/yolo/detect/train.py:57-59) give the step of the
    equivalent float batch"""
    rs = np.random.RandomState(3)
    u8 = torch.from_numpy(rs.randint(0, 256, (2, 3, 160, 160), dtype=np.uint8)).cuda()
    from oracle import synth
    bi, cl, bb = [torch.from_numpy(a) for a in synth.make_targets(2, seed=6, max_per_img=5, empty_images=())]
    eng = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1)
    a = eng.forward_backward(u8, bi, cl, bb, update_bn=False).cpu().numpy()
    ga = eng.tp.grad.clone()
    b = eng.forward_backward(u8.float() / 255.0, bi, cl, bb, update_bn=False).cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-4)
    assert float((eng.tp.grad - ga).norm() / ga.norm()) < 1e-3


def test_train_step_without_targets_matches_oracle(state_dict):
    """a batch with no ground-truth boxes at all (utils/loss.py:392-408 -> empty targets, tal.py:62-70): only the classification term is alive"""
    from oracle import model as om
    img, _, _, _ = cases.train_step_inputs(**cases.TRAIN_STEP_CASES["b2_160"])
    bi, cl, bb = torch.zeros(0), torch.zeros(0, 1), torch.zeros(0, 4)
    loss_ref, items_ref, grads_ref, _, _ = om.train_step_grads(state_dict, torch.from_numpy(img), bi, cl, bb)
    eng = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1)
    out4 = eng.forward_backward(torch.from_numpy(img).cuda(), bi, cl, bb).cpu().numpy()
    assert np.isfinite(out4).all()
    assert abs(out4[3] - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
    assert out4[0] == 0.0 and out4[2] == 0.0
    for k in ("model.33.cv3.weight", "model.2.cv1.conv.weight", "model.10.m.0.attn.to_out.0.weight"):
        gr, rf = eng.tp.g(k).cpu(), grads_ref[k]
        assert float((gr - rf).norm() / rf.norm()) < 1e-2, k
    assert float(eng.tp.g("model.33.cv2.weight").abs().max()) == 0.0  # the box branch receives no gradient


def test_backward_is_the_derivative_of_the_forward(state_dict):
    """Size-independent property (no oracle): for a fixed random cotangent R on the three raw head outputs, the whole-model backward must give
    the derivative of S(w) = sum(outputs(w) * R): along a random parameter direction d, (S(w + e d) - S(w - e d)) / (2 e) == <grad, d>.
    (The detection loss itself is not used here: its targets are assignment-dependent constants for autograd, so a finite difference of the
    loss is not its gradient; the loss kernels have their own parity tests.)  fp32 build, 320 x 320, batch 4, batch-statistics BatchNorm."""
    from oracle import synth
    from yolo_ad_refine_b200 import training as T
    img = torch.from_numpy(synth.make_images(4, 320, 320, seed=9)).cuda()
    eng = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1)
    tp = eng.tp
    gen = torch.Generator(device="cuda").manual_seed(1)
    cot = None

    def run(backward):
        nonlocal cot
        tp.zero_grad()
        tp.pack()
        g = T.Graph(tp, 1, update_bn=False)
        outs, _ = T.forward_model(g, img)
        if cot is None:
            cot = [torch.randn(o.torch().shape, generator=gen, device="cuda") for o in outs]
        s = sum(float((o.torch().double() * r.double()).sum()) for o, r in zip(outs, cot))
        if backward:
            for o, r in zip(outs, cot):
                g.mark(o)
                g.grad(o).torch().copy_(r)
            g.backward()
            tp.unpack_grads()
        return s

    run(True)
    grad = tp.grad.clone()
    w0 = tp.flat.clone()
    head = torch.zeros(tp.total, device="cuda")
    for k in tp.keys:
        if k.startswith("model.33."):
            head[tp.off[k]:tp.off[k] + int(np.prod(tp.shape[k], dtype=np.int64))] = 1.0
    # (direction mask, step, tolerance): the network is strongly non-linear in its early layers (33 layers, batch-statistics BatchNorm), so the
    # central difference converges slowly there (measured: 23.4k, 39.5k, 44.3k, 47.7k for e = 3e-3 .. 1e-4 against <grad, d> = 47.4k); fp32
    # rounding of S (and the run-to-run order of the statistics atomics) forbids smaller steps, hence the loose bound on the all-layer direction
    for mask, eps, tol in ((torch.ones_like(head), 3e-4, 0.12), (head, 1e-3, 0.04)):
        d = torch.randn(tp.total, generator=gen, device="cuda") * mask * (tp.group != 255) * tp.flat.abs().clamp_min(1e-3)  # relative perturbation
        vals = []
        for sgn in (1.0, -1.0):
            tp.flat.copy_(w0 + sgn * eps * d)
            vals.append(run(False))
        tp.flat.copy_(w0)
        fd = (vals[0] - vals[1]) / (2 * eps)
        an = float((grad.double() * d.double()).sum())
        assert abs(fd - an) < tol * abs(an), (fd, an)


def test_gradient_accumulation_adds_up(state_dict):
    """zero_grad=False: two forward / backward passes on two half batches leave the sum of their gradients (accumulate > 1 in the reference)"""
    from oracle import synth
    img = torch.from_numpy(synth.make_images(2, 160, 160, seed=12)).cuda()
    bi, cl, bb = [torch.from_numpy(a) for a in synth.make_targets(2, seed=13, max_per_img=4, empty_images=())]
    eng = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1)
    eng.forward_backward(img, bi, cl, bb, update_bn=False)
    g1 = eng.tp.grad.clone()
    eng.forward_backward(img, bi, cl, bb, update_bn=False, zero_grad=False)
    assert float((eng.tp.grad - 2 * g1).norm() / (2 * g1).norm()) < 1e-3


def test_full_size_training_reduces_the_loss(state_dict):
    """BASELINE.json configs[3] size (batch 128, 640 x 640, bf16): four optimizer steps on one synthetic batch -- every loss and gradient norm
    is finite, the loss goes down, the EMA follows, BatchNorm buffers move (size-independent sanity of the complete step at full size)."""
    from yolo_ad_refine_b200 import synth
    rs = np.random.RandomState(11)
    img = torch.from_numpy(rs.randint(0, 256, (128, 3, 640, 640), dtype=np.uint8)).cuda()
    bi, cl, bb = [torch.from_numpy(a).cuda() for a in synth.make_targets(128, seed=12, max_per_img=8, empty_images=())]
    eng = TrainEngine(state_dict, dtype=torch.bfloat16)
    buf0 = eng.tp.bufs.clone()
    losses, norms = [], []
    for _ in range(4):
        out4 = eng.step(img, bi, cl, bb)
        losses.append(float(out4[3]))
        norms.append(float(eng.tp.norm_sq.sqrt()))
    assert np.isfinite(losses).all() and np.isfinite(norms).all()
    assert all(b < a for a, b in zip(losses, losses[1:])) and losses[-1] < 0.9 * losses[0], losses  # measured: 2.05M -> 1.59M
    assert float((eng.tp.ema - eng.tp.flat).abs().max()) > 0 and float((eng.tp.bufs - buf0).abs().max()) > 0
    assert bool(torch.isfinite(eng.tp.flat).all())


def test_two_adamw_steps_fp32_match_reference(gold, state_dict):
    """the reference's other optimizer (`optimizer=auto` on short runs, engine/trainer.py:773-782): clip_grad_norm_(10) + AdamW, two steps"""
    g = gold("opt_step_adamw.npz")
    eng = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1, lr=1e-3, momentum=0.937, weight_decay=5e-4, optimizer="AdamW")
    inp = _inputs("b2_160")
    for step in range(2):
        out4 = eng.step(*inp).cpu().numpy()
        # Adam divides every gradient element by its own running magnitude: elements whose gradient is rounding noise move by +-lr in
        # implementation-dependent directions, so the second step's loss agrees to 2e-3 only (SGD: 2e-4)
        assert abs(out4[3] - float(g[f"loss{step}"])) < 2e-3 * abs(float(g[f"loss{step}"]))
        assert abs(float(eng.tp.norm_sq.sqrt()) - float(g[f"gradnorm{step}"])) < 1e-2 * float(g[f"gradnorm{step}"])
    sd = eng.tp.state_dict()
    checked = 0
    all_got, all_ref = [], []
    for k, v0 in state_dict.items():
        if not v0.dtype.is_floating_point or k.endswith(("running_mean", "running_var")):
            continue
        if k.endswith((".conv.bias", ".conv1.bias", ".temps")):
            continue  # exact gradient 0 (a bias in front of a batch-statistics BatchNorm; the TSSA temperatures): Adam amplifies rounding noise
        pos = cases.sample_positions(v0.numel(), 16)
        got, ref = (sd[k].cpu().double() - v0.double()).numpy().reshape(-1)[pos], g[f"{k}|delta"]
        # per-parameter aggregate instead of element-wise: single elements whose gradient hovers around zero legitimately move differently
        all_got.append(got)
        all_ref.append(ref)
        if np.linalg.norm(ref) > 1e-5:
            assert np.linalg.norm(got - ref) <= 0.35 * np.linalg.norm(ref), (k, got, ref)
        checked += 1
    assert checked > 300
    a, b = np.concatenate(all_got), np.concatenate(all_ref)
    assert float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))) > 0.99
    assert np.linalg.norm(a - b) < 0.08 * np.linalg.norm(b)


def _sync_state(dst, src):
    """copy the training state of `src` into the (graph-captured, so in place) arenas of `dst`"""
    for name in ("flat", "mom", "ema", "bufs", "ema_bufs"):
        getattr(dst.tp, name).copy_(getattr(src.tp, name))
    if hasattr(src.tp, "mom2") and hasattr(dst.tp, "mom2"):
        dst.tp.mom2.copy_(src.tp.mom2)


@pytest.mark.parametrize("optimizer", ["SGD", "AdamW"])
def test_graphed_step_equals_eager_step(state_dict, optimizer):
    """TrainEngine.capture / step_graphed (one CUDA graph: forward + loss + backward + clip / optimizer / EMA with device-resident scalars, padded
    static targets) against the eager step, over three steps with a changing learning rate.  Every step starts from the SAME state in all three
    engines (eager A, graphed B, eager C: the state of A is copied into B and C after each step), so each comparison is a one-step comparison:
    loss to 1e-5, parameter / EMA update to 1e-3 of its norm (SGD) -- the summation order of the fp atomics is the only difference -- or to twice the
    spread between the two eager runs.  (Letting the three trajectories run free compared chaotic quantities: two EAGER runs of this random-init
    model drift apart by 0.05 % .. 1 % of the loss within three steps, and the graphed run failed a 3x-the-eager-spread bound about one time in
    ten.)  A stale learning rate or a missing update in the graph moves the parameters by the whole update, orders of magnitude above the bounds."""
    from oracle import synth
    rs = np.random.RandomState(3)
    img = torch.from_numpy(rs.randint(0, 256, (4, 3, 160, 160), dtype=np.uint8)).cuda()
    bi, cl, bb = [torch.from_numpy(a) for a in synth.make_targets(4, seed=6, max_per_img=6)]
    lrs = [0.002, 0.006, 0.01]
    ea, ec = (TrainEngine(state_dict, dtype=torch.float32, conv_impl=1, optimizer=optimizer) for _ in range(2))
    eb = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1, optimizer=optimizer)
    eb.capture(4, 160, n_max=8)
    f0 = state_dict_flat(ea, state_dict)
    # AdamW's step is lr * m / sqrt(v) element-wise (lr * sign(g) on the first step): gradient elements at the noise floor of the fp atomics flip
    # their whole update, hence the looser bound (measured 1.5 - 2.6 % of the update norm between two runs, eager or graphed)
    tol = 1e-3 if optimizer == "SGD" else 5e-2
    updates = []
    for i, lr in enumerate(lrs):
        start, ema_start = ea.tp.flat.clone(), ea.tp.ema.clone()
        la = ea.step(img, bi, cl, bb, lr=lr).cpu().numpy()
        lb = eb.step_graphed(img, bi, cl, bb, lr=lr).cpu().numpy()
        lc = ec.step(img, bi, cl, bb, lr=lr).cpu().numpy()
        assert np.isfinite(lb).all()
        np.testing.assert_allclose(la, lb, rtol=1e-5 if i == 0 else 1e-4)
        upd = float((ea.tp.flat - start).norm())
        spread = float((ea.tp.flat - ec.tp.flat).norm())  # the same quantity between the two EAGER runs
        assert upd > 0 and float((ea.tp.flat - eb.tp.flat).norm()) < max(tol * upd, 2 * spread), (i, upd, spread)
        assert float((ea.tp.ema - eb.tp.ema).norm()) < max(tol, 2 * spread / upd) * float((ea.tp.ema - ema_start).norm() + 1e-12)
        np.testing.assert_allclose(ea.tp.bufs.cpu().numpy(), eb.tp.bufs.cpu().numpy(), rtol=1e-4, atol=1e-6)
        updates.append(upd)
        _sync_state(eb, ea)
        _sync_state(ec, ea)
    assert min(updates) > 0 and float((ea.tp.flat - f0).norm()) > 0
    assert ea.tp.steps == eb.tp.steps == 3 and ea.tp.ema_updates == eb.tp.ema_updates
    hy = eb.tp._hyper.cpu().numpy()
    assert abs(hy[0] - lrs[-1]) < 1e-7 and abs(hy[11] - 10.0) < 1e-6      # the device-resident scalars of the last step
    assert int(eb.tp.other["model.0.bn.num_batches_tracked"]) == 3


def state_dict_flat(eng, sd):
    flat = torch.zeros_like(eng.tp.flat)
    for k in eng.tp.keys:
        n = sd[k].numel()
        flat[eng.tp.off[k]:eng.tp.off[k] + n] = sd[k].float().reshape(-1).to(flat.device)
    return flat


def test_adamw_checkpoint_resume_matches_uninterrupted_run(state_dict):
    """checkpoint() / load_checkpoint() carry AdamW's second-moment arena, the step counters and the optimizer's name: a run resumed after two
    steps takes the same third step as the uninterrupted run"""
    inp = _inputs("b2_160")
    a = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1, optimizer="AdamW", lr=0.002)
    for _ in range(2):
        a.step(*inp)
    ck = a.tp.checkpoint()
    assert "mom2" in ck and ck["optimizer"] == "AdamW" and int(ck["other"]["model.0.bn.num_batches_tracked"]) == 2
    a.step(*inp)
    b = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1, optimizer="AdamW", lr=0.002)
    b.tp.load_checkpoint(ck)
    b.step(*inp)
    step3 = float((a.tp.flat - ck["flat"].cuda()).norm())
    assert step3 > 0 and float((a.tp.flat - b.tp.flat).norm()) < 2e-3 * step3
    c = TrainEngine(state_dict, dtype=torch.float32, conv_impl=1, optimizer="SGD")
    c.tp.optimizer_name = "SGD"
    with pytest.raises(AssertionError):
        c.tp.load_checkpoint(ck)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_train_step_with_a_class_count_that_is_not_a_multiple_of_8(dtype):
    """custom datasets (the normal use of the fork): nc = 3.  The head's level outputs carry pad8(nc) class channels (zero weight rows), yad_head_pack /
    yad_head_unpack hand the loss the first nc and give the padding a zero gradient; against the CPU oracle's autograd on the same seeded model"""
    from oracle import model as om
    from oracle import synth
    nc = 3
    spec = [[k, ([nc] + list(sh[1:]) if k in ("model.33.cv3.weight", "model.33.cv3.bias") else sh), dt] for k, sh, dt in synth.load_spec()]
    sd = synth.make_state_dict(seed=5, spec=spec)
    img, bi, cl, bb = cases.train_step_inputs(**cases.TRAIN_STEP_CASES["b2_160"])
    cl = (cl % nc).astype(cl.dtype)
    img_t, bi_t, cl_t, bb_t = torch.from_numpy(img), torch.from_numpy(bi), torch.from_numpy(cl), torch.from_numpy(bb)
    loss_ref, items_ref, grads_ref, _, _ = om.train_step_grads(sd, img_t, bi_t, cl_t, bb_t)
    eng = TrainEngine(sd, dtype=dtype, conv_impl=1 if dtype == torch.float32 else 0, nc=nc)
    out4 = eng.forward_backward(img_t.cuda(), bi_t, cl_t, bb_t).cpu().numpy()
    assert np.isfinite(out4).all()
    if dtype == torch.float32:
        assert abs(out4[3] - float(loss_ref)) < 1e-4 * abs(float(loss_ref)), (out4, float(loss_ref))
        np.testing.assert_allclose(out4[:3], items_ref.numpy(), rtol=1e-4)
        for k in ("model.33.cv3.weight", "model.33.cv3.bias", "model.33.cv2.weight", "model.33.rep_block_cls.conv2.conv.weight", "model.2.cv1.conv.weight",
                  "model.10.m.0.attn.to_out.0.weight"):
            gr, rf = eng.tp.g(k).cpu(), grads_ref[k]
            assert gr.shape == rf.shape and float((gr - rf).norm() / rf.norm()) < 1e-2, k
    else:  # bf16 storage changes the assigner's discrete choices on random-init weights (see the module docstring): consistency only
        # (measured over three runs, tools/nc3_margin.py: loss 5.7 - 6.9 % off the fp32 oracle, cv3 gradient cosine 0.9999)
        assert abs(out4[3] - float(loss_ref)) < 0.20 * abs(float(loss_ref)), (out4, float(loss_ref))
        gr, rf = eng.tp.g("model.33.cv3.weight").cpu(), grads_ref["model.33.cv3.weight"]
        assert gr.shape == rf.shape and float(torch.nn.functional.cosine_similarity(gr.reshape(-1), rf.reshape(-1), dim=0)) > 0.9
