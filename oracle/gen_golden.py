"""TEST INFRASTRUCTURE -- generates the committed golden fixtures in tests/golden/ by running the LIVE reference
(/root/reference, imported read-only through oracle/ref_shims.py) on deterministic synthetic weights and inputs
(oracle/synth.py).  Run in the build container only:   python -m oracle.gen_golden

The reference's own tests hold no golden vectors for this path, so these fixtures -- outputs of the
reference's own PyTorch code -- are what pins the oracle restatement (tests/test_oracle_*.py) and, through it, the CUDA
kernels.  Fixtures are kept small (sub-sampled activations, index outputs as int32) so that they can be committed.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shims, synth  # noqa: E402

ref_shims.install()
import torch  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
from ultralytics.nn.tasks import DetectionModel  # noqa: E402
from ultralytics.utils import ops as ref_ops  # noqa: E402
from ultralytics.utils.loss import v8DetectionLoss  # noqa: E402
from ultralytics.utils.tal import TaskAlignedAssigner  # noqa: E402

YAML = os.path.join(ref_shims.REFERENCE_ROOT, "z-yaml", "yolo11-701-YOLO-AD-Refine.yaml")
from oracle.cases import NMS_CASES, TRAIN_CASES, sample_positions, tal_inputs, train_inputs  # noqa: E402


def build_reference_model(fuse=True):
    m = DetectionModel(YAML, ch=3, nc=80, verbose=False)
    spec = [(k, list(v.shape), str(v.dtype)) for k, v in m.state_dict().items()]
    os.makedirs(GOLD, exist_ok=True)
    with open(synth.SPEC_PATH, "w") as f:
        json.dump(spec, f)
    sd = synth.make_state_dict(seed=1, spec=spec)
    m.load_state_dict(sd, strict=True)
    m.eval()
    if fuse:
        m.fuse(verbose=False)
    return m, spec


def run_model(m, img):
    outs = {}
    hooks = []
    for i, layer in enumerate(m.model):
        hooks.append(layer.register_forward_hook(lambda mod, inp, out, i=i: outs.__setitem__(i, out)))
    m.model[-1].shape = None  # anchors are cached by shape (head.py:1184)
    with torch.inference_mode():
        y, feats = m(torch.from_numpy(img))
    for h in hooks:
        h.remove()
    return y, feats, outs


def layer_fixture(outs):
    d = {}
    for i in range(33):
        o = outs[i].detach().float().numpy().reshape(-1)
        pos = sample_positions(o.size)
        d[f"L{i}_shape"] = np.asarray(outs[i].shape, np.int32)
        d[f"L{i}_samples"] = o[pos]
        d[f"L{i}_absmean"] = np.float64(np.abs(o).mean())
        d[f"L{i}_mean"] = np.float64(o.astype(np.float64).mean())
    return d


def gen_model(m):
    # --- 160x160, batch 2: everything kept
    img = synth.make_images(2, 160, 160, seed=2)
    y, feats, outs = run_model(m, img)
    d = layer_fixture(outs)
    d["y"] = y.numpy()
    for i, f in enumerate(feats):
        d[f"feat{i}"] = f.numpy()
    np.savez_compressed(os.path.join(GOLD, "model_160.npz"), **d)
    print("model_160: y", tuple(y.shape), "max cls", float(y[:, 4:].max()))

    # --- 640x640, batch 1: sub-sampled y + NMS outputs (BASELINE.json configs[0])
    img = synth.make_images(1, 640, 640, seed=2)
    y, feats, outs = run_model(m, img)
    d = layer_fixture(outs)
    d["y_sub"] = y.numpy()[:, :, ::7]
    d["y_absmean"] = np.float64(y.abs().mean())
    for name, kw in (("predict", dict(conf_thres=0.25, iou_thres=0.7, max_det=300)),
                     ("val", dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True))):
        out = ref_ops.non_max_suppression(y.clone(), max_time_img=1e9, **kw)
        d[f"nms_{name}"] = out[0].numpy()
        print("model_640 nms", name, tuple(out[0].shape))
    np.savez_compressed(os.path.join(GOLD, "model_640.npz"), **d)


def gen_nms():
    d = {}
    for name, (pk, nk) in NMS_CASES.items():
        pred = synth.make_predictions(**pk)
        out = ref_ops.non_max_suppression(torch.from_numpy(pred.copy()), max_time_img=1e9, **nk)
        d[f"{name}_count"] = np.asarray([o.shape[0] for o in out], np.int32)
        d[f"{name}_rows"] = np.concatenate([o.numpy() for o in out], 0) if out else np.zeros((0, 6), np.float32)
        print("nms", name, d[f"{name}_count"])
    # decode + NMS on raw head logits (reduced batch)
    raw = synth.make_head_logits(4, 8400, seed=0)
    from ultralytics.nn.modules.block import DFL
    from ultralytics.utils.tal import dist2bbox, make_anchors
    x = torch.from_numpy(raw)
    feats = [torch.zeros(1, 1, 80, 80), torch.zeros(1, 1, 40, 40), torch.zeros(1, 1, 20, 20)]
    anchors, strides = (t.transpose(0, 1) for t in make_anchors(feats, torch.tensor([8.0, 16.0, 32.0]), 0.5))
    dfl = DFL(16)
    box, cls = x.split((64, 80), 1)
    dbox = dist2bbox(dfl(box), anchors.unsqueeze(0), xywh=True, dim=1) * strides
    y = torch.cat((dbox, cls.sigmoid()), 1)
    d["decode_y_sub"] = y.numpy()[:, :, ::11]
    out = ref_ops.non_max_suppression(y.clone(), conf_thres=0.25, iou_thres=0.7, max_det=300, max_time_img=1e9)
    d["decode_count"] = np.asarray([o.shape[0] for o in out], np.int32)
    d["decode_rows"] = np.concatenate([o.numpy() for o in out], 0)
    print("decode+nms", d["decode_count"])
    np.savez_compressed(os.path.join(GOLD, "nms_cases.npz"), **d)


class _Args:
    box, cls, dfl = 7.5, 0.5, 1.5


def gen_train(m):
    m.args = _Args()
    crit = v8DetectionLoss(m)
    d = {}
    for name, kw in TRAIN_CASES.items():
        feats_np, bi, cl, bb = train_inputs(**kw)
        feats = [torch.from_numpy(f).requires_grad_(True) for f in feats_np]
        batch = dict(batch_idx=torch.from_numpy(bi), cls=torch.from_numpy(cl), bboxes=torch.from_numpy(bb))
        captured = {}
        orig = crit.assigner.forward

        def spy(*a, **k):
            out = orig(*a, **k)
            captured["in"] = [t.detach().clone() for t in a]
            captured["out"] = [t.detach().clone() for t in out]
            return out

        crit.assigner.forward = spy
        loss, items = crit(feats, batch)
        crit.assigner.forward = orig
        loss.backward()
        tl, tb, ts, fg, tgi = captured["out"]
        d[f"{name}_loss"] = np.float64(loss.item())
        d[f"{name}_items"] = items.numpy().astype(np.float64)
        d[f"{name}_target_labels"] = tl.numpy().astype(np.int32)
        d[f"{name}_fg_mask"] = fg.numpy().astype(np.uint8)
        d[f"{name}_target_gt_idx"] = tgi.numpy().astype(np.int32)
        fgm = fg.numpy().astype(bool)
        d[f"{name}_target_bboxes_fg"] = tb.numpy()[fgm]
        d[f"{name}_target_scores_fgmax"] = ts.numpy().max(-1)[fgm]
        d[f"{name}_target_scores_sum"] = np.float64(ts.sum().item())
        for i, f in enumerate(feats):
            g = f.grad.numpy().reshape(-1)
            pos = sample_positions(g.size, 256)
            d[f"{name}_grad{i}_samples"] = g[pos]
            d[f"{name}_grad{i}_abssum"] = np.float64(np.abs(g).sum())
        print("train", name, "loss", loss.item(), items.numpy(), "fg", int(fgm.sum()))
    np.savez_compressed(os.path.join(GOLD, "train_cases.npz"), **d)

    # stand-alone TaskAlignedAssigner call with padded gts (tal.py:38-88)
    pd_scores, pd_bboxes, anc_px, gt_labels, gt_bboxes, mask_gt = tal_inputs()
    nc = pd_scores.shape[-1]
    tal = TaskAlignedAssigner(topk=10, num_classes=nc, alpha=0.5, beta=6.0)
    out = tal(torch.from_numpy(pd_scores), torch.from_numpy(pd_bboxes), torch.from_numpy(anc_px), torch.from_numpy(gt_labels),
              torch.from_numpy(gt_bboxes), torch.from_numpy(mask_gt))
    tl, tb, ts, fg, tgi = out
    fgm = fg.numpy().astype(bool)
    np.savez_compressed(os.path.join(GOLD, "tal_case.npz"), target_labels=tl.numpy().astype(np.int32), fg_mask=fg.numpy().astype(np.uint8),
                        target_gt_idx=tgi.numpy().astype(np.int32), target_bboxes_fg=tb.numpy()[fgm],
                        target_scores_fgmax=ts.numpy().max(-1)[fgm], target_scores_sum=np.float64(ts.sum().item()))
    print("tal fg", int(fgm.sum()))


def gen_train_step():
    """One full training forward + backward of the UNFUSED reference model in train() mode (batch-statistics BatchNorm) through
    BaseModel.loss (nn/tasks.py:270-297): pins oracle.model.train_step_grads and, through it, the CUDA backward path."""
    from oracle.cases import TRAIN_STEP_CASES, train_step_inputs
    m, _ = build_reference_model(fuse=False)
    m.args = _Args()
    m.train()
    d = {}
    for name, kw in TRAIN_STEP_CASES.items():
        img, bi, cl, bb = train_step_inputs(**kw)
        sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
        m.zero_grad()
        m.criterion = None
        m.model[-1].shape = None
        batch = dict(img=torch.from_numpy(img), batch_idx=torch.from_numpy(bi), cls=torch.from_numpy(cl), bboxes=torch.from_numpy(bb))
        loss, items = m.loss(batch)
        loss.backward()
        d[f"{name}_loss"] = np.float64(loss.item())
        d[f"{name}_items"] = items.detach().numpy().astype(np.float64)
        for k, p in m.named_parameters():
            if p.grad is None:
                d[f"{name}|{k}|none"] = np.int32(1)
                continue
            g = p.grad.detach().numpy().reshape(-1)
            d[f"{name}|{k}|norm"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
            d[f"{name}|{k}|samples"] = g[sample_positions(g.size, 16)]
        sd1 = m.state_dict()
        for k in sd1:
            if k.endswith(("running_mean", "running_var")):
                d[f"{name}|{k}|buf"] = sd1[k].numpy().reshape(-1)[:8].copy()
        m.load_state_dict(sd0)  # restore the BatchNorm buffers for the next case
        print("train_step", name, "loss", loss.item(), items.detach().numpy())
    # optimizer parameter groups (engine/trainer.py:784-808): 0 = weight with decay, 1 = norm weight (no decay), 2 = bias (no decay)
    from torch import nn
    bn = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
    groups = {}
    for module_name, module in m.named_modules():
        for param_name, param in module.named_parameters(recurse=False):
            fullname = f"{module_name}.{param_name}" if module_name else param_name
            groups[fullname] = 2 if "bias" in fullname else (1 if isinstance(module, bn) else 0)
    with open(os.path.join(GOLD, "optimizer_groups.json"), "w") as f:
        json.dump(groups, f)
    np.savez_compressed(os.path.join(GOLD, "train_step.npz"), **d)


def gen_train_step_nc3():
    """The training step of gen_train_step for a model built with nc = 3 (a custom dataset: DetectionModel(yaml, nc=3) resizes the head's cv3,
    nn/tasks.py:319-326, nn/modules/head.py:1094-1097): pins oracle.model.train_step_grads for a class count that is not a multiple of 8, the case
    the CUDA training head pads (tests/test_gpu_train_step.py::test_train_step_with_a_class_count_that_is_not_a_multiple_of_8).  The synthetic
    state-dict spec is the committed one with cv3 resized; the committed spec file is not touched."""
    from oracle.cases import TRAIN_STEP_CASES, train_step_inputs
    nc = 3
    m = DetectionModel(YAML, ch=3, nc=nc, verbose=False)
    spec = [(k, list(v.shape), str(v.dtype)) for k, v in m.state_dict().items()]
    want = [[k, ([nc] + list(sh[1:]) if k in ("model.33.cv3.weight", "model.33.cv3.bias") else list(sh)), dt] for k, sh, dt in synth.load_spec()]
    assert [list(e) for e in spec] == want, "the nc = 3 model differs from the committed spec in more than cv3"
    m.load_state_dict(synth.make_state_dict(seed=5, spec=spec), strict=True)
    m.args = _Args()
    m.train()
    m.criterion = None
    m.model[-1].shape = None
    img, bi, cl, bb = train_step_inputs(**TRAIN_STEP_CASES["b2_160"])
    cl = (cl % nc).astype(cl.dtype)
    loss, items = m.loss(dict(img=torch.from_numpy(img), batch_idx=torch.from_numpy(bi), cls=torch.from_numpy(cl), bboxes=torch.from_numpy(bb)))
    loss.backward()
    d = {"loss": np.float64(loss.item()), "items": items.detach().numpy().astype(np.float64)}
    for k, p in m.named_parameters():
        if p.grad is None:
            d[f"{k}|none"] = np.int32(1)
            continue
        g = p.grad.detach().numpy().reshape(-1)
        d[f"{k}|norm"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        d[f"{k}|samples"] = g[sample_positions(g.size, 16)]
    print("train_step_nc3 loss", loss.item(), items.detach().numpy())
    np.savez_compressed(os.path.join(GOLD, "train_step_nc3.npz"), **d)


OPT_HYP = dict(lr=0.01, momentum=0.937, weight_decay=5e-4, steps=2)
ADAMW_LR = 1e-3


def gen_opt_step(name="SGD", out="opt_step.npz"):
    """Two optimizer steps of the live reference on one batch: BaseTrainer.build_optimizer's own parameter grouping + torch.optim.SGD
    (engine/trainer.py:754-808), clip_grad_norm_(10) + step (trainer.py:580-588), ModelEMA.update (utils/torch_utils.py:511-541).
    Stores sampled parameter / EMA DELTAS (after - before): pins oracle/optim.py and the yad_sgd_step / yad_ema_update kernels."""
    from types import SimpleNamespace
    from oracle.cases import TRAIN_STEP_CASES, train_step_inputs
    from ultralytics.engine.trainer import BaseTrainer
    from ultralytics.utils.torch_utils import ModelEMA
    m, _ = build_reference_model(fuse=False)
    m.args = _Args()
    m.train()
    for p_ in m.parameters():
        p_.requires_grad_(True)
    m.model[-1].dfl.conv.weight.requires_grad_(False)  # block.py:72
    h = dict(OPT_HYP)
    if name == "AdamW":
        h["lr"] = ADAMW_LR
    opt = BaseTrainer.build_optimizer(SimpleNamespace(args=SimpleNamespace(warmup_bias_lr=0.1)), m, name=name, lr=h["lr"], momentum=h["momentum"],
                                      decay=h["weight_decay"], iterations=1e5)
    ema = ModelEMA(m)
    img, bi, cl, bb = train_step_inputs(**TRAIN_STEP_CASES["b2_160"])
    batch = dict(img=torch.from_numpy(img), batch_idx=torch.from_numpy(bi), cls=torch.from_numpy(cl), bboxes=torch.from_numpy(bb))
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    d = {}
    for step in range(h["steps"]):
        m.criterion = None
        m.model[-1].shape = None
        loss, items = m.loss(batch)
        loss.backward()
        norm = torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=10.0)
        opt.step()
        opt.zero_grad()
        ema.update(m)
        d[f"loss{step}"] = np.float64(loss.item())
        d[f"gradnorm{step}"] = np.float64(norm.item())
        print("opt step", step, "loss", loss.item(), "grad norm", norm.item())
    sd1, sde = m.state_dict(), ema.ema.state_dict()
    for k, v in sd1.items():
        if not v.dtype.is_floating_point:
            continue
        pos = sample_positions(v.numel(), 16)
        d[f"{k}|delta"] = (v.double() - sd0[k].double()).numpy().reshape(-1)[pos]
        d[f"{k}|ema_delta"] = (sde[k].double() - sd0[k].double()).numpy().reshape(-1)[pos]
        d[f"{k}|delta_norm"] = np.float64((v.double() - sd0[k].double()).norm().item())
    np.savez_compressed(os.path.join(GOLD, out), **d)


def gen_preprocess():
    """LetterBox + the predictor's BGR->RGB / HWC->CHW step (engine/predictor.py:115-133, 144-156) and scale_boxes (utils/ops.py:88-123) of the
    live reference (cv2 %s) on the seeded images of oracle/cases.py: pins oracle/preprocess.py and the yad_letterbox / yad_scale_boxes kernels."""
    import cv2
    from oracle.cases import PREPROCESS_CASES, preprocess_image, scale_boxes_inputs
    from ultralytics.data.augment import LetterBox
    d = {"cv2_version": np.array(cv2.__version__)}
    for name, (h, w, size, auto, seed) in PREPROCESS_CASES.items():
        img = preprocess_image(h, w, seed)
        lb = LetterBox((size, size), auto=auto, stride=32)(image=img)
        chw = np.ascontiguousarray(np.stack([lb])[..., ::-1].transpose((0, 3, 1, 2)))[0]  # predictor.py:127-129
        d[name] = chw
        boxes = torch.from_numpy(scale_boxes_inputs(seed + 100) * (size / 640.0))
        d[name + "_boxes"] = ref_ops.scale_boxes(chw.shape[1:], boxes.clone(), (h, w)).numpy()
        print("preprocess", name, chw.shape)
    import zlib
    from oracle.cases import PREPROCESS_SWEEP
    crc = []
    for h, w, seed in PREPROCESS_SWEEP:
        lb = LetterBox((96, 96), auto=False, stride=32)(image=preprocess_image(h, w, seed))
        crc.append(zlib.crc32(np.ascontiguousarray(lb[..., ::-1].transpose(2, 0, 1)).tobytes()))
    d["sweep_crc"] = np.array(crc, np.uint32)
    np.savez_compressed(os.path.join(GOLD, "preprocess.npz"), **d)


def gen_match():
    """DetectionValidator._process_batch (models/yolo/detect/val.py:209-227 -> box_iou, utils/metrics.py:52-71 -> BaseValidator.match_predictions,
    engine/validator.py:221-261) of the live reference on the seeded cases of oracle/cases.py: pins oracle/metrics.py and yad_match_predictions."""
    from types import SimpleNamespace
    from oracle.cases import MATCH_CASES, match_inputs
    from ultralytics.engine.validator import BaseValidator
    from ultralytics.utils.metrics import box_iou
    me = SimpleNamespace(iouv=torch.linspace(0.5, 0.95, 10))
    d = {}
    for name, args in MATCH_CASES.items():
        det, gt, gt_cls = (torch.from_numpy(a) for a in match_inputs(*args))
        iou = box_iou(gt, det[:, :4])
        d[name] = BaseValidator.match_predictions(me, det[:, 5], gt_cls, iou).numpy()
        d[name + "_iou"] = iou.numpy()
        print("match", name, d[name].shape, int(d[name].sum()))
    np.savez_compressed(os.path.join(GOLD, "match_cases.npz"), **d)


def gen_val():
    """The live reference's validation statistics on the seeded sets of oracle/cases.py: DetectionValidator._prepare_batch / _prepare_pred /
    _process_batch per image exactly as update_metrics drives them (models/yolo/detect/val.py:104-176), then ap_per_class (utils/metrics.py:1144-1231)
    on the concatenated stats as get_stats / DetMetrics.process do (val.py:183-191).  Pins oracle/metrics.py and the yad_val_* kernels."""
    from types import SimpleNamespace
    from oracle.cases import VAL_CASES, val_inputs
    from ultralytics.engine.validator import BaseValidator
    from ultralytics.models.yolo.detect.val import DetectionValidator
    from ultralytics.utils.metrics import ap_per_class
    me = SimpleNamespace(iouv=torch.linspace(0.5, 0.95, 10), device=torch.device("cpu"))
    me.match_predictions = lambda *a: BaseValidator.match_predictions(me, *a)
    d = {}
    for name, args in VAL_CASES.items():
        v = val_inputs(*args)
        batch = dict(batch_idx=torch.from_numpy(v["batch_idx"]), cls=torch.from_numpy(v["cls"])[:, None], bboxes=torch.from_numpy(v["bboxes"]),
                     ori_shape=v["ori_shape"], ratio_pad=v["ratio_pad"], img=torch.zeros(len(v["dets"]), 3, v["imgsz"], v["imgsz"]))
        tp, conf, pcls, tcls, labels = [], [], [], [], []
        for si, det in enumerate(v["dets"]):
            pb = DetectionValidator._prepare_batch(me, si, batch)
            cls, bbox = pb.pop("cls"), pb.pop("bbox")
            labels.append(bbox.numpy().reshape(-1, 4))
            if len(det) == 0:
                if len(cls):
                    tcls.append(cls.numpy())
                continue
            predn = DetectionValidator._prepare_pred(me, torch.from_numpy(det), pb)
            t = DetectionValidator._process_batch(me, predn, bbox, cls).numpy() if len(cls) else np.zeros((len(det), 10), bool)
            tp.append(t), conf.append(predn[:, 4].numpy()), pcls.append(predn[:, 5].numpy()), tcls.append(cls.numpy())
        tp, conf, pcls, tcls = np.concatenate(tp), np.concatenate(conf), np.concatenate(pcls), np.concatenate(tcls)
        r = ap_per_class(tp, conf, pcls, tcls)
        d[name + "_labels"] = np.concatenate(labels)
        d[name + "_tp"] = np.packbits(tp, axis=0)
        d[name + "_n"] = np.array(len(tp))
        for k, a in zip(("tpn", "fpn", "p", "r", "f1", "ap", "classes", "p_curve", "r_curve", "f1_curve"), r[:10]):
            d[f"{name}_{k}"] = np.asarray(a)[:, ::8] if k.endswith("curve") else np.asarray(a)   # curves: every 8th of the 1000 columns
        from ultralytics.utils.metrics import smooth
        d[name + "_f1_index"] = np.array(smooth(r[9].mean(0), 0.1).argmax())
        print("val", name, tp.shape, "mAP50 %.4f mAP %.4f" % (r[5][:, 0].mean(), r[5].mean()))
    np.savez_compressed(os.path.join(GOLD, "val_cases.npz"), **d)


def gen_mona():
    """Mona.forward of the live reference (nn/modules/mona.py:36-64, eval mode) on the seeded parameters / inputs of oracle/mona.py: pins the oracle
    restatement and, through it, yad_ln_mix + the fused composition of yolo_ad_refine_b200.functional.mona."""
    from oracle.mona import MONA_CASES, make_input, make_state
    from ultralytics.nn.modules.mona import Mona
    d = {}
    for name, (c, n, h, w, seed) in MONA_CASES.items():
        m = Mona(c).eval()
        missing = m.load_state_dict(make_state(c, seed), strict=True)
        with torch.no_grad():
            y = m(make_input(c, n, h, w, seed))
        d[name] = y.numpy() if h * w < 100 else y.numpy()[:, :, ::2, ::2]   # larger maps: every other row / column
        d[name + "_keys"] = np.array(sorted(m.state_dict().keys()))
        print("mona", name, tuple(y.shape), float(y.abs().mean()), missing)
    np.savez_compressed(os.path.join(GOLD, "mona.npz"), **d)


def gen_mona_block():
    """C2TSSA_DYT_Mona_EDFFN.forward of the live reference (nn/modules/block.py:1685-1709, eval) on seeded parameters: pins oracle/mona.py's block
    restatement (DynamicTanh, AttentionTSSA, Mona, EDFFN inside the C2PSA split) and yad_attention_tssa + the fused composition.  The seeded state
    dicts are stored with the outputs (their key set and shapes come from the reference module)."""
    from oracle.mona import BLOCK_CASES, make_block_state, make_input
    from ultralytics.nn.modules.block import C2TSSA_DYT_Mona_EDFFN
    d, shapes = {}, {}
    for name, (c, nb, n, h, w, seed) in BLOCK_CASES.items():
        m = C2TSSA_DYT_Mona_EDFFN(c, c, nb).eval()
        for mod in m.modules():  # what initialize_weights does to every BatchNorm2d of a DetectionModel (utils/torch_utils.py:426-436)
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.eps, mod.momentum = 1e-3, 0.03
        shapes[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
        sd = make_block_state(shapes[name], seed)
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            y = m(make_input(c, n, h, w, seed))
        d[name] = y.numpy()[:, ::4]   # every fourth channel
        print("mona block", name, tuple(y.shape), float(y.abs().mean()), len(sd), "keys")
    np.savez_compressed(os.path.join(GOLD, "mona_block.npz"), **d)
    with open(os.path.join(GOLD, "mona_block_spec.json"), "w") as f:
        json.dump(shapes, f)


def gen_psa_block():
    """C2PSA.forward of the live reference (nn/modules/block.py:874-964, 1010-1049, eval) on seeded parameters: pins oracle/psa.py (Attention with
    its per-head [q | k | v] interleave and key_dim = head_dim / 2, PSABlock, the C2PSA split) and, through it, the yad_mha-based composition of
    yolo_ad_refine_b200.functional.c2psa.  Key sets and shapes come from the reference module."""
    from oracle.mona import make_block_state, make_input
    from oracle.psa import PSA_CASES
    from ultralytics.nn.modules.block import C2PSA
    d, shapes = {}, {}
    for name, (c, nb, n, h, w, seed) in PSA_CASES.items():
        m = C2PSA(c, c, nb).eval()
        for mod in m.modules():  # what initialize_weights does to every BatchNorm2d of a DetectionModel (utils/torch_utils.py:426-436)
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.eps, mod.momentum = 1e-3, 0.03
        shapes[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
        sd = make_block_state(shapes[name], seed)
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            y = m(make_input(c, n, h, w, seed))
        d[name] = y.numpy()[:, ::4]   # every fourth channel
        print("psa block", name, tuple(y.shape), float(y.abs().mean()), len(sd), "keys")
    np.savez_compressed(os.path.join(GOLD, "psa_block.npz"), **d)
    with open(os.path.join(GOLD, "psa_block_spec.json"), "w") as f:
        json.dump(shapes, f)


def gen_sfa_block():
    """C2SFA.forward of the live reference (nn/modules/block.py:2049-2202, 2358-2373, eval) on seeded parameters: pins oracle/psa.py's C2SFA
    restatement and the composition yolo_ad_refine_b200.functional.c2sfa.  residual_weight1 / 2 are lifted to O(1) so that both branches carry weight."""
    from oracle.mona import make_block_state, make_input
    from oracle.psa import SFA_CASES
    from ultralytics.nn.modules.block import C2SFA
    d, shapes = {}, {}
    for name, (c, nb, n, h, w, seed) in SFA_CASES.items():
        m = C2SFA(c, c, nb).eval()
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.eps, mod.momentum = 1e-3, 0.03
        shapes[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
        sd = sfa_state(shapes[name], seed)
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            y = m(make_input(c, n, h, w, seed))
        d[name] = y.numpy()[:, ::4]   # every fourth channel
        print("sfa block", name, tuple(y.shape), float(y.abs().mean()), len(sd), "keys")
    np.savez_compressed(os.path.join(GOLD, "sfa_block.npz"), **d)
    with open(os.path.join(GOLD, "sfa_block_spec.json"), "w") as f:
        json.dump(shapes, f)


def sfa_state(shapes, seed):
    from oracle.psa import sfa_state as f
    return f(shapes, seed)


def gen_augment():
    """RandomHSV / RandomFlip / Mosaic._mosaic4 of the live reference (data/augment.py:1301-1378, 1380-1472, 657-713; cv2 of this image) on seeded
    images and seeded random streams: pins oracle/augment.py and the yad_hsv_lut / yad_flip / yad_mosaic4 kernels.  Also the exhaustive cv2 checks
    of the restated 8-bit BGR<->HSV arithmetic (all 2^24 colours / all 180 x 256 x 256 HSV triples), recorded as mismatch counts (must be 0)."""
    import random
    import zlib

    import cv2
    from ultralytics.data.augment import Mosaic, RandomFlip, RandomHSV
    from ultralytics.utils.instance import Instances

    from oracle import augment as oa
    from oracle.cases import AUG_HSV_CASES, AUG_MOSAIC_CASES, aug_image
    d = {"cv2_version": np.frombuffer(cv2.__version__.encode(), np.uint8)}
    v = np.arange(256, dtype=np.uint8)
    b, g, r = np.meshgrid(v, v, v, indexing="ij")
    allc = np.stack([b.ravel(), g.ravel(), r.ravel()], 1).reshape(4096, 4096, 3)
    d["bgr2hsv_exhaustive_mismatches"] = np.int64((oa.bgr2hsv(allc) != cv2.cvtColor(allc, cv2.COLOR_BGR2HSV)).any(-1).sum())
    hh, ss, vv = np.meshgrid(v[:180], v, v, indexing="ij")
    allh = np.stack([hh.ravel(), ss.ravel(), vv.ravel()], 1).reshape(-1, 4096, 3)
    d["hsv2bgr_exhaustive_mismatches"] = np.int64((oa.hsv2bgr(allh, "simd") != cv2.cvtColor(allh, cv2.COLOR_HSV2BGR)).any(-1).sum())
    print("exhaustive cv2 checks:", int(d["bgr2hsv_exhaustive_mismatches"]), int(d["hsv2bgr_exhaustive_mismatches"]))
    for name, (h, w, seed) in AUG_HSV_CASES.items():
        img = aug_image(h, w, seed)
        np.random.seed(seed)
        rr = np.random.uniform(-1, 1, 3) * [0.015, 0.7, 0.4] + 1     # the draw RandomHSV makes (cfg/default.yaml hsv_h / hsv_s / hsv_v)
        np.random.seed(seed)
        lab = {"img": img.copy()}
        RandomHSV(0.015, 0.7, 0.4)(lab)
        d[f"hsv_{name}_r"] = rr
        d[f"hsv_{name}_crc"] = np.int64(zlib.crc32(lab["img"].tobytes()))
        d[f"hsv_{name}_sample"] = lab["img"][::7, ::5].copy()
        # flips of the same image with boxes
        boxes = np.random.RandomState(seed).uniform(0.1, 0.9, (5, 4)).astype(np.float32)
        for direction in ("vertical", "horizontal"):
            random.seed(seed)
            hit = random.random() < 0.5
            random.seed(seed)
            inst = Instances(boxes.copy(), np.zeros((0, 1000, 2), np.float32), None, bbox_format="xywh", normalized=True)
            out = RandomFlip(p=0.5, direction=direction)({"img": img.copy(), "instances": inst})
            d[f"flip_{name}_{direction}_hit"] = np.int64(hit)
            d[f"flip_{name}_{direction}_crc"] = np.int64(zlib.crc32(out["img"].tobytes()))
            d[f"flip_{name}_{direction}_boxes"] = out["instances"].bboxes.copy()
    for name, (s, shapes, seed) in AUG_MOSAIC_CASES.items():
        imgs = [aug_image(h, w, seed + i) for i, (h, w) in enumerate(shapes)]
        mz = Mosaic.__new__(Mosaic)
        mz.imgsz, mz.border, mz.n = s, (-s // 2, -s // 2), 4
        random.seed(seed)
        yc, xc = (int(random.uniform(-x, 2 * s + x)) for x in mz.border)
        random.seed(seed)

        def lab(i):
            h, w = shapes[i]
            bx = np.random.RandomState(seed + 10 + i).uniform(0.2, 0.8, (3, 4)).astype(np.float32)
            bx[:, 2:] *= 0.3
            return {"img": imgs[i], "resized_shape": (h, w), "im_file": f"img{i}.jpg", "ori_shape": (h, w), "cls": np.full((3, 1), float(i), np.float32),
                    "instances": Instances(bx, np.zeros((0, 1000, 2), np.float32), None, bbox_format="xywh", normalized=True)}
        labels = lab(0)
        labels["mix_labels"] = [lab(1), lab(2), lab(3)]
        out = mz._mosaic4(labels)
        d[f"mosaic_{name}_center"] = np.asarray([yc, xc], np.int64)
        d[f"mosaic_{name}_crc"] = np.int64(zlib.crc32(out["img"].tobytes()))
        d[f"mosaic_{name}_sample"] = out["img"][::9, ::9].copy()
        out["instances"].convert_bbox("xyxy")
        d[f"mosaic_{name}_boxes_xyxy"] = out["instances"].bboxes.copy()
        d[f"mosaic_{name}_cls"] = out["cls"].copy()
        print("mosaic", name, out["img"].shape, yc, xc)
    np.savez_compressed(os.path.join(GOLD, "augment.npz"), **d)


def main():
    if sys.argv[1:] == ["augment"]:
        return gen_augment()
    if sys.argv[1:] == ["sfa_block"]:
        return gen_sfa_block()
    if sys.argv[1:] == ["psa_block"]:
        return gen_psa_block()
    if sys.argv[1:] == ["mona_block"]:
        return gen_mona_block()
    if sys.argv[1:] == ["mona"]:
        return gen_mona()
    if sys.argv[1:] == ["match"]:
        return gen_match()
    if sys.argv[1:] == ["val"]:
        return gen_val()
    if sys.argv[1:] == ["preprocess"]:
        return gen_preprocess()
    if sys.argv[1:] == ["train_step"]:
        return gen_train_step()
    if sys.argv[1:] == ["train_step_nc3"]:
        return gen_train_step_nc3()
    if sys.argv[1:] == ["opt_step"]:
        return gen_opt_step()
    if sys.argv[1:] == ["opt_step_adamw"]:
        return gen_opt_step("AdamW", "opt_step_adamw.npz")
    m, spec = build_reference_model()
    sd = synth.make_state_dict_np(seed=1, spec=spec)
    with open(os.path.join(GOLD, "state_checksum.json"), "w") as f:
        json.dump({"seed": 1, "checksum": synth.state_checksum(sd), "n_keys": len(spec)}, f)
    which = sys.argv[1:] or ["model", "nms", "train"]
    if "model" in which:
        gen_model(m)
    if "nms" in which:
        gen_nms()
    if "train" in which:
        gen_train(m)


if __name__ == "__main__":
    main()
