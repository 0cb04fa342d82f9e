"""Row f4 (checkpoint format), CPU: the reference-format checkpoint writer / reader of yolo_ad_refine_b200/checkpoint.py against the LIVE reference
(/root/reference in the build container, oracle/_ref on the GPU box; skipped when neither exists).
  * a file written by save_train_params loads in a clean interpreter that has never seen this package, through the reference's own
    attempt_load_one_weight, and its optimizer state loads into the optimizer the reference's build_optimizer logic creates;
  * a checkpoint assembled the way BaseTrainer.save_model does (pickled fp16 EMA module, convert_optimizer_state_dict_to_fp16) reads back into
    TrainParams: weights, BatchNorm buffers, momentum buffers in build_optimizer's parameter order, the EMA update count."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_ref():
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("no reference tree (oracle/_ref or /root/reference)")
    ref_shims.install()
    return ref_shims


def _tp(sd):
    from yolo_ad_refine_b200.train_params import TrainParams
    return TrainParams(sd, torch.float32, "cpu")


def test_writer_output_loads_in_the_unmodified_reference(tmp_path, state_dict):
    rs = _need_ref()
    from yolo_ad_refine_b200 import checkpoint as ck
    yaml = os.path.join(rs.REFERENCE_ROOT, "z-yaml", "yolo11-701-YOLO-AD-Refine.yaml")
    tp = _tp(state_dict)
    g = torch.Generator().manual_seed(0)
    tp.mom.copy_(torch.randn(tp.total, generator=g) * 0.01)
    tp.ema.copy_(tp.flat + torch.randn(tp.total, generator=g) * 1e-3)
    tp.steps, tp.ema_updates, tp.optimizer_name = 7, 7, "SGD"
    path = str(tmp_path / "last.pt")
    ck.save_train_params(path, tp, yaml, nc=80, epoch=3, best_fitness=0.25, lr=0.01, train_args=dict(imgsz=640, batch=16))
    # a clean interpreter: the reference alone (import shims for the packages this image lacks), no yolo_ad_refine_b200 anywhere
    code = r"""
import json, sys
sys.path.insert(0, %r)
from oracle import ref_shims
ref_shims.install()
assert not any(m.startswith("yolo_ad_refine_b200") for m in sys.modules)
import torch
from ultralytics.nn.tasks import attempt_load_one_weight
from ultralytics.nn import tasks
model, ckpt = attempt_load_one_weight(%r)
assert type(model).__module__ == "ultralytics.nn.tasks" and type(model.model[10]).__module__ == "ultralytics.nn.modules.block"
assert not any(m.startswith("yolo_ad_refine_b200") for m in sys.modules)
sd = model.state_dict()
# the optimizer the reference's build_optimizer would create takes the saved state (engine/trainer.py:725, 784-808)
import torch.nn as nn
bn = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
g = [], [], []
for mn, m in model.named_modules():
    for pn, p in m.named_parameters(recurse=False):
        fn = f"{mn}.{pn}" if mn else pn
        (g[2] if "bias" in fn else g[1] if isinstance(m, bn) else g[0]).append(p)
opt = torch.optim.SGD(g[2], lr=0.01, momentum=0.937, nesterov=True)
opt.add_param_group({"params": g[0], "weight_decay": 5e-4})
opt.add_param_group({"params": g[1], "weight_decay": 0.0})
opt.load_state_dict(ckpt["optimizer"])
w = dict(model.named_parameters())["model.33.cv3.weight"]
mb = opt.state[w]["momentum_buffer"]
with torch.inference_mode():
    y = model(torch.zeros(1, 3, 160, 160))[0]
print("RESULT " + json.dumps(dict(epoch=ckpt["epoch"], updates=ckpt["updates"], best=ckpt["best_fitness"], n=len(sd), yshape=list(y.shape),
      w0=float(sd["model.0.conv.weight"].flatten()[0]), mb_sum=float(mb.float().sum()), states=len(ckpt["optimizer"]["state"]),
      imgsz=ckpt["train_args"]["imgsz"], ema_dtype=str(next(ckpt["ema"].parameters()).dtype))))
""" % (ROOT, path)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    assert res["epoch"] == 3 and res["updates"] == 7 and res["best"] == 0.25 and res["imgsz"] == 640 and res["n"] == len(state_dict)
    assert res["yshape"] == [1, 84, 525]
    ema0 = float(tp.state_dict(ema=True)["model.0.conv.weight"].half().float().flatten()[0])
    assert abs(res["w0"] - ema0) < 1e-6
    n = tp.shape["model.33.cv3.weight"]
    want = float(tp.mom[tp.off["model.33.cv3.weight"]:tp.off["model.33.cv3.weight"] + int(np.prod(n))].half().float().sum())
    assert abs(res["mb_sum"] - want) < 1e-3 * (abs(want) + 1e-3)
    from yolo_ad_refine_b200.train_params import is_frozen
    # frozen tensors (DFL's projection, AdaptiveDynamicTanh.scale_weights) carry no optimizer state; attempt_load_one_weight calls .float()
    assert res["states"] == sum(not is_frozen(k) for k in tp.keys) and "float32" in res["ema_dtype"]


def test_reader_takes_a_checkpoint_the_reference_trainer_would_write(tmp_path, state_dict):
    rs = _need_ref()
    from copy import deepcopy

    from ultralytics.nn.tasks import DetectionModel
    from ultralytics.utils.torch_utils import convert_optimizer_state_dict_to_fp16
    from yolo_ad_refine_b200 import checkpoint as ck
    yaml = os.path.join(rs.REFERENCE_ROOT, "z-yaml", "yolo11-701-YOLO-AD-Refine.yaml")
    model = DetectionModel(yaml, ch=3, nc=80, verbose=False)
    model.load_state_dict(state_dict, strict=True)
    # the reference's optimizer (engine/trainer.py:784-808) after one step on synthetic gradients
    import torch.nn as nn
    bn = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
    g = [], [], []
    for mn, m in model.named_modules():
        for pn, p in m.named_parameters(recurse=False):
            fn = f"{mn}.{pn}" if mn else pn
            (g[2] if "bias" in fn else g[1] if isinstance(m, bn) else g[0]).append(p)
    opt = torch.optim.SGD(g[2], lr=0.01, momentum=0.937, nesterov=True)
    opt.add_param_group({"params": g[0], "weight_decay": 5e-4})
    opt.add_param_group({"params": g[1], "weight_decay": 0.0})
    gen = torch.Generator().manual_seed(1)
    for p in model.parameters():
        if p.requires_grad:
            p.grad = torch.randn(p.shape, generator=gen) * 0.01
    opt.step()
    path = str(tmp_path / "ref_last.pt")
    torch.save({"epoch": 5, "best_fitness": 0.5, "model": None, "ema": deepcopy(model).half(), "updates": 123,
                "optimizer": convert_optimizer_state_dict_to_fp16(deepcopy(opt.state_dict())), "train_args": {"imgsz": 640}}, path)
    tp = _tp(state_dict)
    info = ck.load_into_train_params(path, tp)
    assert info["epoch"] == 5 and tp.ema_updates == 123 and tp.optimizer_name == "SGD"
    named = dict(model.named_parameters())
    for k in ("model.0.conv.weight", "model.10.m.0.ffn.fft", "model.33.cv3.bias", "model.33.stems.0.gn.weight"):
        n = int(np.prod(tp.shape[k]))
        got = tp.flat[tp.off[k]:tp.off[k] + n]
        np.testing.assert_allclose(got.numpy(), named[k].detach().half().float().reshape(-1).numpy(), rtol=0, atol=0)
        mb = tp.mom[tp.off[k]:tp.off[k] + n]
        np.testing.assert_allclose(mb.numpy(), opt.state[named[k]]["momentum_buffer"].half().float().reshape(-1).numpy(), rtol=0, atol=0)
    k = "model.0.bn.running_var"
    n = int(np.prod(tp.shape[k]))
    np.testing.assert_allclose(tp.bufs[tp.buf_off[k]:tp.buf_off[k] + n].numpy(), state_dict[k].half().float().numpy())
