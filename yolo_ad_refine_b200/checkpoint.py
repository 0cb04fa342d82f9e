"""Checkpoint I/O in the reference's own format (row f4 of SURVEY.md section 8): what `BaseTrainer.save_model` writes (engine/trainer.py:507-540)
and what `attempt_load_one_weight` (nn/tasks.py:917-940) / `BaseTrainer.resume_training` (engine/trainer.py:718-743) read.

The reference pickles the EMA *module* (`deepcopy(self.ema.ema).half()`), i.e. class paths `ultralytics.nn.tasks.DetectionModel`,
`ultralytics.nn.modules.block.*`, ...; a state-dict file would not load there.  save_reference_checkpoint therefore builds a genuine reference
`DetectionModel` (the reference's own classes, restored through plugin.originals() when the plugin is installed), loads the weights into it and
pickles that -- so `last.pt` written from a TrainEngine run loads in an unmodified reference checkout, resumes in its trainer (optimizer state in
torch.optim's state-dict layout, parameter order of `build_optimizer`, engine/trainer.py:784-808) and predicts with `YOLO(path)`.
load_reference_checkpoint is the inverse: any checkpoint of the reference (its own classes or the libyad mirrors) -> reference-layout state dict,
optimizer arenas and counters for TrainParams / RefineEngine.

Needs the `ultralytics` package of the reference importable (it is the environment the plugin lives in).  File I/O only: no kernels, no device.
"""
import importlib
import math
from copy import deepcopy
from datetime import datetime

import torch
import torch.nn as nn


def _build_reference_model(state_dict, yaml, nc, ultralytics_pkg):
    from . import plugin
    with plugin.originals():
        tasks = importlib.import_module(f"{ultralytics_pkg}.nn.tasks")
        model = tasks.DetectionModel(yaml, ch=3, nc=nc, verbose=False)
    model.load_state_dict({k: v.detach().float().cpu() for k, v in state_dict.items()}, strict=True)
    return model


def optimizer_param_order(model):
    """(names per group) in the order `BaseTrainer.build_optimizer` hands parameters to torch.optim (engine/trainer.py:784-808): group 0 of the
    optimizer = biases, group 1 = weights with decay, group 2 = normalisation weights"""
    bn = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
    g = [], [], []
    for module_name, module in model.named_modules():
        for param_name, _ in module.named_parameters(recurse=False):
            fullname = f"{module_name}.{param_name}" if module_name else param_name
            if "bias" in fullname:
                g[2].append(fullname)
            elif isinstance(module, bn):
                g[1].append(fullname)
            else:
                g[0].append(fullname)
    return [g[2], g[0], g[1]]


def optimizer_state_dict(model, arenas, optimizer="SGD", lr=0.01, momentum=0.937, weight_decay=5e-4, step=0, beta2=0.999, eps=1e-8, has_grad=None):
    """torch.optim state dict (the layout `optimizer.state_dict()` produces) from per-parameter tensors.  arenas: dict name -> dict of state tensors
    ({"momentum_buffer": t} for SGD; {"exp_avg", "exp_avg_sq"} for AdamW).  has_grad(name) -> False for parameters torch never creates state for
    (frozen: DFL's projection, AdaptiveDynamicTanh.scale_weights)."""
    groups = optimizer_param_order(model)
    state, param_groups, idx = {}, [], 0
    for gi, names in enumerate(groups):
        ids = []
        for name in names:
            if has_grad is None or has_grad(name):
                st = {k: v.detach().clone().cpu() for k, v in arenas[name].items()}
                if optimizer == "AdamW":
                    st["step"] = torch.tensor(float(step))
                state[idx] = st
            ids.append(idx)
            idx += 1
        if optimizer == "SGD":
            pg = dict(lr=lr, momentum=momentum, dampening=0, weight_decay=weight_decay if gi == 1 else 0.0, nesterov=True, maximize=False, foreach=None,
                      differentiable=False, fused=None, initial_lr=lr, params=ids)
        else:
            pg = dict(lr=lr, betas=(momentum, beta2), eps=eps, weight_decay=weight_decay if gi == 1 else 0.0, amsgrad=False, maximize=False, foreach=None,
                      capturable=False, differentiable=False, fused=None, initial_lr=lr, params=ids)
        param_groups.append(pg)
    return dict(state=state, param_groups=param_groups)


def save_reference_checkpoint(path, state_dict, yaml, nc=80, epoch=0, best_fitness=None, updates=0, optimizer=None, train_args=None,
                              train_metrics=None, half=True, ultralytics_pkg="ultralytics"):
    """Write `path` as BaseTrainer.save_model would: {"epoch", "best_fitness", "model": None, "ema": <DetectionModel, fp16>, "updates", "optimizer",
    "train_args", "train_metrics", "train_results", "date", "version", "license", "docs"}.  state_dict: the EMA weights in the reference's
    state-dict layout (TrainParams.state_dict(ema=True)); optimizer: a torch.optim state dict (optimizer_state_dict above) or None."""
    model = _build_reference_model(state_dict, yaml, nc, ultralytics_pkg)
    pkg = importlib.import_module(ultralytics_pkg)
    ckpt = {"epoch": epoch, "best_fitness": best_fitness, "model": None, "ema": (model.half() if half else model), "updates": updates,
            "optimizer": optimizer, "train_args": dict(train_args or {}), "train_metrics": dict(train_metrics or {}), "train_results": {},
            "date": datetime.now().isoformat(), "version": getattr(pkg, "__version__", "unknown"),
            "license": "AGPL-3.0 (https://ultralytics.com/license)", "docs": "https://docs.ultralytics.com"}
    if optimizer is not None:  # convert_optimizer_state_dict_to_fp16 (utils/torch_utils.py:640-651)
        optimizer = deepcopy(optimizer)
        for st in optimizer["state"].values():
            for k, v in st.items():
                if k != "step" and isinstance(v, torch.Tensor) and v.dtype is torch.float32:
                    st[k] = v.half()
        ckpt["optimizer"] = optimizer
    from . import plugin
    with plugin.originals():  # pickle looks the classes up by qualified name: they must resolve to the objects being pickled
        torch.save(ckpt, path)
    return model


def load_reference_checkpoint(path, ultralytics_pkg="ultralytics"):
    """Read a checkpoint the reference (or a plugin-driven run) wrote.  Returns dict(state_dict (fp32, reference layout), epoch, best_fitness, updates,
    train_args, optimizer (torch.optim state dict or None), param_order (names per optimizer group))."""
    importlib.import_module(ultralytics_pkg)  # the pickled classes live there
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    model = ckpt.get("ema") or ckpt["model"]
    sd = {k: (v.float() if v.dtype.is_floating_point else v) for k, v in model.state_dict().items()}
    return dict(state_dict=sd, epoch=ckpt.get("epoch", -1), best_fitness=ckpt.get("best_fitness"), updates=ckpt.get("updates", 0),
                train_args=ckpt.get("train_args", {}), optimizer=ckpt.get("optimizer"), param_order=optimizer_param_order(model))


# ---- TrainParams <-> checkpoint ----------------------------------------------------------------------------------------------------
def save_train_params(path, tp, yaml, nc=80, epoch=0, best_fitness=None, lr=0.01, momentum=0.937, weight_decay=5e-4, train_args=None,
                      ultralytics_pkg="ultralytics"):
    """`last.pt` of a TrainEngine run: EMA weights + BatchNorm buffers as the pickled fp16 reference model, the flat momentum (and AdamW second
    moment) arenas re-cut into torch.optim's per-parameter state in build_optimizer's order, ModelEMA's update count."""
    from .train_params import is_frozen
    sd = tp.state_dict(ema=True)
    model = _build_reference_model(sd, yaml, nc, ultralytics_pkg)
    opt = getattr(tp, "optimizer_name", None) or "SGD"

    def view(arena, k):
        n = math.prod(tp.shape[k])
        return arena[tp.off[k]:tp.off[k] + n].view(tp.shape[k])

    arenas = {}
    for k in tp.keys:
        arenas[k] = {"momentum_buffer": view(tp.mom, k)} if opt == "SGD" else {"exp_avg": view(tp.mom, k), "exp_avg_sq": view(tp.mom2, k)}
    osd = optimizer_state_dict(model, arenas, opt, lr, momentum, weight_decay, step=tp.steps, has_grad=lambda k: not is_frozen(k)) if tp.steps else None
    save_reference_checkpoint(path, sd, yaml, nc, epoch, best_fitness, tp.ema_updates, osd, train_args, ultralytics_pkg=ultralytics_pkg)


def load_into_train_params(path, tp, ultralytics_pkg="ultralytics"):
    """Resume a TrainEngine from a reference-format checkpoint: parameters AND EMA start from the checkpoint's (EMA) weights, as the reference's
    resume does (trainer.py:728; the raw model is `None` in its files), optimizer state goes back into the flat arenas."""
    ck = load_reference_checkpoint(path, ultralytics_pkg)
    sd = ck["state_dict"]
    with torch.no_grad():
        for k in tp.keys:
            n = math.prod(tp.shape[k])
            v = sd[k].reshape(-1).to(tp.device)
            tp.flat[tp.off[k]:tp.off[k] + n].copy_(v)
            tp.ema[tp.off[k]:tp.off[k] + n].copy_(v)
        for k in tp.buf_keys:
            n = math.prod(tp.shape[k])
            v = sd[k].reshape(-1).to(tp.device)
            tp.bufs[tp.buf_off[k]:tp.buf_off[k] + n].copy_(v)
            tp.ema_bufs[tp.buf_off[k]:tp.buf_off[k] + n].copy_(v)
        for k in tp.other:
            if k in sd:
                tp.other[k] = sd[k].clone()
        osd = ck["optimizer"]
        if osd is not None:
            names = [n for grp in ck["param_order"] for n in grp]
            adam = any("exp_avg" in st for st in osd["state"].values())
            if adam and not hasattr(tp, "mom2"):
                tp.mom2 = torch.zeros_like(tp.mom)
            for idx, st in osd["state"].items():
                k = names[int(idx)]
                n = math.prod(tp.shape[k])
                sl = slice(tp.off[k], tp.off[k] + n)
                if adam:
                    tp.mom[sl].copy_(st["exp_avg"].float().reshape(-1).to(tp.device))
                    tp.mom2[sl].copy_(st["exp_avg_sq"].float().reshape(-1).to(tp.device))
                    tp.steps = int(st.get("step", tp.steps))
                else:
                    tp.mom[sl].copy_(st["momentum_buffer"].float().reshape(-1).to(tp.device))
                    tp.steps = max(tp.steps, 1)
            tp.optimizer_name = "AdamW" if adam else "SGD"
    tp.ema_updates = int(ck["updates"] or 0)
    return ck
