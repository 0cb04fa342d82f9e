"""Drop-in installation into an Ultralytics tree (the reference fork): rebinds the names that `parse_model`, the predictor / validator and the
trainer look up at call time (SURVEY.md section 8b "How to bind"), so that `YOLO(yaml).predict/val` run the YOLO-AD-Refine hot path
through libyad.so without touching the reference's source.

    import yolo_ad_refine_b200.plugin as yad
    yad.install()                       # before building the model
    model = YOLO("z-yaml/yolo11-701-YOLO-AD-Refine.yaml", task="detect")
    yad.convert_model(model.model)      # swaps the yaml's plain nn.Conv2d / nn.ConvTranspose2d layers for libyad-backed twins

Training keeps the reference's loop as it is (engine/trainer.py:382-397, 580-588: `self.model(batch)`, `.backward()`, clip, optimizer, EMA): install()
also wraps `DetectionModel.loss` (nn/tasks.py:290-302) so that a training-mode `model(batch)` runs forward + TaskAlignedAssigner + v8DetectionLoss +
the complete backward in libyad.so (bridge.TrainBridge) and returns a loss tensor whose `.backward()` hands every parameter its gradient.
`YadDetectionTrainer` (below) is the reference's DetectionTrainer with exactly the two changes a libyad model needs: the yaml's plain torch
layers are converted after the model is built, and AMP's fp16 autocast / GradScaler is off (activations are bf16 inside the kernels).
"""
import importlib

import torch.nn as nn

from . import bridge as ybridge
from . import loss as yloss
from . import modules as M
from . import postprocess as ypost
from . import tal as ytal

BLOCKS = ["Conv", "C3k2", "C3k2_MLCA", "SPPF", "C2PTSSA", "C2ProgressiveTSSA_Fusion", "ELA_HSFPN", "Multiply", "Add", "Fusion",
          # layer 10 of the yolo11-mona / 687 / 689 / 697 sibling yamls
          "C2TSSA_DYT_Mona_EDFFN", "TSSAlock_DYT_Mona_EDFFN", "DynamicTanh", "AttentionTSSA", "Mona", "MonaOp",
          # the stock yolo11 attention block the other sibling yamls keep at layer 10
          "C2PSA", "PSABlock", "Attention",
          # yolo11-hsfpn+C2SFA.yaml and its copies
          "C2SFA", "ProgressiveTSSA_Fusion0", "SimpleFeatureProcessor", "SEBlock", "StandardFFN"]
HEADS = ["AYHead", "AYHead1"]


_ORIGINALS = {}  # (module object, attribute) -> the reference's own object, recorded by install()


class originals:
    """Context manager: the reference's own classes / functions are bound back for the duration of the block (and the libyad ones again
    afterwards).  checkpoint.save_reference_checkpoint uses it to build and pickle a genuine reference DetectionModel, whose classes are recorded by
    qualified name -- so the file loads in an unmodified reference checkout."""

    def __enter__(self):
        self._swapped = [(m, a, getattr(m, a)) for (m, a) in _ORIGINALS]
        for (m, a), obj in _ORIGINALS.items():
            setattr(m, a, obj)
        return self

    def __exit__(self, *exc):
        for m, a, obj in self._swapped:
            setattr(m, a, obj)
        return False


def install(ultralytics_pkg="ultralytics"):
    """Rebind the reference's names to the libyad-backed implementations.  Returns the list of rebound qualified names."""
    done = []

    def bind(modname, attr, obj):
        try:
            mod = importlib.import_module(f"{ultralytics_pkg}.{modname}")
        except Exception:
            return
        if hasattr(mod, attr):
            _ORIGINALS.setdefault((mod, attr), getattr(mod, attr))
            setattr(mod, attr, obj)
            done.append(f"{modname}.{attr}")

    for name in BLOCKS:
        obj = getattr(M, name)
        bind("nn.tasks", name, obj)          # parse_model resolves classes through globals() (nn/tasks.py:969)
        bind("nn.modules", name, obj)
        bind("nn.modules.block", name, obj)  # unpickling of checkpoints
    bind("nn.modules.conv", "Conv", M.Conv)
    bind("nn.modules.mona", "Mona", M.Mona)
    bind("nn.modules.mona", "MonaOp", M.MonaOp)
    for name in HEADS:
        bind("nn.tasks", name, M.AYHead)
        bind("nn.modules", name, M.AYHead)
        bind("nn.modules.head", name, M.AYHead)
    bind("utils.ops", "non_max_suppression", ypost.non_max_suppression)  # looked up at call time (detect/predict.py:25, detect/val.py:93)
    bind("nn.tasks", "v8DetectionLoss", yloss.v8DetectionLoss)           # nn/tasks.py:396-398
    bind("utils.loss", "v8DetectionLoss", yloss.v8DetectionLoss)
    bind("utils.loss", "TaskAlignedAssigner", ytal.TaskAlignedAssigner)   # utils/loss.py:379
    bind("utils.tal", "TaskAlignedAssigner", ytal.TaskAlignedAssigner)
    try:  # training-mode model(batch) -> libyad forward + loss + backward (bridge.py)
        tasks = importlib.import_module(f"{ultralytics_pkg}.nn.tasks")
        if not getattr(tasks.DetectionModel.loss, "_yad_wrapped", False):
            _ORIGINALS.setdefault((tasks.DetectionModel, "loss"), tasks.DetectionModel.loss)
            tasks.DetectionModel.loss = ybridge.model_loss(tasks.DetectionModel.loss)
            done.append("nn.tasks.DetectionModel.loss")
    except Exception:
        pass
    return done


def make_trainer_class(ultralytics_pkg="ultralytics"):
    """The reference's DetectionTrainer (models/yolo/detect/train.py) for a libyad-backed model: same data pipeline, loop, optimizer, EMA,
    validation and checkpoints; get_model() converts the yaml's plain torch layers and AMP is forced off (bf16 lives inside the kernels, there is
    no fp16 autocast to scale for).  Usage: `make_trainer_class()(overrides=dict(model=yaml, data=..., epochs=...)).train()` or
    `YOLO(yaml).train(trainer=make_trainer_class(), ...)`."""
    install(ultralytics_pkg)
    det = importlib.import_module(f"{ultralytics_pkg}.models.yolo.detect")

    class YadDetectionTrainer(det.DetectionTrainer):
        def __init__(self, cfg=None, overrides=None, _callbacks=None):
            overrides = dict(overrides or {})
            overrides["amp"] = False
            kw = {} if cfg is None else {"cfg": cfg}
            super().__init__(overrides=overrides, _callbacks=_callbacks, **kw)

        def get_model(self, cfg=None, weights=None, verbose=True):
            model = super().get_model(cfg=cfg, weights=weights, verbose=verbose)
            return convert_model(model)

    return YadDetectionTrainer


def convert_model(model):
    """Replace the plain torch layers the yaml names (`nn.Conv2d` laterals, `nn.ConvTranspose2d` upsamplers) in a parsed model by their
    libyad-backed twins, keeping parameters (same state-dict keys) and the routing attributes parse_model attached (i, f, type, np)."""
    seq = model.model if hasattr(model, "model") else model
    for idx, m in enumerate(seq):
        new = None
        if type(m) is nn.Conv2d and m.groups == 1:
            new = M.YadConv2d(m.in_channels, m.out_channels, m.kernel_size[0], m.stride[0], m.padding[0], m.bias is not None)
        elif type(m) is nn.ConvTranspose2d:
            new = M.YadConvTranspose2d(m.in_channels, m.out_channels, m.kernel_size[0], m.stride[0], m.padding[0], m.output_padding[0])
        if new is None:
            continue
        new.weight, new.bias = m.weight, m.bias
        for attr in ("i", "f", "type", "np"):
            if hasattr(m, attr):
                setattr(new, attr, getattr(m, attr))
        new.train(m.training)
        seq[idx] = new
    return model
