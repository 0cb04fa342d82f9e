"""Import shims that let the LIVE reference (/root/reference, read-only) be imported in the
build container so that the oracle restatement (oracle/model.py, oracle/postprocess.py, ...)
can be validated against it and golden vectors can be generated (oracle/gen_golden.py).

TEST INFRASTRUCTURE ONLY.  Nothing under yolo_ad_refine_b200/ imports this file, and it is never
used on the GPU box (/root/reference does not exist there).

Packages the reference imports but this image lacks:
  matplotlib, seaborn, thop            -> empty dummy modules (plotting / flop counting only)
  timm.{models.,}layers                -> DropPath=Identity, trunc_normal_ (classes outside the path)
  efficientnet_pytorch.model           -> MemoryEfficientSwish=SiLU (classes outside the path)
  mmcv.ops.ModulatedDeformConv2d       -> torchvision.ops.deform_conv2d (same call the reference's
                                          own DCNv2 class makes, ultralytics/nn/modules/head.py:1373-1385)
  mmcv.cnn.build_norm_layer            -> ('gn', GroupNorm(num_groups, c))
mmcv is un-pinned and absent: DCN parity is "unpinned" beyond torchvision's operator.
"""
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """/root/reference in the build container; oracle/_ref (the copy made by oracle/build_ref.py, git-ignored) on the GPU box"""
    for cand in (os.environ.get("YAD_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "ultralytics")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


class _Anything:
    """Attribute sink: any attribute access / call returns another sink."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


def _dummy(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    m.__dict__.update(attrs)

    def _getattr(attr):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Anything()

    m.__getattr__ = _getattr
    sys.modules[name] = m
    return m


def install():
    """Install the shims and put the reference on sys.path. Idempotent."""
    if getattr(install, "_done", False):
        return
    import torch
    import torch.nn as nn

    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/yad_ultralytics_cfg")
    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    # F8: `import ultralytics` pins OMP_NUM_THREADS=1 when it is unset.
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.font_manager",
                 "matplotlib.ticker", "seaborn", "thop"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _dummy(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    try:
        import timm  # noqa: F401
    except Exception:
        layers = dict(DropPath=nn.Identity, trunc_normal_=nn.init.trunc_normal_, to_2tuple=lambda x: (x, x))
        _dummy("timm")
        _dummy("timm.layers", **layers)
        _dummy("timm.models")
        _dummy("timm.models.layers", **layers)
        _dummy("timm.models.registry", register_model=lambda f: f)
    try:
        import efficientnet_pytorch  # noqa: F401
    except Exception:
        _dummy("efficientnet_pytorch")
        _dummy("efficientnet_pytorch.model", MemoryEfficientSwish=nn.SiLU)

    try:
        import mmcv  # noqa: F401
    except Exception:
        import torchvision

        class ModulatedDeformConv2d(nn.Module):
            def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                         groups=1, deform_groups=1, bias=True):
                super().__init__()
                k = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
                self.stride, self.padding, self.dilation = stride, padding, dilation
                self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *k))
                self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
                n = in_channels * k[0] * k[1]
                self.weight.data.uniform_(-1.0 / n ** 0.5, 1.0 / n ** 0.5)

            def forward(self, x, offset, mask):
                return torchvision.ops.deform_conv2d(x, offset, self.weight, self.bias, self.stride,
                                                     self.padding, self.dilation, mask)

        # picklable under the name a real mmcv install would give it (reference checkpoints pickle the module tree)
        ModulatedDeformConv2d.__module__, ModulatedDeformConv2d.__qualname__ = "mmcv.ops", "ModulatedDeformConv2d"

        def build_norm_layer(cfg, num_features, postfix=""):
            assert cfg.get("type") == "GN"
            return "gn", nn.GroupNorm(cfg.get("num_groups", 16), num_features)

        _dummy("mmcv")
        _dummy("mmcv.ops", ModulatedDeformConv2d=ModulatedDeformConv2d)
        _dummy("mmcv.cnn", build_norm_layer=build_norm_layer)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    install._done = True


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ultralytics"))
