"""CPU: the nn.Module mirrors (yolo_ad_refine_b200/modules.py) expose exactly the reference's state-dict keys and shapes (so reference
checkpoints load and parse_model can build them by name), and the plugin rebinding covers the names SURVEY.md section 8b lists."""
import os

import pytest
import torch

from yolo_ad_refine_b200 import modules as M
from yolo_ad_refine_b200 import synth

REF = "/root/reference/ultralytics"


def test_state_dict_keys_match_reference_spec():
    model = M.YoloADRefine(nc=80)
    sd = model.state_dict()
    spec = {k: (tuple(s), d) for k, s, d in synth.load_spec()}
    assert set(sd) == set(spec), (sorted(set(sd) - set(spec))[:10], sorted(set(spec) - set(sd))[:10])
    for k, v in sd.items():
        assert tuple(v.shape) == spec[k][0], (k, tuple(v.shape), spec[k][0])
        assert str(v.dtype) == spec[k][1], (k, v.dtype, spec[k][1])
    assert list(sd) == [k for k, _, _ in synth.load_spec()]  # same registration order as the reference
    model.load_state_dict(synth.make_state_dict(seed=1), strict=True)
    assert sum(p.numel() for p in model.parameters()) == 4098193


def test_constructor_signatures_and_head_attributes():
    h = M.AYHead(80, (128, 128, 128))
    assert (h.nc, h.nl, h.reg_max, h.no) == (80, 3, 16, 144)
    assert h.stride.tolist() == [8.0, 16.0, 32.0]
    for attr in ("dynamic", "export", "shape", "anchors", "strides", "format", "bias_init"):
        assert hasattr(h, attr)
    assert M.Conv(3, 16, 3, 2).conv.stride == (2, 2) and isinstance(M.Conv.default_act, torch.nn.SiLU)
    assert M.C2PTSSA is M.C2ProgressiveTSSA_Fusion
    with pytest.raises(NotImplementedError):
        M.Fusion([128, 128], "concat")
    with pytest.raises(NotImplementedError):
        M.Conv(8, 8).train()(torch.zeros(1, 8, 4, 4))  # training forward is not built; must fail loudly, not fall back


def test_cpu_tensors_are_rejected_not_silently_computed():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.Conv(8, 8).eval()(torch.zeros(1, 8, 4, 4))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_plugin_rebinds_reference_names_and_parse_model_builds_mirrors():
    from oracle import ref_shims
    ref_shims.install()
    import ultralytics.nn.tasks as tasks
    import ultralytics.utils.ops as uops
    from yolo_ad_refine_b200 import plugin
    saved = {n: getattr(tasks, n) for n in plugin.BLOCKS + plugin.HEADS + ["v8DetectionLoss"] if hasattr(tasks, n)}
    saved_nms = uops.non_max_suppression
    try:
        done = plugin.install()
        assert "nn.tasks.C3k2_MLCA" in done and "utils.ops.non_max_suppression" in done and "nn.tasks.AYHead" in done
        assert "nn.tasks.C2TSSA_DYT_Mona_EDFFN" in done and "nn.modules.mona.Mona" in done and tasks.C2TSSA_DYT_Mona_EDFFN is M.C2TSSA_DYT_Mona_EDFFN
        m = tasks.DetectionModel("/root/reference/z-yaml/yolo11-701-YOLO-AD-Refine.yaml", ch=3, nc=80, verbose=False)
        plugin.convert_model(m)
        assert type(m.model[6]) is M.C3k2_MLCA and type(m.model[33]) is M.AYHead and type(m.model[13]) is M.YadConvTranspose2d
        spec = {k: tuple(s) for k, s, _ in synth.load_spec()}
        sd = m.state_dict()
        assert set(sd) == set(spec) and all(tuple(v.shape) == spec[k] for k, v in sd.items())
    finally:
        for n, o in saved.items():
            setattr(tasks, n, o)
        uops.non_max_suppression = saved_nms
