"""CPU: the training graph (yolo_ad_refine_b200/training.py + trainer pieces) traces end to end with consistent shapes and every libyad call
it would issue matches the C ABI's arity and argument types (include/yad.h via _lib.SIGNATURES).  No kernels run: ops._call is replaced by a
signature checker -- this is host-logic coverage, the numerics are checked on the GPU (tests/test_gpu_train_step.py)."""
import ctypes as C

import pytest
import torch

from yolo_ad_refine_b200 import _lib, ops
from yolo_ad_refine_b200 import training as T
from yolo_ad_refine_b200.train_params import TrainParams, is_frozen, optimizer_group


@pytest.fixture()
def traced(monkeypatch, state_dict):
    calls = []

    def fake_call(name, *args, meta=None):
        res, argtypes = _lib.SIGNATURES[name]
        assert len(args) == len(argtypes), f"{name}: {len(args)} arguments, the C ABI takes {len(argtypes)}"
        for i, (a, t) in enumerate(zip(args, argtypes)):
            try:
                t.from_param(a)
            except Exception as e:  # noqa: BLE001
                raise AssertionError(f"{name}: argument {i} ({a!r}) does not convert to {t}: {e}")
        calls.append(name)

    monkeypatch.setattr(ops, "_call", fake_call)
    monkeypatch.setattr(ops, "stream_ptr", lambda: C.c_void_p(0))
    tp = TrainParams(state_dict, torch.float32, "cpu")
    return tp, calls


def test_training_graph_traces(traced):
    tp, calls = traced
    g = T.Graph(tp, conv_impl=1)
    img = torch.zeros(2, 3, 96, 160)
    outs, layers = T.forward_model(g, img)
    assert [(o.h, o.w, o.c) for o in outs] == [(12, 20, 144), (6, 10, 144), (3, 5, 144)]
    n_fwd = len(calls)
    for o in outs:
        g.mark(o)
    g.backward()
    assert len(calls) > n_fwd + 400
    tp.pack()
    tp.unpack_grads()
    # every trainable conv weight has a gradient-bearing kernel layout
    packed = {k for (k, kind) in tp._packed}
    for k in tp.keys:
        if k.endswith(".weight") and len(tp.shape[k]) == 4 and tp.shape[k][1] > 1 and not is_frozen(k) and "importance_gate" not in k and "la_conv" not in k:
            assert k in packed, k  # (the tiny gate MLPs read the master fp32 weights directly)
    assert {"yad_conv_wgrad", "yad_norm_bwd", "yad_mlca_bwd", "yad_mha_bwd", "yad_tssa_bwd", "yad_deform_col_bwd", "yad_maxpool5_bwd",
            "yad_patch_filter_bwd", "yad_adt_bwd", "yad_rowcol_gate_bwd", "yad_fusion_weights"} <= set(calls)


def test_optimizer_groups_match_reference(gold, state_dict):
    import json
    import os
    from conftest import GOLD
    ref = json.load(open(os.path.join(GOLD, "optimizer_groups.json")))
    for k, grp in ref.items():
        if k in state_dict and not is_frozen(k):
            assert optimizer_group(k) == grp, k


def test_checkpoint_roundtrip(state_dict):
    """flat arenas + counters survive checkpoint() / load_checkpoint(); state_dict() views follow the reference's keys"""
    tp = TrainParams(state_dict, torch.float32, "cpu")
    tp.flat.add_(0.5)
    tp.mom.fill_(0.25)
    tp.steps, tp.ema_updates = 7, 7
    ck = tp.checkpoint()
    tp2 = TrainParams(state_dict, torch.float32, "cpu")
    tp2.load_checkpoint(ck)
    assert torch.equal(tp2.flat, tp.flat) and torch.equal(tp2.mom, tp.mom) and tp2.steps == 7 and tp2.ema_updates == 7
    sd = tp2.state_dict()
    assert set(k for k in state_dict if state_dict[k].dtype.is_floating_point) <= set(sd)
    k = "model.0.conv.weight"
    assert torch.allclose(sd[k], state_dict[k].float() + 0.5)
