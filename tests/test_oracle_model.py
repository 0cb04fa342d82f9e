"""CPU: the oracle restatement of the model forward (oracle/model.py) against golden fixtures produced by the live reference
(oracle/gen_golden.py).  fp32 on both sides; tolerance 1e-4 relative to the layer's mean magnitude (BN folding / op order)."""
import json
import os

import numpy as np
import torch

from oracle import model as om
from oracle import synth
from oracle.cases import sample_positions

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_state_dict_is_reproducible():
    ck = json.load(open(os.path.join(GOLD, "state_checksum.json")))
    sd = synth.make_state_dict_np(seed=ck["seed"])
    assert len(sd) == ck["n_keys"] == 541
    assert synth.state_checksum(sd) == ck["checksum"]
    assert sum(v.size for k, v in sd.items() if v.dtype == np.float32 and "running" not in k) == 4098193


def _check_layers(g, ys, rtol=2e-4):
    for i in range(33):
        o = ys[i].numpy().reshape(-1)
        assert tuple(g[f"L{i}_shape"]) == tuple(ys[i].shape), i
        ref = g[f"L{i}_samples"]
        got = o[sample_positions(o.size)]
        scale = float(g[f"L{i}_absmean"]) + 1e-6
        err = np.abs(got - ref).max() / scale
        assert err < rtol, f"layer {i}: rel err {err}"


def test_forward_160_matches_reference(gold, state_dict):
    g = gold("model_160.npz")
    img = torch.from_numpy(synth.make_images(2, 160, 160, seed=2))
    (y, feats), ys = om.forward(state_dict, img, return_layers=True)
    _check_layers(g, ys)
    for i, f in enumerate(feats):
        np.testing.assert_allclose(f.numpy(), g[f"feat{i}"], rtol=0, atol=2e-4 * float(np.abs(g[f"feat{i}"]).mean()) + 1e-5)
    np.testing.assert_allclose(y.numpy()[:, 4:], g["y"][:, 4:], rtol=0, atol=2e-5)
    np.testing.assert_allclose(y.numpy()[:, :4], g["y"][:, :4], rtol=1e-4, atol=1e-3)


def test_forward_640_matches_reference(gold, state_dict):
    g = gold("model_640.npz")
    img = torch.from_numpy(synth.make_images(1, 640, 640, seed=2))
    (y, feats), ys = om.forward(state_dict, img, return_layers=True)
    _check_layers(g, ys)
    np.testing.assert_allclose(y.numpy()[:, 4:, ::7], g["y_sub"][:, 4:], rtol=0, atol=2e-5)
    np.testing.assert_allclose(y.numpy()[:, :4, ::7], g["y_sub"][:, :4], rtol=1e-4, atol=2e-3)
