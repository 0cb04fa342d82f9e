"""CPU: oracle/psa.py (Attention / PSABlock / C2PSA restated from a state dict) against the fixtures written by the live reference
(oracle/gen_golden.py psa_block), and the nn.Module mirror's state-dict keys against the reference's."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import psa as op
from oracle.mona import make_block_state, make_input


def _case(name):
    spec = json.load(open(os.path.join(GOLD, "psa_block_spec.json")))[name]
    c, nb, n, h, w, seed = op.PSA_CASES[name]
    return make_block_state(spec, seed), make_input(c, n, h, w, seed), nb, c, spec


@pytest.mark.parametrize("name", list(op.PSA_CASES))
def test_psa_oracle_matches_reference(gold, name):
    sd, x, nb, _, _ = _case(name)
    y = op.c2psa({"m." + k: v for k, v in sd.items()}, x, nb).numpy()
    np.testing.assert_allclose(y[:, ::4], gold("psa_block.npz")[name], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["c256_n1_20", "c256_n2_ragged", "c512_n1_7"])
def test_c2psa_module_state_dict_keys(name):
    from yolo_ad_refine_b200.modules import C2PSA
    sd, _, nb, c, spec = _case(name)
    m = C2PSA(c, c, nb)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == spec
    m.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.eval()(torch.zeros(1, c, 8, 8))  # no CPU fallback


def test_plugin_lists_psa_blocks():
    from yolo_ad_refine_b200 import modules as M
    from yolo_ad_refine_b200 import plugin
    for name in ("C2PSA", "PSABlock", "Attention"):
        assert name in plugin.BLOCKS and hasattr(M, name)


def _sfa_case(name):
    spec = json.load(open(os.path.join(GOLD, "sfa_block_spec.json")))[name]
    c, nb, n, h, w, seed = op.SFA_CASES[name]
    return op.sfa_state(spec, seed), make_input(c, n, h, w, seed), nb, c, spec


@pytest.mark.parametrize("name", list(op.SFA_CASES))
def test_sfa_oracle_matches_reference(gold, name):
    sd, x, nb, _, _ = _sfa_case(name)
    y = op.c2sfa({"m." + k: v for k, v in sd.items()}, x, nb).numpy()
    np.testing.assert_allclose(y[:, ::4], gold("sfa_block.npz")[name], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", list(op.SFA_CASES))
def test_c2sfa_module_state_dict_keys(name):
    from yolo_ad_refine_b200.modules import C2SFA
    sd, _, nb, c, spec = _sfa_case(name)
    m = C2SFA(c, c, nb)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == spec
    m.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.eval()(torch.zeros(1, c, 8, 8))  # no CPU fallback
