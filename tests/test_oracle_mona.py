"""CPU: oracle/mona.py (Mona adapter restated from a state dict) against the fixtures written by the live reference (oracle/gen_golden.py mona), and
the nn.Module mirror's state-dict keys against the reference's."""
import numpy as np
import pytest
import torch

from oracle import mona as om


def _sub(y, h, w):
    return y if h * w < 100 else y[:, :, ::2, ::2]


@pytest.mark.parametrize("name", list(om.MONA_CASES))
def test_mona_oracle_matches_reference(gold, name):
    g = gold("mona.npz")
    c, n, h, w, seed = om.MONA_CASES[name]
    y = om.mona_forward(om.make_state(c, seed), om.make_input(c, n, h, w, seed)).numpy()
    np.testing.assert_allclose(_sub(y, h, w), g[name], rtol=2e-5, atol=2e-5)


def test_mona_module_state_dict_keys(gold):
    from yolo_ad_refine_b200.modules import Mona
    g = gold("mona.npz")
    m = Mona(128)
    assert sorted(m.state_dict().keys()) == list(g["c128_20_keys"])
    ref = om.make_state(128, 1)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.items()}
    m.load_state_dict(ref, strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.eval()(torch.zeros(1, 128, 8, 8))  # no CPU fallback


def _block_case(name):
    import json, os
    from conftest import GOLD
    spec = json.load(open(os.path.join(GOLD, "mona_block_spec.json")))[name]
    c, nb, n, h, w, seed = om.BLOCK_CASES[name]
    return om.make_block_state(spec, seed), om.make_input(c, n, h, w, seed), nb, spec


@pytest.mark.parametrize("name", list(om.BLOCK_CASES))
def test_mona_block_oracle_matches_reference(gold, name):
    sd, x, nb, _ = _block_case(name)
    y = om.c2tssa_dyt_mona_edffn({"m." + k: v for k, v in sd.items()}, x, nb).numpy()
    np.testing.assert_allclose(y[:, ::4], gold("mona_block.npz")[name], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["c256_n1_20", "c256_n2_ragged"])
def test_mona_block_module_state_dict_keys(name):
    from yolo_ad_refine_b200.modules import C2TSSA_DYT_Mona_EDFFN
    sd, _, nb, spec = _block_case(name)
    m = C2TSSA_DYT_Mona_EDFFN(256, 256, nb)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == spec
    m.load_state_dict(sd, strict=True)
