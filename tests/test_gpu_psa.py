"""GPU parity of the stock position-sensitive attention block (nn/modules/block.py:874-964, 1010-1049) through the
C ABI: the C2PSA nn.Module mirror (yad_conv2d with the re-ordered / zero-padded qkv rows, yad_mha, yad_dwconv for pe(v) + the attention output,
residuals in the conv epilogues) against the fixtures written by the live reference and against the oracle.
Tolerances: fp32 build <= 1e-3 relative (north_star).  bf16 storage: each PSABlock is 7 materialised tensors deep with a softmax in the middle;
stated tolerance per stacked block: max error <= 6e-2 and mean absolute error <= 1e-2, both relative to the mean output magnitude, against the
oracle evaluated on the same bf16-rounded input."""
import json
import os

import pytest
import torch

from conftest import GOLD
from oracle import psa as op
from oracle.mona import make_block_state, make_input
from util_gpu import DEV, DTYPES, rel_err
from yolo_ad_refine_b200.modules import C2PSA

pytestmark = pytest.mark.gpu


def _case(name):
    spec = json.load(open(os.path.join(GOLD, "psa_block_spec.json")))[name]
    c, nb, n, h, w, seed = op.PSA_CASES[name]
    return make_block_state(spec, seed), make_input(c, n, h, w, seed), nb, c


@pytest.mark.parametrize("dtype", DTYPES, ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", list(op.PSA_CASES))
def test_c2psa_matches_reference_golden(gold, name, dtype):
    sd, x, nb, c = _case(name)
    m = C2PSA(c, c, nb).eval()
    m.load_state_dict(sd, strict=True)
    y = m(x.to(DEV).to(dtype)).float().cpu()
    want = op.c2psa({"m." + k: v for k, v in sd.items()}, x.to(dtype).float(), nb)
    if dtype == torch.float32:
        assert rel_err(y, want) < 1e-3
        assert rel_err(y[:, ::4], gold("psa_block.npz")[name]) < 1e-3  # the live reference's output
    else:
        assert rel_err(y, want) < 6e-2 * nb
        assert float((y - want).abs().mean() / want.abs().mean()) < 1e-2 * nb


def test_c2psa_batch_64_layer10_geometry():
    """batch 64 at the layer-10 geometry (256 channels, 20x20): every image equals the same image run alone (no cross-image leakage in yad_mha)"""
    sd, x, nb, c = _case("c256_n1_20")
    m = C2PSA(c, c, nb).eval()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(5)
    xb = torch.randn(64, c, 20, 20, generator=g).to(DEV).bfloat16()
    y = m(xb).float().cpu()
    y1 = m(xb[17:18].contiguous()).float().cpu()
    assert torch.equal(y[17:18], y1)


# ---- C2SFA (nn/modules/block.py:2358-2373): GroupNorm + depthwise + 1x1 processors, SE gate, GELU FFN ---------------------------------------------
def _sfa_case(name):
    spec = json.load(open(os.path.join(GOLD, "sfa_block_spec.json")))[name]
    c, nb, n, h, w, seed = op.SFA_CASES[name]
    return op.sfa_state(spec, seed), make_input(c, n, h, w, seed), nb, c


@pytest.mark.parametrize("dtype", DTYPES, ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", list(op.SFA_CASES))
def test_c2sfa_matches_reference_golden(gold, name, dtype):
    """bf16: each ProgressiveTSSA_Fusion0 is 10 materialised tensors deep; stated tolerance per stacked block: max 6e-2 / mean 1e-2 (relative to the
    mean output magnitude) against the oracle on the bf16-rounded input"""
    from yolo_ad_refine_b200.modules import C2SFA
    sd, x, nb, c = _sfa_case(name)
    m = C2SFA(c, c, nb).eval()
    m.load_state_dict(sd, strict=True)
    y = m(x.to(DEV).to(dtype)).float().cpu()
    want = op.c2sfa({"m." + k: v for k, v in sd.items()}, x.to(dtype).float(), nb)
    if dtype == torch.float32:
        assert rel_err(y, want) < 1e-3
        assert rel_err(y[:, ::4], gold("sfa_block.npz")[name]) < 1e-3  # the live reference's output
    else:
        assert rel_err(y, want) < 6e-2 * nb
        assert float((y - want).abs().mean() / want.abs().mean()) < 1e-2 * nb
