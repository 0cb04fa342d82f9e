"""GPU parity of the Mona adapter (nn/modules/mona.py:36-64) through the C ABI: yad_ln_mix alone against a torch fp32
LayerNorm, and the fused five-launch composition (yad_ln_mix, yad_conv2d, yad_dwconv with the merged 7x7 kernel, yad_conv2d + GELU with W + I,
yad_conv2d + residual) through the nn.Module mirror against the fixtures written by the live reference and against the oracle.
Tolerances: fp32 build <= 1e-3 relative (north_star).  bf16 storage (8 mantissa bits = 2^-9 relative per materialised tensor, five tensors deep,
worst case over ~10^5 outputs): max element-wise relative error <= 4e-2 and mean absolute error <= 5e-3 of the mean magnitude, against the oracle
evaluated on the same bf16-rounded input."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mona as om
from util_gpu import DEV, DTYPES, from_act, rel_err, to_act
from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.modules import Mona
from yolo_ad_refine_b200.ops import Act

pytestmark = pytest.mark.gpu
BF16_MAX, BF16_MEAN = 4e-2, 5e-3


def check(y, want, dtype):
    if dtype == torch.float32:
        assert rel_err(y, want) < 1e-3
    else:
        assert rel_err(y, want) < BF16_MAX
        assert float((y - want).abs().mean() / want.abs().mean()) < BF16_MEAN


@pytest.mark.parametrize("dtype", DTYPES, ids=["fp32", "bf16"])
@pytest.mark.parametrize("c,n,h,w", [(128, 2, 20, 20), (64, 1, 7, 9), (8, 3, 5, 5), (24, 1, 33, 2), (256, 2, 12, 12), (1024, 1, 3, 5), (320, 1, 9, 4)])
def test_ln_mix_matches_torch_layer_norm(dtype, c, n, h, w):
    g = torch.Generator().manual_seed(c + h)
    x = torch.randn(n, c, h, w, generator=g) * 2 + 0.5
    wt, b, ga, gx = (torch.randn(c, generator=g) for _ in range(4))
    xa = to_act(x, dtype)
    xr = from_act(xa)  # the values the kernel actually reads (bf16-rounded in the bf16 build)
    ref = F.layer_norm(xr.permute(0, 2, 3, 1), (c,), wt, b, 1e-5).permute(0, 3, 1, 2) * ga[None, :, None, None] + xr * gx[None, :, None, None]
    y = ops.ln_mix(xa, wt.to(DEV), b.to(DEV), ga.to(DEV), gx.to(DEV), 1e-5, Act.empty(n, h, w, c, dtype, DEV))
    assert rel_err(from_act(y), ref) < (1e-5 if dtype == torch.float32 else 6e-3)  # bf16: one output rounding


@pytest.mark.parametrize("dtype", DTYPES, ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", list(om.MONA_CASES))
def test_mona_module_matches_reference_golden(gold, name, dtype):
    g = gold("mona.npz")
    c, n, h, w, seed = om.MONA_CASES[name]
    sd, x = om.make_state(c, seed), om.make_input(c, n, h, w, seed)
    m = Mona(c).eval()
    m.load_state_dict(sd, strict=True)
    before = ops.LAUNCHES
    y = m(x.to(DEV).to(dtype)).float().cpu()
    assert ops.LAUNCHES - before == 5
    assert tuple(y.shape) == (n, c, h, w)
    check(y, om.mona_forward(sd, x.to(dtype).float()), dtype)
    if dtype == torch.float32:  # the live reference's output on the same fp32 input
        sub = y if h * w < 100 else y[:, :, ::2, ::2]
        assert rel_err(sub, g[name]) < 1e-3


def test_mona_parameter_refresh_and_training_mode():
    sd = om.make_state(64, 5)
    m = Mona(64)
    m.load_state_dict(sd)
    with pytest.raises(NotImplementedError):
        m.train()(torch.zeros(1, 64, 8, 8, device=DEV))
    m.eval()
    x = om.make_input(64, 1, 8, 8, 5)
    y0 = m(x.to(DEV)).cpu()
    sd2 = {k: v * 1.5 for k, v in sd.items()}
    m.load_state_dict(sd2)  # prepared weights are dropped on load
    y1 = m(x.to(DEV)).cpu()
    assert rel_err(y1, om.mona_forward(sd2, x)) < 1e-3 and rel_err(y0, om.mona_forward(sd, x)) < 1e-3


def test_mona_full_size_batch64_residual_property():
    """batch 64 at the layer-10 geometry in bf16: with project2 zeroed the adapter is the identity (bit for bit), and the full output stays within the
    bf16 tolerance of the oracle on a sampled image"""
    c, n, h, w = 128, 64, 20, 20
    sd = om.make_state(c, 9)
    x = torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(3)).bfloat16()
    m = Mona(c).eval()
    m.load_state_dict({**sd, "project2.weight": torch.zeros_like(sd["project2.weight"]), "project2.bias": torch.zeros_like(sd["project2.bias"])})
    assert torch.equal(m(x.to(DEV)).cpu(), x)
    m.load_state_dict(sd)
    y = m(x.to(DEV)).float().cpu()
    check(y[17:18], om.mona_forward(sd, x[17:18].float()), torch.bfloat16)


# ---- AttentionTSSA kernel and the C2TSSA_DYT_Mona_EDFFN block ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES, ids=["fp32", "bf16"])
@pytest.mark.parametrize("c,heads,n,h,w", [(128, 2, 2, 20, 20), (64, 1, 1, 7, 9), (256, 4, 3, 5, 5), (128, 2, 1, 80, 80), (512, 8, 1, 3, 3), (128, 16, 1, 6, 6)])
def test_attention_tssa_kernel_matches_oracle(dtype, c, heads, n, h, w):
    g = torch.Generator().manual_seed(c + heads + h)
    x = torch.randn(n, c, h, w, generator=g)
    xa = to_act(x, dtype)
    xr = from_act(xa)
    temp = 0.5 + torch.rand(heads, 1, generator=g)
    sd = {"a.qkv.weight": torch.eye(c), "a.temp": temp, "a.to_out.0.weight": torch.eye(c), "a.to_out.0.bias": torch.zeros(c)}
    ref = om.attention_tssa(sd, "a", xr.flatten(2).permute(0, 2, 1), heads).permute(0, 2, 1).reshape(n, c, h, w)
    y = ops.attention_tssa(xa, temp.reshape(-1).to(DEV), heads, Act.empty(n, h, w, c, dtype, DEV))
    assert rel_err(from_act(y), ref) < (2e-5 if dtype == torch.float32 else 6e-3)


def test_attention_tssa_argument_errors():
    x = Act.empty(1, 4, 4, 72, torch.float32, DEV)
    with pytest.raises(RuntimeError, match="head_dim"):
        ops.attention_tssa(x, torch.ones(5, device=DEV), 5, Act.empty(1, 4, 4, 72, torch.float32, DEV))
    x = Act.empty(1, 300, 300, 128, torch.float32, DEV)
    with pytest.raises(RuntimeError, match="shared memory"):
        ops.attention_tssa(x, torch.ones(2, device=DEV), 2, Act.empty(1, 300, 300, 128, torch.float32, DEV))


def _block_case(name):
    import json
    import os
    from conftest import GOLD
    spec = json.load(open(os.path.join(GOLD, "mona_block_spec.json")))[name]
    c, nb, n, h, w, seed = om.BLOCK_CASES[name]
    return om.make_block_state(spec, seed), om.make_input(c, n, h, w, seed), nb, c


@pytest.mark.parametrize("dtype", DTYPES, ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", list(om.BLOCK_CASES))
def test_c2tssa_dyt_mona_edffn_matches_reference_golden(gold, name, dtype):
    """bf16: each TSSAlock_DYT_Mona_EDFFN is ~25 materialised tensors deep; stated tolerance per stacked block: max 8e-2 / mean 1e-2 (relative to
    the mean magnitude) against the oracle on the bf16-rounded input"""
    from yolo_ad_refine_b200.modules import C2TSSA_DYT_Mona_EDFFN
    sd, x, nb, c = _block_case(name)
    m = C2TSSA_DYT_Mona_EDFFN(c, c, nb).eval()
    m.load_state_dict(sd, strict=True)
    y = m(x.to(DEV).to(dtype)).float().cpu()
    want = om.c2tssa_dyt_mona_edffn({"m." + k: v for k, v in sd.items()}, x.to(dtype).float(), nb)
    if dtype == torch.float32:
        assert rel_err(y, want) < 1e-3
        assert rel_err(y[:, ::4], gold("mona_block.npz")[name]) < 1e-3  # the live reference's output
    else:
        assert rel_err(y, want) < 8e-2 * nb
        assert float((y - want).abs().mean() / want.abs().mean()) < 1e-2 * nb
