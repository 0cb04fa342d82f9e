"""GPU parity for the resident-weight tcgen05 convolution kernel (csrc/conv_v2.cu, impl=4): stride-1 1x1 and 3x3 convolutions against torch fp32 on
bf16-rounded operands -- the haloed-patch 3x3 path (tap-shifted UMMA descriptors), K tails, ragged tiles, persistent multi-tile CTAs, every fused
epilogue (bias / SiLU / sigmoid / runtime activation / alpha / per-image and per-pixel scales / mul / add / channel-sliced output) and the fused
GroupNorm statistics.  Tolerance: 2e-2 of (|ref| + mean|ref|) per element -- bf16 storage of the output (2^-9 per rounding) on exact-product fp32
accumulation."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util_gpu import DEV, from_act, rel_err, to_act
from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.ops import Act
from yolo_ad_refine_b200.weights import pack_conv

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
TOL = 2e-2


def q(x):
    return x.to(BF).float()


SHAPES = [  # cin, cout, k, hw, n
    (48, 64, 1, 20, 2), (128, 64, 3, 10, 2), (192, 128, 1, 7, 2), (64, 27, 3, 12, 2), (256, 256, 1, 5, 3), (128, 384, 1, 20, 2),
    (128, 512, 1, 10, 2), (64, 64, 3, 80, 5), (64, 64, 3, 13, 2), (96, 128, 1, 21, 2), (128, 64, 3, 20, 3), (32, 32, 3, 40, 2),
    (32, 64, 1, 16, 2), (16, 16, 1, 24, 2), (32, 64, 3, 40, 2), (64, 32, 3, 24, 2), (16, 32, 3, 33, 1), (64, 80, 1, 40, 2),
    (128, 128, 1, 80, 8), (64, 128, 3, 17, 2), (24, 40, 1, 9, 2), (40, 16, 3, 8, 1), (128, 256, 1, 20, 2), (64, 64, 1, 80, 4),
]


@pytest.mark.parametrize("cin,cout,k,hw,n", SHAPES)
def test_v2_bias_silu_add(cin, cout, k, hw, n):
    g = torch.Generator().manual_seed(cin * 1000 + cout + k)
    x = q(torch.randn(n, cin, hw, hw, generator=g))
    w = q(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    add = q(torch.randn(n, cout, hw, hw, generator=g))
    ref = F.silu(F.conv2d(x, w, b, 1, k // 2)) + add
    cw = pack_conv(w, b, BF, DEV, 1)
    out = Act.empty(n, hw, hw, cw.cout, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out, bias=cw.b, kh=k, kw=k, stride=1, pad_h=k // 2, pad_w=k // 2, act=ops.ACT_SILU, add=to_act(add, BF), impl=4)
    assert rel_err(from_act(out, cout), ref) < TOL


@pytest.mark.parametrize("cin,cout,k,hw,n", [(64, 64, 3, 80, 3), (128, 128, 1, 40, 4), (64, 32, 3, 20, 2), (128, 64, 1, 13, 2)])
def test_v2_plain_no_bias(cin, cout, k, hw, n):
    g = torch.Generator().manual_seed(cin + cout)
    x = q(torch.randn(n, cin, hw, hw, generator=g))
    w = q(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    ref = F.conv2d(x, w, None, 1, k // 2)
    cw = pack_conv(w, None, BF, DEV, 1)
    out = Act.empty(n, hw, hw, cw.cout, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out, kh=k, kw=k, stride=1, pad_h=k // 2, pad_w=k // 2, impl=4)
    assert rel_err(from_act(out, cout), ref) < TOL


@pytest.mark.parametrize("act,fn", [(ops.ACT_SIGMOID, torch.sigmoid), (ops.ACT_RELU, F.relu), (ops.ACT_GELU, F.gelu), (ops.ACT_HARDSWISH, F.hardswish)])
@pytest.mark.parametrize("k", [1, 3])
def test_v2_act_mul_add_alpha(act, fn, k):
    g = torch.Generator().manual_seed(11 + act)
    n, cin, cout, hw = 2, 128, 64, 20
    x = q(torch.randn(n, cin, hw, hw, generator=g))
    w = q(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    mul, add = q(torch.randn(n, cout, hw, hw, generator=g)), q(torch.randn(n, cout, hw, hw, generator=g))
    ref = fn(F.conv2d(x, w, b, 1, k // 2)) * 0.7 * mul + add
    cw = pack_conv(w, b, BF, DEV, 1)
    out = Act.empty(n, hw, hw, cw.cout, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out, bias=cw.b, kh=k, kw=k, pad_h=k // 2, pad_w=k // 2, act=act, alpha=0.7, mul=to_act(mul, BF), add=to_act(add, BF),
               impl=4)
    assert rel_err(from_act(out, cout), ref) < TOL


def test_v2_scales_and_concat_slice():
    g = torch.Generator().manual_seed(5)
    x = q(torch.randn(3, 64, 12, 9, generator=g))
    w = q(torch.randn(80, 64, 1, 1, generator=g) / 8)
    b = torch.randn(80, generator=g) * 0.1
    img_scale = torch.rand(3, generator=g) + 0.5
    pix = q(torch.rand(3, 8, 12, 9, generator=g))
    mul = q(torch.randn(3, 80, 12, 9, generator=g))
    ref = torch.sigmoid(F.conv2d(x, w) * img_scale.view(3, 1, 1, 1) * pix[:, :1] + b.view(1, -1, 1, 1)) * 0.7 * mul
    cw = pack_conv(w, b, BF, DEV)
    wide = Act.empty(3, 12, 9, 144, BF, DEV)
    wide.buf.zero_()
    ops.conv2d(to_act(x, BF), cw.w, wide.slice(64, 80), bias=cw.b, act=ops.ACT_SIGMOID, alpha=0.7, img_scale=img_scale.to(DEV),
               pix_scale=to_act(pix, BF), mul=to_act(mul, BF), impl=4)
    assert rel_err(from_act(wide)[:, 64:144], ref) < TOL
    assert float(from_act(wide)[:, :64].abs().max()) == 0.0


def test_v2_sliced_input_and_output_3x3():
    """input = a channel window of a wider buffer (C2f split), output = a window of the concat buffer; the untouched channels stay untouched"""
    g = torch.Generator().manual_seed(9)
    n, hw = 2, 20
    xin = q(torch.randn(n, 128, hw, hw, generator=g))
    w = q(torch.randn(32, 64, 3, 3, generator=g) / 24)
    b = torch.randn(32, generator=g) * 0.1
    ref = F.silu(F.conv2d(xin[:, 64:128], w, b, 1, 1))
    cw = pack_conv(w, b, BF, DEV, 1)
    src = to_act(xin, BF)
    wide = Act.empty(n, hw, hw, 96, BF, DEV)
    wide.buf.fill_(3.0)
    ops.conv2d(src.slice(64, 64), cw.w, wide.slice(32, 32), bias=cw.b, kh=3, kw=3, pad_h=1, pad_w=1, act=ops.ACT_SILU, impl=4)
    got = from_act(wide)
    assert rel_err(got[:, 32:64], ref) < TOL
    assert float((got[:, :32] - 3.0).abs().max()) == 0.0 and float((got[:, 64:] - 3.0).abs().max()) == 0.0


@pytest.mark.parametrize("cin,cout,k,hw,n", [(128, 64, 3, 80, 2), (64, 64, 3, 40, 3), (128, 128, 1, 80, 2), (128, 128, 1, 20, 5), (64, 64, 3, 13, 2),
                                             (64, 256, 1, 10, 3), (64, 64, 1, 20, 3)])
def test_v2_fused_groupnorm_statistics(cin, cout, k, hw, n):
    """conv epilogue statistics double [n][16][2] (sum, sum of squares of the fp32 conv output) against torch, + img_scale on the 1x1 (TaskDecomposition)"""
    g = torch.Generator().manual_seed(cin + 3 * cout + hw)
    x = q(torch.randn(n, cin, hw, hw, generator=g))
    w = q(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    img_scale = (torch.rand(n, generator=g) + 0.5) if k == 1 else None
    ref = F.conv2d(x, w, None, 1, k // 2)
    if img_scale is not None:
        ref = ref * img_scale.view(n, 1, 1, 1)
    groups = 16
    rg = ref.double().reshape(n, groups, -1)
    ref_stats = torch.stack([rg.sum(-1), (rg * rg).sum(-1)], -1)
    cw = pack_conv(w, None, BF, DEV, 1)
    out = Act.empty(n, hw, hw, cw.cout, BF, DEV)
    stats = torch.full((n, groups, 2), 7.0, dtype=torch.float64, device=DEV)  # yad_conv2d zeroes it
    ops.conv2d(to_act(x, BF), cw.w, out, kh=k, kw=k, pad_h=k // 2, pad_w=k // 2, gn_stats=stats, gn_groups=groups,
               img_scale=None if img_scale is None else img_scale.to(DEV), impl=4)
    assert rel_err(from_act(out, cout), ref) < TOL
    got = stats.cpu()
    cnt = rg.shape[-1]
    # sums: absolute error against the scale sqrt(cnt) * rms; squares: relative
    rms = float(ref.pow(2).mean().sqrt())
    assert float((got[..., 0] - ref_stats[..., 0]).abs().max()) < 1e-3 * rms * cnt ** 0.5 + 1e-6 * cnt * rms
    assert float(((got[..., 1] - ref_stats[..., 1]).abs() / ref_stats[..., 1]).max()) < 1e-4


def test_v2_not_eligible_raises():
    x = Act.empty(1, 8, 8, 64, BF, DEV)
    w = torch.zeros(64, 9 * 64, dtype=BF, device=DEV)
    y = Act.empty(1, 4, 4, 64, BF, DEV)
    with pytest.raises(RuntimeError):
        ops.conv2d(x, w, y, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, impl=4)
