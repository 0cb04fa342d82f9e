"""GPU parity for the small-channel tcgen05 convolution (conv3_kernel in csrc/conv_v2.cu, impl=6): 3x3 stride-1 convolutions with 8 / 16 / 32
input and output channels issued straight on a TMA-staged no-swizzle patch (tap pairs per K = 16 step with per-MMA leading-byte offsets).  Against
torch fp32 on bf16-rounded operands; tolerance as tests/test_gpu_conv_v2.py (2e-2 of |ref| + mean|ref|: bf16 storage of the output)."""
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


@pytest.mark.parametrize("cin", [8, 16, 32])
@pytest.mark.parametrize("cout", [8, 16, 32])
@pytest.mark.parametrize("h,w,n", [(160, 160, 4), (83, 45, 20)])  # ragged tiles in both directions; >= 592 tiles of 16 x 8 pixels
def test_c3_bias_silu(cin, cout, h, w, n):
    g = torch.Generator().manual_seed(cin * 100 + cout)
    x = q(torch.randn(n, cin, h, w, generator=g))
    wt = q(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.silu(F.conv2d(x, wt, b, 1, 1))
    cw = pack_conv(wt, b, BF, DEV, 1)
    out = Act.empty(n, h, w, cw.cout, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out, bias=cw.b, kh=3, kw=3, stride=1, pad_h=1, pad_w=1, act=ops.ACT_SILU, impl=6)
    assert rel_err(from_act(out, cout), ref) < TOL


@pytest.mark.parametrize("cin,cout", [(16, 8), (8, 16), (32, 16), (16, 32)])
@pytest.mark.parametrize("act,fn", [(ops.ACT_NONE, lambda t: t), (ops.ACT_SILU, F.silu), (ops.ACT_RELU, F.relu)])
def test_c3_add_alpha_channel_slices(cin, cout, act, fn):
    """residual add, alpha, no bias, and channel-sliced views on both sides (the C3k2 bottleneck reads half of cv1's output and writes into the
    concat buffer): input = channels [8, 8 + cin) of a wider buffer, output = channels [16, 16 + cout) of a wider buffer"""
    g = torch.Generator().manual_seed(cin + 7 * cout + act)
    n, h, w = 12, 80, 80
    xw = q(torch.randn(n, cin + 24, h, w, generator=g))
    x = xw[:, 8:8 + cin]
    wt = q(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    add = q(torch.randn(n, cout, h, w, generator=g))
    ref = fn(F.conv2d(x, wt, None, 1, 1)) * 0.5 + add
    cw = pack_conv(wt, None, BF, DEV, 1)
    wide = Act.empty(n, h, w, cout + 32, BF, DEV)
    wide.buf.zero_()
    ops.conv2d(to_act(xw, BF).slice(8, cin), cw.w, wide.slice(16, cout), kh=3, kw=3, stride=1, pad_h=1, pad_w=1, act=act, alpha=0.5, add=to_act(add, BF),
               impl=6)
    got = from_act(wide, cout + 32)
    assert rel_err(got[:, 16:16 + cout], ref) < TOL
    assert float(got[:, :16].abs().max()) == 0.0 and float(got[:, 16 + cout:].abs().max()) == 0.0  # neighbours of the slice untouched


@pytest.mark.parametrize("cout", [8, 16, 32])
@pytest.mark.parametrize("h,w,n", [(320, 320, 3), (166, 90, 20)])
def test_c3_stride2_pixel_pair_planes(cout, h, w, n):
    """stride 2, 16 input channels (backbone layer 1): two row-parity planes of pixel pairs (SWIZZLE_64B), a tap = (plane, pair offset, 32-byte half);
    even maps whose output grid is ragged against the 16 x 8 tile"""
    g = torch.Generator().manual_seed(cout + h)
    x = q(torch.randn(n, 16, h, w, generator=g))
    wt = q(torch.randn(cout, 16, 3, 3, generator=g) / 12.0)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.silu(F.conv2d(x, wt, b, 2, 1))
    cw = pack_conv(wt, b, BF, DEV, 2)
    out = Act.empty(n, h // 2, w // 2, cw.cout, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out, bias=cw.b, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, act=ops.ACT_SILU, impl=6)
    assert rel_err(from_act(out, cout), ref) < TOL
