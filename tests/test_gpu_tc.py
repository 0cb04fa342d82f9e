"""GPU parity for the tcgen05 / TMEM implicit-GEMM convolution (impl=2): first the bare UMMA GEMM self-test (descriptor encodings, swizzled
staging, TMEM read-back), then every convolution mode against torch fp32 on bf16-rounded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model as om
from util_gpu import DEV, from_act, rel_err, to_act
from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200._lib import check
from yolo_ad_refine_b200.ops import Act
from yolo_ad_refine_b200.weights import pack_conv

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
TOL = 2e-2


def q(x):
    return x.to(BF).float()


@pytest.mark.parametrize("m,n,k", [(128, 16, 64), (128, 64, 64), (256, 128, 128), (300, 80, 72), (1000, 256, 576), (128, 32, 16), (77, 8, 8)])
def test_umma_gemm_selftest(m, n, k):
    g = torch.Generator().manual_seed(m + n + k)
    a, b = q(torch.randn(m, k, generator=g)), q(torch.randn(n, k, generator=g))
    c = torch.full((m, n), float("nan"), device=DEV)
    ad, bd = a.to(DEV).to(BF).contiguous(), b.to(DEV).to(BF).contiguous()
    check(ops.lib().yad_tc_gemm_selftest(C.c_void_p(ad.data_ptr()), C.c_void_p(bd.data_ptr()), C.c_void_p(c.data_ptr()), m, n, k, ops.stream_ptr()),
          "selftest")
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t()
    err = float((c.cpu().double() - ref).abs().max() / ref.abs().mean())
    assert err < 1e-4, err  # exact products, fp32 accumulation


@pytest.mark.parametrize("cin,cout,k,s,hw", [(8, 16, 3, 2, 32), (16, 8, 3, 1, 20), (48, 64, 1, 1, 20), (128, 64, 3, 1, 10), (64, 32, 3, 2, 40),
                                             (192, 128, 1, 1, 7), (64, 27, 3, 1, 12), (32, 1, 3, 1, 9), (256, 256, 1, 1, 5), (128, 384, 1, 1, 20),
                                             (128, 512, 1, 1, 10), (64, 64, 3, 1, 80), (128, 128, 3, 1, 40), (64, 64, 3, 1, 13),
                                             (96, 128, 1, 1, 21), (128, 64, 3, 1, 20), (16, 32, 3, 2, 64), (64, 64, 3, 2, 32), (128, 256, 3, 2, 20),
                                             (32, 32, 3, 1, 40), (16, 8, 3, 1, 32), (8, 16, 3, 1, 32), (32, 64, 1, 1, 16), (16, 16, 1, 1, 24),
                                             (128, 128, 3, 2, 18), (8, 8, 3, 1, 45), (32, 16, 3, 2, 50), (32, 32, 3, 2, 22)])
@pytest.mark.parametrize("impl", [0, 2, 3])  # 0 = auto (single-pass mma.sync kernel for 3x3 with <= 32 channels, else as 2), 2 = tcgen05,
# TMA-fed where eligible, 3 = tcgen05 with thread-gathered operands
def test_conv_tc_bias_silu_add(cin, cout, k, s, hw, impl):
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    x = q(torch.randn(2, cin, hw, hw, generator=g))
    w = q(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ho = (hw + 2 * (k // 2) - k) // s + 1
    add = q(torch.randn(2, cout, ho, ho, generator=g))
    ref = F.silu(F.conv2d(x, w, b, s, k // 2)) + add
    cw = pack_conv(w, b, BF, DEV, s)
    out = Act.empty(2, ho, ho, cw.cout, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out, bias=cw.b, kh=k, kw=k, stride=s, pad_h=k // 2, pad_w=k // 2, act=ops.ACT_SILU, add=to_act(add, BF), impl=impl)
    assert rel_err(from_act(out, cout), ref) < TOL


def test_conv_tc_epilogue_scales_and_concat_slice():
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
               pix_scale=to_act(pix, BF), mul=to_act(mul, BF), impl=2)
    assert rel_err(from_act(wide)[:, 64:144], ref) < TOL
    assert float(from_act(wide)[:, :64].abs().max()) == 0.0


def test_conv_tc_transposed():
    g = torch.Generator().manual_seed(6)
    x = q(torch.randn(2, 32, 7, 9, generator=g))
    w = q(torch.randn(32, 24, 3, 3, generator=g) / 17)
    b = torch.randn(24, generator=g) * 0.1
    ref = F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1)
    cw = pack_conv(w, b, BF, DEV, transposed=True)
    out = Act.empty(2, 14, 18, 24, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out, bias=cw.b, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, mode=ops.CONV_TRANSPOSED, impl=2)
    assert rel_err(from_act(out), ref) < TOL


def test_conv_tc_deformable():
    g = torch.Generator().manual_seed(7)
    x = q(torch.randn(2, 64, 10, 12, generator=g))
    w = q(torch.randn(64, 64, 3, 3, generator=g) / 24)
    om_ = q(torch.randn(2, 27, 10, 12, generator=g) * 1.5)
    ref = om.deform_conv3x3(x, om_[:, :18], om_[:, 18:].sigmoid(), w)
    cw = pack_conv(w, None, BF, DEV)
    out = Act.empty(2, 10, 12, 64, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out, kh=3, kw=3, pad_h=1, pad_w=1, mode=ops.CONV_DEFORM, offmask=to_act(om_, BF), impl=2)
    assert rel_err(from_act(out), ref) < 3e-2  # the bilinear blend is rounded to bf16 before the MMA


def test_conv_tc_7x1():
    g = torch.Generator().manual_seed(8)
    v = q(torch.randn(4, 128, 20, generator=g))
    w = q(torch.randn(128, 128, 7, generator=g) / 30)
    b = torch.randn(128, generator=g) * 0.1
    ref = F.conv1d(v, w, b, padding=3)
    cw = pack_conv(w[:, :, :, None], b, BF, DEV)
    out = Act.empty(4, 20, 1, 128, BF, DEV)
    ops.conv2d(to_act(v[:, :, :, None], BF), cw.w, out, bias=cw.b, kh=7, kw=1, pad_h=3, pad_w=0, impl=2)
    assert rel_err(from_act(out)[:, :, :, 0], ref) < TOL


@pytest.mark.parametrize("dtype", [torch.float32, BF])
@pytest.mark.parametrize("n,h,w,cin,cout,with_add,act", [(2, 12, 9, 64, 128, True, ops.ACT_NONE), (3, 20, 20, 128, 128, True, ops.ACT_NONE),
                                                         (2, 7, 13, 32, 64, False, ops.ACT_NONE), (1, 40, 40, 128, 256, True, ops.ACT_SILU),
                                                         (5, 3, 50, 64, 80, True, ops.ACT_SIGMOID), (2, 16, 16, 64, 64, True, ops.ACT_GELU)])
def test_conv_separable_gate_epilogue(dtype, n, h, w, cin, cout, with_add, act):
    """yad_epilogue.gate_h / gate_w (ELA_HSFPN flag=False + Multiply folded into the lateral 1x1, z-yaml layers 15-18 / 22-25, block.py:1408-1424):
    against torch, and in bf16 bit-identical to the path that materialises the gate map with yad_rowcol_gate and passes it as `mul`"""
    g = torch.Generator().manual_seed(n * 100 + h + cout)
    rnd = (lambda t: q(t)) if dtype == BF else (lambda t: t)
    x = rnd(torch.randn(n, cin, h, w, generator=g))
    wt = rnd(torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    gh = rnd(torch.rand(n, cout, h, 1, generator=g))
    gw = rnd(torch.rand(n, cout, w, 1, generator=g))
    add = rnd(torch.randn(n, cout, h, w, generator=g))
    gate = rnd(gh * gw.permute(0, 1, 3, 2))
    acts = {ops.ACT_NONE: lambda t: t, ops.ACT_SILU: F.silu, ops.ACT_SIGMOID: torch.sigmoid, ops.ACT_GELU: F.gelu}
    ref = acts[act](F.conv2d(x, wt, b)) * gate + (add if with_add else 0)
    cw = pack_conv(wt, b, dtype, DEV)
    impl = 0 if dtype == BF else 1
    gha, gwa = to_act(gh, dtype), to_act(gw, dtype)
    out = Act.empty(n, h, w, cw.cout, dtype, DEV)
    ops.conv2d(to_act(x, dtype), cw.w, out, bias=cw.b, act=act, gate=(gha, gwa), add=to_act(add, dtype) if with_add else None, impl=impl)
    assert rel_err(from_act(out, cout), ref) < (TOL if dtype == BF else 1e-4)
    gmap = ops.rowcol_gate(None, gha, gwa, Act.empty(n, h, w, cw.cout, dtype, DEV))
    out2 = Act.empty(n, h, w, cw.cout, dtype, DEV)
    ops.conv2d(to_act(x, dtype), cw.w, out2, bias=cw.b, act=act, mul=gmap, add=to_act(add, dtype) if with_add else None, impl=impl)
    if dtype == BF:
        assert torch.equal(out.buf, out2.buf)
    else:
        assert rel_err(from_act(out, cout), from_act(out2, cout)) < 1e-6


def test_conv_separable_gate_rejected_outside_its_scope():
    x = to_act(torch.randn(1, 64, 8, 8), BF)
    cw = pack_conv(torch.randn(64, 64, 3, 3) / 24, None, BF, DEV)
    gh, gw = to_act(torch.rand(1, 64, 8, 1), BF), to_act(torch.rand(1, 64, 8, 1), BF)
    with pytest.raises(RuntimeError, match="separable gate"):
        ops.conv2d(x, cw.w, Act.empty(1, 8, 8, 64, BF, DEV), kh=3, kw=3, pad_h=1, pad_w=1, gate=(gh, gw))


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 80, 80, 64, 80), (3, 12, 9, 64, 80), (2, 20, 20, 64, 32), (1, 33, 7, 128, 64)])
def test_conv_v2_pix_scale_alpha_and_relu_variants(n, h, w, cin, cout):
    """the dedicated conv2_kernel variants for the head's cv3 (per-pixel scale only, head.py:1171-1173) and cls_prob_conv.0 (ReLU, head.py:1168)"""
    g = torch.Generator().manual_seed(cin + cout + h)
    x = q(torch.randn(n, cin, h, w, generator=g))
    wt = q(torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    pix = q(torch.rand(n, 8, h, w, generator=g))
    cw = pack_conv(wt, b, BF, DEV)
    out = Act.empty(n, h, w, cw.cout, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out, bias=cw.b, pix_scale=to_act(pix, BF), alpha=0.5)
    assert rel_err(from_act(out, cout), (F.conv2d(x, wt) * pix[:, :1] + b.view(1, -1, 1, 1)) * 0.5) < TOL
    ops.conv2d(to_act(x, BF), cw.w, out, bias=cw.b, act=ops.ACT_RELU)
    assert rel_err(from_act(out, cout), F.relu(F.conv2d(x, wt, b))) < TOL


@pytest.mark.parametrize("n,cin,cout,h,w", [(2, 32, 24, 7, 9), (3, 128, 128, 20, 20), (1, 64, 64, 17, 33), (2, 128, 64, 16, 8), (64, 128, 128, 10, 10)])
def test_conv_transposed_on_conv2_kernel(n, cin, cout, h, w):
    """nn.ConvTranspose2d(k3, s2, p1, op1) (z-yaml layers 13 / 20) as four output phases on conv2_kernel (impl 4 forces that path on small maps):
    run-time tap lists, patch origin at the tile, strided TMA store"""
    g = torch.Generator().manual_seed(cin + cout + h)
    x = q(torch.randn(n, cin, h, w, generator=g))
    wt = q(torch.randn(cin, cout, 3, 3, generator=g) / (cin * 2.25) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv_transpose2d(x, wt, b, stride=2, padding=1, output_padding=1)
    cw = pack_conv(wt, b, BF, DEV, transposed=True)
    out = Act.empty(n, 2 * h, 2 * w, cw.cout, BF, DEV)
    out.buf.fill_(float("nan"))
    ops.conv2d(to_act(x, BF), cw.w, out, bias=cw.b, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, mode=ops.CONV_TRANSPOSED, impl=4)
    assert rel_err(from_act(out, cout), ref) < TOL
    out5 = Act.empty(n, 2 * h, 2 * w, cw.cout, BF, DEV)
    ops.conv2d(to_act(x, BF), cw.w, out5, bias=cw.b, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, mode=ops.CONV_TRANSPOSED, impl=5)  # conv_tma_kernel
    assert rel_err(from_act(out, cout), from_act(out5, cout)) < 1e-2
