"""GPU parity: every libyad.so operator (called through the C ABI via yolo_ad_refine_b200.ops) against the fp32 oracle restatement
(oracle/model.py, torch-CPU) on the same seeded inputs.  Tolerances: util_gpu.tol (1e-3 fp32, 4e-2 bf16, relative to mean magnitude)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model as om
from util_gpu import DEV, DTYPES, from_act, rel_err, to_act, tol
from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.ops import Act
from yolo_ad_refine_b200.weights import edffn_spectral_matrix, pack_conv

pytestmark = pytest.mark.gpu


def q(x, dtype):
    """round inputs to the storage dtype so that both sides see the same values"""
    return x.to(dtype).float()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cin,cout,k,s,hw", [(8, 16, 3, 2, 32), (16, 8, 3, 1, 20), (48, 64, 1, 1, 20), (128, 64, 3, 1, 10), (64, 32, 3, 2, 40),
                                             (192, 128, 1, 1, 7), (64, 27, 3, 1, 12), (32, 1, 3, 1, 9), (256, 256, 1, 1, 5)])
def test_conv_bias_silu_add(dtype, cin, cout, k, s, hw):
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    x = q(torch.randn(2, cin, hw, hw, generator=g), dtype)
    w = q(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5, dtype)
    b = torch.randn(cout, generator=g) * 0.1
    ho = (hw + 2 * (k // 2) - k) // s + 1
    add = q(torch.randn(2, cout, ho, ho, generator=g), dtype)
    ref = F.silu(F.conv2d(x, w, b, s, k // 2)) + add
    cw = pack_conv(w, b, dtype, DEV, s)
    out = Act.empty(2, ho, ho, cw.cout, dtype, DEV)
    ops.conv2d(to_act(x, dtype), cw.w, out, bias=cw.b, kh=k, kw=k, stride=s, pad_h=k // 2, pad_w=k // 2, act=ops.ACT_SILU,
               add=to_act(add, dtype), impl=1)
    assert rel_err(from_act(out, cout), ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_epilogue_scales_and_concat_slice(dtype):
    g = torch.Generator().manual_seed(5)
    x = q(torch.randn(3, 64, 12, 9, generator=g), dtype)
    w = q(torch.randn(80, 64, 1, 1, generator=g) / 8, dtype)
    b = torch.randn(80, generator=g) * 0.1
    img_scale = torch.rand(3, generator=g) + 0.5
    pix = q(torch.rand(3, 8, 12, 9, generator=g), dtype)
    mul = q(torch.randn(3, 80, 12, 9, generator=g), dtype)
    ref = torch.sigmoid(F.conv2d(x, w) * img_scale.view(3, 1, 1, 1) * pix[:, :1] + b.view(1, -1, 1, 1)) * 0.7 * mul
    cw = pack_conv(w, b, dtype, DEV)
    wide = Act.empty(3, 12, 9, 144, dtype, DEV)
    wide.buf.zero_()
    ops.conv2d(to_act(x, dtype), cw.w, wide.slice(64, 80), bias=cw.b, act=ops.ACT_SIGMOID, alpha=0.7, img_scale=img_scale.to(DEV),
               pix_scale=to_act(pix, dtype), mul=to_act(mul, dtype), impl=1)
    assert rel_err(from_act(wide)[:, 64:144], ref) < tol(dtype)
    assert float(from_act(wide)[:, :64].abs().max()) == 0.0  # the neighbouring channel window is untouched


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_transposed(dtype):
    g = torch.Generator().manual_seed(6)
    x = q(torch.randn(2, 32, 7, 9, generator=g), dtype)
    w = q(torch.randn(32, 24, 3, 3, generator=g) / 17, dtype)  # (cin, cout, kh, kw)
    b = torch.randn(24, generator=g) * 0.1
    ref = F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1)
    cw = pack_conv(w, b, dtype, DEV, transposed=True)
    out = Act.empty(2, 14, 18, 24, dtype, DEV)
    ops.conv2d(to_act(x, dtype), cw.w, out, bias=cw.b, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, mode=ops.CONV_TRANSPOSED, impl=1)
    assert rel_err(from_act(out), ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_deformable(dtype):
    g = torch.Generator().manual_seed(7)
    x = q(torch.randn(2, 64, 10, 12, generator=g), dtype)
    w = q(torch.randn(64, 64, 3, 3, generator=g) / 24, dtype)
    om_ = q(torch.randn(2, 27, 10, 12, generator=g) * 1.5, dtype)
    ref = om.deform_conv3x3(x, om_[:, :18], om_[:, 18:].sigmoid(), w)
    cw = pack_conv(w, None, dtype, DEV)
    out = Act.empty(2, 10, 12, 64, dtype, DEV)
    ops.conv2d(to_act(x, dtype), cw.w, out, kh=3, kw=3, pad_h=1, pad_w=1, mode=ops.CONV_DEFORM, offmask=to_act(om_, dtype), impl=1)
    assert rel_err(from_act(out), ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_7x1_ela_branch(dtype):
    g = torch.Generator().manual_seed(8)
    v = q(torch.randn(4, 32, 20, generator=g), dtype)  # (n, c, L)
    w = q(torch.randn(32, 32, 7, generator=g) / 15, dtype)
    b = torch.randn(32, generator=g) * 0.1
    ref = F.conv1d(v, w, b, padding=3)
    cw = pack_conv(w[:, :, :, None], b, dtype, DEV)
    out = Act.empty(4, 20, 1, 32, dtype, DEV)
    ops.conv2d(to_act(v[:, :, :, None], dtype), cw.w, out, bias=cw.b, kh=7, kw=1, pad_h=3, pad_w=0, impl=1)
    assert rel_err(from_act(out)[:, :, :, 0], ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k,gate", [(3, False), (7, False), (3, True)])
def test_dwconv(dtype, k, gate):
    g = torch.Generator().manual_seed(9 + k)
    c = 64 if gate else 32
    x = q(torch.randn(2, c, 11, 13, generator=g), dtype)
    w = torch.randn(c, 1, k, k, generator=g) / k
    b = torch.randn(c, generator=g) * 0.1
    y = F.conv2d(x, w, b, 1, k // 2, groups=c)
    wk = w.reshape(c, k * k).t().contiguous().to(DEV)
    if gate:
        ref = F.gelu(y[:, :32]) * y[:, 32:]
        out = ops.dwconv(to_act(x, dtype), wk, Act.empty(2, 11, 13, 32, dtype, DEV), bias=b.to(DEV), k=k, gate_split=32)
    else:
        sc, sh = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
        add = q(torch.randn(2, c, 11, 13, generator=g), dtype)
        ref = F.gelu(y * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)) + add
        out = ops.dwconv(to_act(x, dtype), wk, Act.empty(2, 11, 13, c, dtype, DEV), bias=b.to(DEV), scale=sc.to(DEV), shift=sh.to(DEV), k=k,
                         act=ops.ACT_GELU, add=to_act(add, dtype))
    assert rel_err(from_act(out), ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k", [3, 5, 7])
def test_dwconv_row_strips(dtype, k):
    """register-tiled depthwise kernel on a shape whose CTAs walk several row strips (grid.y < number of strips) with a ragged last strip"""
    g = torch.Generator().manual_seed(40 + k)
    n, c, h, w_ = 40, 256, 38, 13
    x = q(torch.randn(n, c, h, w_, generator=g), dtype)
    w = torch.randn(c, 1, k, k, generator=g) / k
    b = torch.randn(c, generator=g) * 0.1
    ref = F.conv2d(x, w, b, 1, k // 2, groups=c)
    wk = w.reshape(c, k * k).t().contiguous().to(DEV)
    out = ops.dwconv(to_act(x, dtype), wk, Act.empty(n, h, w_, c, dtype, DEV), bias=b.to(DEV), k=k)
    assert rel_err(from_act(out), ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c,groups,hw", [(128, 16, 20), (64, 16, 10), (64, 16, 41), (32, 8, 6)])
def test_group_norm_silu_add(dtype, c, groups, hw):
    g = torch.Generator().manual_seed(c + hw)
    x = q(torch.randn(3, c, hw, hw, generator=g) * 2 + 0.5, dtype)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    add = q(torch.randn(3, c, hw, hw, generator=g), dtype)
    ref = F.silu(F.group_norm(x, groups, gamma, beta, 1e-5)) + add
    out = ops.group_norm(to_act(x, dtype), Act.empty(3, hw, hw, c, dtype, DEV), torch.empty(3, groups, 2, dtype=torch.float64, device=DEV),
                         groups, gamma.to(DEV), beta.to(DEV), 1e-5, ops.ACT_SILU, to_act(add, dtype))
    assert rel_err(from_act(out), ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,c,h,w", [(2, 32, 20, 20), (3, 48, 45, 23), (1, 16, 7, 5), (2, 16, 40, 40), (1, 24, 9, 9)])
def test_sppf_pool(dtype, n, c, h, w):
    """bit-exact (max is exact in any storage type); multi-tile maps, maps smaller than the 13x13 window, a channel count (24) off the tiled path"""
    x = q(torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(h)), dtype)
    cat = Act.empty(n, h, w, 4 * c, dtype, DEV)
    cat.buf[..., :c] = x.permute(0, 2, 3, 1).to(DEV)
    ops.sppf_pool(cat.slice(0, c), cat.slice(c, c), cat.slice(2 * c, c), cat.slice(3 * c, c))
    y1 = F.max_pool2d(x, 5, 1, 2); y2 = F.max_pool2d(y1, 5, 1, 2); y3 = F.max_pool2d(y2, 5, 1, 2)
    got = from_act(cat)
    for i, r in enumerate((x, y1, y2, y3)):
        np.testing.assert_array_equal(got[:, c * i:c * i + c].numpy(), r.numpy())


@pytest.mark.parametrize("dtype", DTYPES)
def test_gap_rowcol(dtype):
    x = q(torch.randn(3, 64, 10, 14, generator=torch.Generator().manual_seed(2)) + 0.3, dtype)
    a = to_act(x, dtype)
    out = ops.gap(a, torch.empty(3, 64, device=DEV))
    assert rel_err(out.cpu(), x.mean((2, 3))) < 1e-4
    rows, cols = Act.empty(3, 10, 1, 64, dtype, DEV), Act.empty(3, 14, 1, 64, dtype, DEV)
    ops.rowcol_mean(a, rows, cols)
    assert rel_err(from_act(rows)[..., 0], x.mean(3)) < tol(dtype)
    assert rel_err(from_act(cols)[..., 0], x.mean(2)) < tol(dtype)
    gh, gw = q(torch.rand(3, 64, 10, 1), dtype), q(torch.rand(3, 64, 14, 1), dtype)
    y = ops.rowcol_gate(a, to_act(gh, dtype), to_act(gw, dtype), Act.empty(3, 10, 14, 64, dtype, DEV))
    assert rel_err(from_act(y), x * gh * gw.permute(0, 1, 3, 2)) < tol(dtype)
    y = ops.rowcol_gate(None, to_act(gh, dtype), to_act(gw, dtype), Act.empty(3, 10, 14, 64, dtype, DEV))
    assert rel_err(from_act(y), gh * gw.permute(0, 1, 3, 2)) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c,h,w", [(128, 80, 80), (64, 37, 150), (32, 5, 160), (256, 20, 20), (96, 3, 1), (64, 9, 300), (48, 21, 33), (16, 40, 64)])
def test_rowcol_mean_single_pass(dtype, c, h, w):
    """the one-read kernel (32-channel slices, column sums in registers, butterfly row sums): ragged widths up to 160 (1, 2, 3 and 5 column
    positions per lane), a channel-sliced input view (ld > c), and a map wider than 160 that must take the two-pass fallback"""
    x = q(torch.randn(3, c + 32, h, w, generator=torch.Generator().manual_seed(h + w)) + 0.2, dtype)
    a = to_act(x, dtype).slice(32, c)
    xs = x[:, 32:]
    rows, cols = Act.empty(3, h, 1, c, dtype, DEV), Act.empty(3, w, 1, c, dtype, DEV)
    ops.rowcol_mean(a, rows, cols)
    assert rel_err(from_act(rows)[..., 0], xs.mean(3)) < tol(dtype)
    assert rel_err(from_act(cols)[..., 0], xs.mean(2)) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hw,s", [(20, 2), (20, 4), (40, 4), (10, 2), (5, 4), (12, 4)])
def test_pool_upsample(dtype, hw, s):
    x = q(torch.randn(2, 16, hw, hw, generator=torch.Generator().manual_seed(hw + s)), dtype)
    ref = F.interpolate(F.adaptive_avg_pool2d(x, (hw // s, hw // s)), size=(hw, hw), mode="bilinear", align_corners=False)
    y = ops.pool_upsample(to_act(x, dtype), s, Act.empty(2, hw, hw, 16, dtype, DEV))
    assert rel_err(from_act(y), ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("h,w,s", [(20, 20, 2), (20, 20, 4), (13, 22, 2), (9, 7, 4)])
def test_pool_upsample_batched(dtype, h, w, s):
    """batch 24 x 96 channels: the two-phase shared-memory kernel (>= 64 CTAs), incl. maps that s does not divide (ragged adaptive bins)"""
    x = q(torch.randn(24, 96, h, w, generator=torch.Generator().manual_seed(h + s)), dtype)
    ref = F.interpolate(F.adaptive_avg_pool2d(x, (h // s, w // s)), size=(h, w), mode="bilinear", align_corners=False)
    y = ops.pool_upsample(to_act(x, dtype), s, Act.empty(24, h, w, 96, dtype, DEV))
    assert rel_err(from_act(y), ref) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c,hw", [(32, 40), (64, 20), (64, 10), (32, 13)])
def test_mlca(dtype, c, hw):
    g = torch.Generator().manual_seed(c + hw)
    x = q(torch.randn(2, c, hw, hw, generator=g), dtype)
    add = q(torch.randn(2, c, hw, hw, generator=g), dtype)
    sd = {"a.conv.weight": torch.randn(1, 1, 3, generator=g) * 0.5, "a.conv_local.weight": torch.randn(1, 1, 3, generator=g) * 0.5}
    ref = om.mlca(sd, "a", x) + add
    y = ops.mlca(to_act(x, dtype), Act.empty(2, hw, hw, c, dtype, DEV), sd["a.conv.weight"].reshape(-1).to(DEV),
                 sd["a.conv_local.weight"].reshape(-1).to(DEV), 3, torch.empty(2, 25, c, device=DEV), torch.empty(2, 25, c, device=DEV),
                 add=to_act(add, dtype))
    assert rel_err(from_act(y), ref) < tol(dtype)


def test_gate_mlp():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 128, generator=g)
    w1, b1, w2, b2 = torch.randn(32, 128, generator=g) / 11, torch.randn(32, generator=g), torch.randn(3, 32, generator=g) / 5, torch.randn(3, generator=g)
    hid = F.relu(x @ w1.t() + b1)
    out = ops.gate_mlp(x.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), torch.empty(5, 3, device=DEV), kind=1)
    assert rel_err(out.cpu(), (hid @ w2.t() + b2).softmax(1)) < 1e-5
    out = ops.gate_mlp(x.to(DEV), w1.to(DEV), b1.to(DEV), w2[:1].contiguous().to(DEV), b2[:1].contiguous().to(DEV), torch.empty(5, 1, device=DEV), kind=0)
    assert rel_err(out.cpu(), torch.sigmoid(hid @ w2[:1].t() + b2[:1])) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
def test_adt_and_eltwise_and_group_mean(dtype):
    g = torch.Generator().manual_seed(4)
    x = q(torch.randn(2, 32, 6, 7, generator=g), dtype)
    imp = torch.rand(2, 3, generator=g).softmax(1)
    al, wv, bv = torch.tensor([0.3, 0.7, 1.2]), torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
    ref = sum(torch.tanh(al[i] * x) * imp[:, i].view(2, 1, 1, 1) for i in range(3)) * wv.view(1, -1, 1, 1) + bv.view(1, -1, 1, 1)
    y = ops.adt_apply(to_act(x, dtype), imp.to(DEV), al.to(DEV), wv.to(DEV), bv.to(DEV), Act.empty(2, 6, 7, 32, dtype, DEV))
    assert rel_err(from_act(y), ref) < tol(dtype)
    b_, c_, d_ = (q(torch.randn(2, 32, 6, 7, generator=g), dtype) for _ in range(3))
    E = lambda: Act.empty(2, 6, 7, 32, dtype, DEV)
    assert rel_err(from_act(ops.eltwise(0, to_act(x, dtype), to_act(b_, dtype), E(), alpha=0.3, beta=0.6)), 0.3 * x + 0.6 * b_) < tol(dtype)
    assert rel_err(from_act(ops.eltwise(1, to_act(x, dtype), to_act(b_, dtype), E())), x * b_) < tol(dtype)
    assert rel_err(from_act(ops.eltwise(3, to_act(x, dtype), to_act(b_, dtype), E(), c3=to_act(c_, dtype), d4=to_act(d_, dtype), alpha=0.2,
                                        beta=0.3, gamma=0.4)), 0.2 * x + 0.3 * b_ + 0.4 * c_ + d_) < tol(dtype)
    t = q(torch.randn(2, 16, 3 * 10, 1, generator=g), dtype)  # (n, c, 3T, 1)
    y = ops.group_mean(to_act(t, dtype), 3, Act.empty(2, 10, 1, 16, dtype, DEV))
    assert rel_err(from_act(y), t.view(2, 16, 3, 10, 1).mean(2)) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("T", [25, 100, 400])
def test_tssa_and_mha(dtype, T):
    g = torch.Generator().manual_seed(T)
    C_, heads = 128, 2
    qkv = q(torch.randn(2, T, 3 * C_, generator=g), dtype)
    temps = torch.rand(heads, generator=g) + 0.5
    # oracle pieces of cross_scale_attention_tssa (block.py:2459-2474)
    qq, kk, vv = (t.view(2, T, heads, 64).transpose(1, 2) for t in qkv.chunk(3, -1))
    wn = F.normalize(qq, dim=-1)
    pi = torch.softmax((wn ** 2).sum(-1) * temps.view(1, heads, 1), -1)
    ref = (-(vv * pi.unsqueeze(-1)) * (1.0 / (1 + pi.unsqueeze(-2) @ (kk ** 2)))).transpose(1, 2).reshape(2, T, C_)
    a = Act(qkv.view(2, T, 1, 3 * C_).to(DEV).to(dtype).contiguous())
    out = Act.empty(2, 2 * T, 1, C_, dtype, DEV)
    out.buf.zero_()
    ops.tssa(a, temps.to(DEV), heads, out, T)
    got = out.torch().float().cpu().view(2, 2 * T, C_)
    assert rel_err(got[:, T:], ref) < tol(dtype)
    assert float(got[:, :T].abs().max()) == 0.0
    # MHA core vs softmax(QK^T/sqrt(d))V
    att = torch.softmax(qq @ kk.transpose(-1, -2) / 8.0, -1) @ vv
    ref2 = att.transpose(1, 2).reshape(2, T, C_)
    o2 = ops.mha(a, heads, Act.empty(2, T, 1, C_, dtype, DEV))
    assert rel_err(o2.torch().float().cpu().view(2, T, C_), ref2) < tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hw,n,c", [(20, 2, 16), (16, 2, 16), (10, 2, 16),
                                    (20, 9, 24), (13, 17, 40), (8, 70, 8)])  # >= 64 patches in the batch: the matrix-resident kernel
def test_patch_filter(dtype, hw, n, c):
    g = torch.Generator().manual_seed(hw + n)
    x = q(torch.randn(n, c, hw, hw, generator=g), dtype)
    add = q(torch.randn(n, c, hw, hw, generator=g), dtype)
    fft = 1.0 + 0.1 * torch.randn(c, 1, 1, 8, 5, generator=g)
    hn = (8 - hw % 8) % 8
    xp = F.pad(x, (0, hn, 0, hn), mode="reflect")
    H = xp.shape[2]
    p = xp.view(n, c, H // 8, 8, H // 8, 8).permute(0, 1, 2, 4, 3, 5)
    p = torch.fft.irfft2(torch.fft.rfft2(p) * fft, s=(8, 8))
    ref = add + 0.4 * p.permute(0, 1, 2, 4, 3, 5).reshape(n, c, H, H)[:, :, :hw, :hw]
    y = ops.patch_filter(to_act(x, dtype), edffn_spectral_matrix(fft).to(DEV), Act.empty(n, hw, hw, c, dtype, DEV), alpha=0.4, add=to_act(add, dtype))
    assert rel_err(from_act(y), ref) < tol(dtype)


def test_nchw_to_nhwc():
    img = torch.rand(2, 3, 32, 64)
    y = ops.nchw_to_nhwc(img.to(DEV), Act.empty(2, 32, 64, 8, torch.float32, DEV))
    got = from_act(y)
    np.testing.assert_array_equal(got[:, :3].numpy(), img.numpy())
    assert float(got[:, 3:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("u8", [False, True])
def test_stem_conv(dtype, u8):
    g = torch.Generator().manual_seed(11)
    img = torch.randint(0, 256, (2, 3, 64, 96), generator=g, dtype=torch.uint8) if u8 else torch.rand(2, 3, 64, 96, generator=g)
    w = torch.randn(16, 3, 3, 3, generator=g) / 5
    b = torch.randn(16, generator=g) * 0.1
    x = img.float() / 255.0 if u8 else img
    ref = F.silu(F.conv2d(x, w, b, 2, 1))
    wq = (w / 255.0 if u8 else w).permute(0, 2, 3, 1).reshape(16, -1).contiguous()
    y = ops.stem_conv(img.to(DEV), wq.to(DEV), b.to(DEV), Act.empty(2, 32, 48, 16, dtype, DEV))
    assert rel_err(from_act(y), ref) < tol(dtype)


@pytest.mark.parametrize("n,h,w", [(1, 640, 640), (3, 62, 304), (2, 34, 1040), (1, 8, 16), (2, 30, 300)])
def test_stem_conv_tensor_core_u8(n, h, w):
    """stem_tc_kernel (uint8 image, bf16 storage: mma.sync implicit GEMM) on whole / ragged 128-pixel row segments, odd output heights, image borders
    and a map narrower than one MMA tile; f16 weights (11 bits) + bf16 output rounding: <= tol(bf16) against the fp32 convolution"""
    g = torch.Generator().manual_seed(h + w)
    img = torch.randint(0, 256, (n, 3, h, w), generator=g, dtype=torch.uint8)
    wt = torch.randn(16, 3, 3, 3, generator=g) / 5
    b = torch.randn(16, generator=g) * 0.1
    ref = F.silu(F.conv2d(img.float() / 255.0, wt, b, 2, 1))
    wq = (wt / 255.0).permute(0, 2, 3, 1).reshape(16, -1).contiguous()
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = ops.stem_conv(img.to(DEV), wq.to(DEV), b.to(DEV), Act.empty(n, ho, wo, 16, torch.bfloat16, DEV))
    assert rel_err(from_act(y), ref) < tol(torch.bfloat16)
    # the f16 tensor-core path against the fp32-weight SIMT kernel on the same input: they differ by the weight rounding only
    y32 = ops.stem_conv(img.to(DEV), wq.to(DEV), b.to(DEV), Act.empty(n, ho, wo, 16, torch.float32, DEV))
    assert rel_err(from_act(y), from_act(y32)) < tol(torch.bfloat16)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,h,w,c,mip,oup", [(2, 80, 80, 64, 8, 64), (3, 7, 13, 64, 8, 64), (1, 20, 5, 512, 16, 256), (2, 3, 40, 1024, 32, 64), (65, 1, 1, 8, 8, 8)])
def test_coordatt_mlp(dtype, n, h, w, c, mip, oup):
    """yad_coordatt_mlp against nn/modules/head.py:694-703 (conv1 with the eval-mode bn1 folded in -> Hardswish -> conv_h / conv_w -> sigmoid)"""
    g = torch.Generator().manual_seed(n + h + c)
    rows, cols = q(torch.randn(n, c, h, 1, generator=g), dtype), q(torch.randn(n, c, w, 1, generator=g), dtype)
    w1, b1 = torch.randn(mip, c, generator=g) / c ** 0.5, torch.randn(mip, generator=g)
    wh, bh = torch.randn(oup, mip, generator=g) / mip ** 0.5, torch.randn(oup, generator=g) * 0.3
    ww, bw = torch.randn(oup, mip, generator=g) / mip ** 0.5, torch.randn(oup, generator=g) * 0.3
    ref_h = torch.sigmoid(F.conv2d(F.hardswish(F.conv2d(rows, w1.view(mip, c, 1, 1), b1)), wh.view(oup, mip, 1, 1), bh))
    ref_w = torch.sigmoid(F.conv2d(F.hardswish(F.conv2d(cols, w1.view(mip, c, 1, 1), b1)), ww.view(oup, mip, 1, 1), bw))
    d = lambda t: t.contiguous().to(DEV)  # noqa: E731
    gh, gw = ops.coordatt_mlp(to_act(rows, dtype), to_act(cols, dtype), d(w1), d(b1), d(wh), d(bh), d(ww), d(bw),
                              Act.empty(n, h, 1, oup, dtype, DEV), Act.empty(n, w, 1, oup, dtype, DEV))
    t = 1e-5 if dtype == torch.float32 else 5e-3  # bf16: the only rounding is the store of the gate (values in (0, 1))
    assert rel_err(from_act(gh), ref_h) < t and rel_err(from_act(gw), ref_w) < t


@pytest.mark.parametrize("n,T,heads", [(2, 65, 2), (1, 128, 1), (3, 129, 2), (2, 1200, 2), (1, 1237, 4), (2, 4800, 2)])
def test_mha_bf16_long_sequences(n, T, heads):
    """yad_mha on the 128-query tensor-core kernel (two 16-row tiles per warp, ldmatrix.x4 fragments) against softmax(q k^T / sqrt(d)) v
    (nn.MultiheadAttention's core, block.py:2479-2488): tile tails, the T = 1200 of the benchmark and the T = 4800 of 1280^2 inputs"""
    g = torch.Generator().manual_seed(T + heads)
    c = heads * 64
    qkv = q(torch.randn(n, T, 3 * c, generator=g), torch.bfloat16)
    qq, kk, vv = (t.view(n, T, heads, 64).transpose(1, 2) for t in qkv.chunk(3, -1))
    ref = F.scaled_dot_product_attention(qq, kk, vv).transpose(1, 2).reshape(n, T, c)
    a = Act(qkv.view(n, T, 1, 3 * c).to(DEV).to(torch.bfloat16).contiguous())
    out = ops.mha(a, heads, Act.empty(n, T, 1, c, torch.bfloat16, DEV))
    assert rel_err(out.torch().float().cpu().view(n, T, c), ref) < tol(torch.bfloat16)
