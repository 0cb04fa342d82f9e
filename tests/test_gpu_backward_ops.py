"""GPU parity of the backward building blocks against torch autograd of the same operator (fp32 torch-CUDA reference, computed from the SAME
rounded inputs): convolution weight gradients (yad_conv_wgrad: SIMT fp32, SIMT bf16, mma.sync bf16) and input gradients (yad_conv2d on dy
with the permuted weight layouts of train_params.TrainParams.conv) for every geometry the model uses; depthwise; normalisation backward."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util_gpu import DEV, rel_err
from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.ops import Act
from yolo_ad_refine_b200.train_params import TrainParams

pytestmark = pytest.mark.gpu
# the torch-CUDA reference must be true fp32 (cuDNN / cuBLAS default to TF32 for convolutions)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# (cin, cout, kh, kw, stride, n, h, w)
GEOMS = [
    (8, 16, 3, 3, 2, 2, 32, 32),      # layer 0 (3 -> 8 padded input channels)
    (16, 32, 3, 3, 2, 2, 16, 24),
    (64, 64, 3, 3, 1, 2, 20, 12),
    (128, 64, 3, 3, 1, 3, 10, 10),
    (64, 32, 1, 1, 1, 2, 20, 20),
    (192, 128, 1, 1, 1, 2, 9, 7),
    (256, 256, 7, 1, 1, 4, 20, 1),    # ELA_HSFPN Conv1d(k=7) on the (n, L, 1, c) view
    (64, 27, 3, 3, 1, 2, 10, 10),     # offset / mask conv (27 -> 32 padded outputs)
    (32, 1, 3, 3, 1, 2, 10, 10),      # cls_prob (1 -> 8 padded outputs)
    (576, 64, 1, 1, 1, 2, 10, 10),    # deformable conv on its column tensor
    (16, 8, 3, 3, 1, 2, 40, 40),      # small-channel wgrad path (cin, cout <= 32): ragged 8 x 32 tiles
    (32, 32, 3, 3, 1, 3, 40, 24),
    (32, 16, 3, 3, 1, 1, 9, 7),
    (8, 32, 3, 3, 2, 2, 48, 80),
    (128, 128, 3, 3, 1, 2, 40, 40),   # tcgen05 wgrad: 3 M-tile groups x pixel splits
    (128, 256, 1, 1, 1, 3, 20, 20),
    (64, 80, 1, 1, 1, 2, 16, 24),     # cout not a multiple of 64: second dy box partly zero-filled
    (64, 128, 3, 3, 2, 2, 40, 40),    # stride 2 through the TMA traversal stride
    (128, 128, 3, 3, 2, 2, 20, 28),
    (32, 64, 3, 3, 1, 2, 24, 40),     # cin < 64: partial channel group, zero-filled by TMA
    (96, 128, 1, 1, 1, 2, 20, 20),
    (48, 64, 1, 1, 1, 1, 33, 17),     # flat 1x1 with a ragged pixel count
]


def _tp(w, b, dtype):
    sd = {"w": w, "b": b}
    return TrainParams(sd, dtype, DEV)


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("mode", ["fp32", "bf16_simt", "bf16_tc", "bf16_mma"])
def test_conv_wgrad_dgrad(geom, mode):
    """bf16_tc: impl 0 = tcgen05 / TMEM wgrad where eligible (stride 1, cin % 64 == 0), else the mma.sync kernels; bf16_mma: impl 3 forces
    the mma.sync wgrad kernels (and the thread-gathered tcgen05 conv for the input gradient)"""
    cin, cout, kh, kw, stride, n, h, w = geom
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    impl = {"bf16_tc": 0, "bf16_mma": 3}.get(mode, 1)
    rs = np.random.RandomState(cin * 7 + cout)
    wt = torch.from_numpy(rs.standard_normal((cout, cin, kh, kw)).astype(np.float32) * 0.1)
    tp = _tp(wt, torch.zeros(cout), dtype)
    x = torch.from_numpy(rs.standard_normal((n, cin, h, w)).astype(np.float32)).to(DEV).to(dtype)
    ph, pw = kh // 2, kw // 2
    ho, wo = (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1
    dy = torch.from_numpy(rs.standard_normal((n, cout, ho, wo)).astype(np.float32)).to(DEV).to(dtype)
    # reference: torch autograd in fp32 from the rounded operands
    xr = x.float().requires_grad_(True)
    wr = wt.to(DEV).to(dtype).float().requires_grad_(True)
    F.conv2d(xr, wr, None, stride, (ph, pw)).backward(dy.float())
    xa, dya = Act.from_nchw(x, dtype), Act.from_nchw(dy, dtype)
    W = tp.conv("w", "fwd")
    Wd = tp.conv("w", "dgrad" if stride == 1 else "dgrad_t")
    tp.zero_grad()
    ops.conv_wgrad(xa, dya, W.gw, kh, kw, stride, ph, pw, impl=impl)
    tp.unpack_grads()
    tol = 1e-3 if dtype == torch.float32 else 4e-3  # fp32 accumulation everywhere; bf16: products of exactly representable operands
    assert rel_err(tp.g("w").cpu().numpy(), wr.grad.cpu().numpy()) < tol
    dx = Act.empty(n, h, w, xa.c, dtype, DEV)
    if stride == 1:
        ops.conv2d(dya, Wd.w, dx, kh=kh, kw=kw, stride=1, pad_h=kh - 1 - ph, pad_w=kw - 1 - pw, impl=impl)
    else:
        ops.conv2d(dya, Wd.w, dx, kh=kh, kw=kw, stride=2, pad_h=ph, pad_w=pw, mode=ops.CONV_TRANSPOSED, impl=impl)
    got = dx.nchw().float().cpu().numpy()[:, :cin]
    assert rel_err(got, xr.grad.cpu().numpy()) < (1e-3 if dtype == torch.float32 else 1.5e-2)  # bf16: the stored dx is rounded once
    # accumulate form (epilogue add aliasing the output)
    if stride == 1:
        ops.conv2d(dya, Wd.w, dx, kh=kh, kw=kw, stride=1, pad_h=kh - 1 - ph, pad_w=kw - 1 - pw, add=dx, impl=impl)
        got2 = dx.nchw().float().cpu().numpy()[:, :cin]
        assert rel_err(got2, 2 * xr.grad.cpu().numpy()) < (1e-3 if dtype == torch.float32 else 3e-2)


@pytest.mark.parametrize("mode", ["fp32", "bf16_simt", "bf16_tc"])
def test_conv_transpose_wgrad_dgrad(mode):
    """nn.ConvTranspose2d(c, c, 3, 2, 1, 1) of the neck (yaml layers 13 / 20)"""
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    impl = 0 if mode == "bf16_tc" else 1
    rs = np.random.RandomState(3)
    ci, co, n, h, w = 64, 32, 2, 10, 12
    wt = torch.from_numpy(rs.standard_normal((ci, co, 3, 3)).astype(np.float32) * 0.1)
    tp = _tp(wt, torch.zeros(co), dtype)
    x = torch.from_numpy(rs.standard_normal((n, ci, h, w)).astype(np.float32)).to(DEV).to(dtype)
    dy = torch.from_numpy(rs.standard_normal((n, co, 2 * h, 2 * w)).astype(np.float32)).to(DEV).to(dtype)
    xr = x.float().requires_grad_(True)
    wr = wt.to(DEV).to(dtype).float().requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, None, 2, 1, 1)
    yr.backward(dy.float())
    xa, dya = Act.from_nchw(x, dtype), Act.from_nchw(dy, dtype)
    Wf, Wt = tp.conv("w", "convT_fwd"), tp.conv("w", "convT_dgrad")
    y = Act.empty(n, 2 * h, 2 * w, co, dtype, DEV)
    ops.conv2d(xa, Wf.w, y, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, mode=ops.CONV_TRANSPOSED, impl=impl)
    assert rel_err(y.nchw().float().cpu().numpy(), yr.detach().cpu().numpy()) < (1e-3 if dtype == torch.float32 else 1.5e-2)
    tp.zero_grad()
    ops.conv_wgrad(dya, xa, Wt.gw, 3, 3, 2, 1, 1, impl=impl)
    tp.unpack_grads()
    assert rel_err(tp.g("w").cpu().numpy(), wr.grad.cpu().numpy()) < (1e-3 if dtype == torch.float32 else 4e-3)
    dx = Act.empty(n, h, w, ci, dtype, DEV)
    ops.conv2d(dya, Wt.w, dx, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, impl=impl)
    assert rel_err(dx.nchw().float().cpu().numpy(), xr.grad.cpu().numpy()) < (1e-3 if dtype == torch.float32 else 1.5e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("k", [3, 7])
def test_dwconv_backward(dtype, k):
    rs = np.random.RandomState(k)
    c, n, h, w = 128, 2, 20, 20
    wt = torch.from_numpy(rs.standard_normal((c, 1, k, k)).astype(np.float32) * 0.2)
    tp = _tp(wt, torch.zeros(c), dtype)
    x = torch.from_numpy(rs.standard_normal((n, c, h, w)).astype(np.float32)).to(DEV).to(dtype)
    dy = torch.from_numpy(rs.standard_normal((n, c, h, w)).astype(np.float32)).to(DEV).to(dtype)
    xr, wr = x.float().requires_grad_(True), wt.to(DEV).requires_grad_(True)
    F.conv2d(xr, wr, None, 1, k // 2, groups=c).backward(dy.float())
    xa, dya = Act.from_nchw(x, dtype), Act.from_nchw(dy, dtype)
    _, gw = tp.f32("w", "dw")
    wflip, _ = tp.f32("w", "dw_flip")
    tp.zero_grad()
    ops.dwconv_wgrad(xa, dya, k, gw)
    tp.unpack_grads()
    assert rel_err(tp.g("w").cpu().numpy(), wr.grad.cpu().numpy()) < 1e-3
    dx = ops.dwconv(dya, wflip, Act.empty(n, h, w, c, dtype, DEV), k=k)
    assert rel_err(dx.nchw().float().cpu().numpy(), xr.grad.cpu().numpy()) < (1e-3 if dtype == torch.float32 else 1.5e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kind", ["bn_silu", "bn_gelu", "bn_hswish", "gn_silu", "gn_sigmoid", "gn_none"])
def test_norm_backward(dtype, kind):
    """yad_norm_bwd against autograd of F.batch_norm(training=True) / F.group_norm followed by the activation"""
    rs = np.random.RandomState(11)
    c, n, h, w = 64, 3, 12, 10
    x = torch.from_numpy(rs.standard_normal((n, c, h, w)).astype(np.float32) * 2 + 0.5).to(DEV).to(dtype)
    dy = torch.from_numpy(rs.standard_normal((n, c, h, w)).astype(np.float32)).to(DEV).to(dtype)
    gamma = torch.from_numpy(rs.uniform(0.5, 1.5, c).astype(np.float32)).to(DEV)
    beta = torch.from_numpy(rs.standard_normal(c).astype(np.float32) * 0.3).to(DEV)
    norm, actn = kind.split("_")
    act = dict(silu=ops.ACT_SILU, gelu=ops.ACT_GELU, hswish=ops.ACT_HARDSWISH, sigmoid=ops.ACT_SIGMOID, none=ops.ACT_NONE)[actn]
    fn = dict(silu=F.silu, gelu=F.gelu, hswish=F.hardswish, sigmoid=torch.sigmoid, none=lambda t: t)[actn]
    xr, gr, br = x.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    eps = 1e-3 if norm == "bn" else 1e-5
    u = F.batch_norm(xr, None, None, gr, br, True, 0.0, eps) if norm == "bn" else F.group_norm(xr, 16, gr, br, eps)
    yr = fn(u)
    yr.backward(dy.float())
    xa, dya = Act.from_nchw(x, dtype), Act.from_nchw(dy, dtype)
    xv, dyv = (xa.reshape(1, n * h, w), dya.reshape(1, n * h, w)) if norm == "bn" else (xa, dya)
    groups = c if norm == "bn" else 16
    stats = torch.empty((xv.n, groups, 2), dtype=torch.float64, device=DEV)
    y = Act.empty(xv.n, xv.h, xv.w, c, dtype, DEV)
    ops.group_norm(xv, y, stats, groups, gamma, beta, eps, act)
    tol = 1e-3 if dtype == torch.float32 else 1.5e-2
    assert rel_err(y.reshape(n, h, w).nchw().float().cpu().numpy(), yr.detach().cpu().numpy()) < tol
    dgam, dbet = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    dx = Act.empty(xv.n, xv.h, xv.w, c, dtype, DEV)
    ops.norm_bwd(xv, dyv, stats, groups, gamma, beta, eps, act, torch.empty_like(stats), dgam, dbet, dx, 0)
    assert rel_err(dx.reshape(n, h, w).nchw().float().cpu().numpy(), xr.grad.cpu().numpy()) < tol
    assert rel_err(dgam.cpu().numpy(), gr.grad.cpu().numpy()) < 2e-3
    assert rel_err(dbet.cpu().numpy(), br.grad.cpu().numpy()) < 2e-3
    if norm == "bn":  # running statistics (momentum 0.03, unbiased variance)
        rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        ops.bn_running_update(stats, c, n * h * w, 0.03, rm, rv)
        rm_r, rv_r = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        F.batch_norm(x.float(), rm_r, rv_r, None, None, True, 0.03, eps)
        np.testing.assert_allclose(rm.cpu().numpy(), rm_r.cpu().numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(rv.cpu().numpy(), rv_r.cpu().numpy(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T", [150, 1200])
def test_mha_backward(dtype, T):
    """yad_mha_bwd (fp32: SIMT kernels; bf16: flash-style mma.sync kernels) against autograd of softmax(q k^T / sqrt(d)) v, 2 heads x 64"""
    rs = np.random.RandomState(T)
    n, heads, c = 2, 2, 128
    qkv = torch.from_numpy(rs.standard_normal((n, T, 3 * c)).astype(np.float32)).to(DEV).to(dtype)
    dout = torch.from_numpy(rs.standard_normal((n, T, c)).astype(np.float32)).to(DEV).to(dtype)
    ref_in = qkv.float().requires_grad_(True)
    q, k, v = [t.view(n, T, heads, 64).transpose(1, 2) for t in ref_in.split(c, -1)]
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1)
    o_ref = (att @ v).transpose(1, 2).reshape(n, T, c)
    o_ref.backward(dout.float())
    qa = Act(qkv.view(n, 1, T, 3 * c).contiguous())
    out = ops.mha(qa, heads, Act.empty(n, 1, T, c, dtype, DEV))
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert rel_err(out.torch().float().cpu().numpy().reshape(n, T, c), o_ref.detach().cpu().numpy()) < tol
    dq = Act.empty(n, 1, T, 3 * c, dtype, DEV)
    ops.mha_bwd(qa, heads, out, Act(dout.view(n, 1, T, c).contiguous()), dq, torch.empty((n * heads, T, 2), dtype=torch.float32, device=DEV))
    got = dq.torch().float().cpu().numpy().reshape(n, T, 3 * c)
    ref = ref_in.grad.cpu().numpy()
    for i, name in enumerate("qkv"):
        e = rel_err(got[..., i * c:(i + 1) * c], ref[..., i * c:(i + 1) * c])
        assert e < (2e-3 if dtype == torch.float32 else 3e-2), (name, e)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("hw", [(12, 20), (8, 8)])
def test_deform_col_forward_backward(dtype, hw):
    """yad_deform_col / yad_deform_col_bwd (DCNv2 sampling as a column tensor) against autograd of the oracle's modulated deformable conv
    (oracle.model.deform_conv3x3, pinned to torchvision.ops.deform_conv2d): output, d(input), d(offsets), d(mask logits)"""
    from oracle.model import deform_conv3x3
    h, w = hw
    rs = np.random.RandomState(h * w)
    n, c, co = 2, 64, 16
    x = torch.from_numpy(rs.standard_normal((n, c, h, w)).astype(np.float32)).to(dtype).float()
    om = torch.from_numpy(rs.standard_normal((n, 27, h, w)).astype(np.float32) * 1.5).to(dtype).float()
    wt = torch.from_numpy(rs.standard_normal((co, c, 3, 3)).astype(np.float32) * 0.1)
    dy = torch.from_numpy(rs.standard_normal((n, co, h, w)).astype(np.float32))
    xr, omr = x.clone().requires_grad_(True), om.clone().requires_grad_(True)
    y_ref = deform_conv3x3(xr, omr[:, :18], omr[:, 18:].sigmoid(), wt)
    y_ref.backward(dy)
    xa, oma = Act.from_nchw(x.to(DEV), dtype), Act.from_nchw(om.to(DEV), dtype)  # 27 -> 32 channels
    col = ops.deform_col(xa, oma, Act.empty(n, h, w, 9 * c, dtype, DEV))
    colt = col.torch().float().cpu().view(n, h, w, 9, c)
    y = torch.einsum("nhwtc,otc->nohw", colt, wt.permute(0, 2, 3, 1).reshape(co, 9, c))
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert rel_err(y.numpy(), y_ref.detach().numpy()) < tol
    dcol = torch.einsum("nohw,otc->nhwtc", dy, wt.permute(0, 2, 3, 1).reshape(co, 9, c)).reshape(n, h, w, 9 * c)
    dcol_a = Act(dcol.to(DEV).to(dtype).contiguous())
    dxf = torch.empty((n, h, w, c), dtype=torch.float32, device=DEV)
    dom = Act.empty(n, h, w, 32, dtype, DEV)
    ops.deform_col_bwd(xa, oma, dcol_a, dxf, dom)
    tolb = 2e-3 if dtype == torch.float32 else 4e-2
    assert rel_err(dxf.permute(0, 3, 1, 2).cpu().numpy(), xr.grad.numpy()) < tolb
    got = dom.nchw().float().cpu().numpy()
    assert rel_err(got[:, :27], omr.grad.numpy()) < tolb
    assert float(np.abs(got[:, 27:]).max()) == 0.0


@pytest.mark.parametrize("geom", [(64, 64, 3, 1, 4, 20, 20), (128, 256, 1, 1, 3, 20, 12), (64, 128, 3, 2, 2, 40, 40), (48, 96, 1, 1, 2, 16, 24)])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_conv_fused_batch_statistics(geom, mode):
    """gn_groups = 0: per-channel (sum, sum of squares) over the whole batch accumulated in the conv epilogue (train-mode BatchNorm) equal the
    stand-alone statistics kernel run on the stored conv output"""
    cin, cout, k, stride, n, h, w = geom
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    rs = np.random.RandomState(cin + cout)
    from yolo_ad_refine_b200.weights import pack_conv
    wt = torch.from_numpy(rs.standard_normal((cout, cin, k, k)).astype(np.float32) / (cin * k * k) ** 0.5)
    cw = pack_conv(wt, None, dtype, DEV, stride)
    x = Act.from_nchw(torch.from_numpy(rs.standard_normal((n, cin, h, w)).astype(np.float32) + 0.3).to(DEV), dtype)
    ho, wo = (h + 2 * (k // 2) - k) // stride + 1, (w + 2 * (k // 2) - k) // stride + 1
    y = Act.empty(n, ho, wo, cout, dtype, DEV)
    fused = torch.empty((1, cout, 2), dtype=torch.float64, device=DEV)
    ops.conv2d(x, cw.w, y, kh=k, kw=k, stride=stride, pad_h=k // 2, pad_w=k // 2, impl=1 if mode == "fp32" else 0, gn_stats=fused, gn_groups=0)
    ref = torch.empty_like(fused)
    ops._call("yad_gn_stats", y.reshape(1, n * ho, wo).yt(), cout, ops._p(ref), ops.dt(dtype), ops.stream_ptr())
    torch.cuda.synchronize()
    np.testing.assert_allclose(fused.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=1e-3)
