"""The YOLO-AD-Refine forward pass expressed as calls into libyad.so (via ops.py), block by block.

Every function names the reference module whose `forward` it replaces (paths relative to /root/reference/ultralytics).  Activations
are NHWC `Act` views; concatenations are eliminated by writing producers into channel windows of one buffer.  This file contains no
arithmetic of its own: all math happens in the CUDA kernels (weights are prepared once in weights.py).
"""
import torch

from . import ops
from .ops import ACT_HARDSWISH, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SILU, Act
from .weights import GN_EPS, ConvW, edffn_spectral_matrix, fold_bn, gn_groups


class Ctx:
    """Execution context: prepared weights + allocation helpers (torch caching allocator = plumbing)."""

    def __init__(self, P, conv_impl=0, parallel_levels=True, dcn_col=False):
        self.P, self.dtype, self.device, self.conv_impl = P, P.dtype, P.device, conv_impl
        self._parallel = parallel_levels
        self.dcn_col = dcn_col  # deformable conv as column tensor + 1x1 GEMM (False: the gathered-operand deformable mode of yad_conv2d)
        self._side = []

    @property
    def parallel_levels(self):
        """Independent branches (pyramid levels, lateral gates, TSSA scales, ...) are forked onto side streams ONLY while a CUDA graph is being
        captured: there they become parallel graph branches over the capture's private memory pool.  In eager mode everything stays on one
        stream -- a tensor allocated on a side stream and consumed on the main stream could be recycled by the caching allocator while the
        main stream still reads it (it tracks only the allocating stream)."""
        return self._parallel and torch.cuda.is_current_stream_capturing()

    def side_streams(self, n):
        while len(self._side) < n:
            self._side.append(torch.cuda.Stream(device=self.device))
        return self._side[:n]

    def act(self, n, h, w, c, ld=None):
        return Act.empty(n, h, w, c, self.dtype, self.device, ld)

    def f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def f64(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.device)


# --------------------------------------------------------------------------------------------------------------------
# generic pieces
# --------------------------------------------------------------------------------------------------------------------
def conv(ctx, x, cw: ConvW, out=None, act=ACT_NONE, mode=ops.CONV_NORMAL, **epi):
    assert x.c == cw.cin, f"conv: input has {x.c} channels, weight expects {cw.cin}"
    if mode == ops.CONV_TRANSPOSED:
        ho, wo = 2 * x.h, 2 * x.w
    elif mode == ops.CONV_DEFORM:
        ho, wo = x.h, x.w
    else:
        ho = (x.h + 2 * cw.pad_h - cw.kh) // cw.stride + 1
        wo = (x.w + 2 * cw.pad_w - cw.kw) // cw.stride + 1
    if out is None:
        out = ctx.act(x.n, ho, wo, cw.cout)
    assert (out.n, out.h, out.w, out.c) == (x.n, ho, wo, cw.cout), ((out.n, out.h, out.w, out.c), (x.n, ho, wo, cw.cout))
    return ops.conv2d(x, cw.w, out, bias=cw.b, kh=cw.kh, kw=cw.kw, stride=cw.stride, pad_h=cw.pad_h, pad_w=cw.pad_w, act=act, mode=mode,
                      impl=ctx.conv_impl, **epi)


def conv_bn_act(ctx, p, x, stride=1, out=None, add=None):
    """nn/modules/conv.py:36-54 Conv.forward_fuse: conv (BN folded) + SiLU"""
    return conv(ctx, x, ctx.P.conv_bn(p, stride), out=out, act=ACT_SILU, add=add)


def conv_gn_act(ctx, p, x, out=None, add=None, act=ACT_SILU, img_scale=None):
    """nn/modules/head.py:1265-1279 Conv_GN: conv(bias=False) -> GroupNorm -> SiLU"""
    cw = ctx.P.conv(p + ".conv.weight")
    g = gn_groups(ctx.P.sd[p + ".conv.weight"].shape[0])
    stats = ctx.f64(x.n, g, 2)
    t = conv(ctx, x, cw, img_scale=img_scale, gn_stats=stats, gn_groups=g)  # statistics accumulated in the conv epilogue
    if out is None:
        out = ctx.act(t.n, t.h, t.w, t.c)
    return ops.group_norm(t, out, stats, g, ctx.P.f32(p + ".gn.weight"), ctx.P.f32(p + ".gn.bias"), GN_EPS, act, add, stats_ready=True)


def mlca(ctx, p, x, out, add=None):
    """nn/modules/block.py:1540-1584 MLCA (+ the residual add of Bottleneck_MLCA)"""
    P = ctx.P
    wg, wl = P.f32(p + ".conv.weight"), P.f32(p + ".conv_local.weight")
    k = wg.numel()
    return ops.mlca(x, out, wg, wl, k, ctx.f32(x.n, 25, x.c), ctx.f32(x.n, 25, x.c), 5, 0.5, add)


def bottleneck(ctx, p, x, out=None, attention=False):
    """nn/modules/block.py:341-354 Bottleneck / :1586-1594 Bottleneck_MLCA (shortcut=True, c1 == c2)"""
    t = conv_bn_act(ctx, p + ".cv1", x)
    if out is None:
        out = ctx.act(x.n, x.h, x.w, x.c)
    if attention:
        t2 = conv_bn_act(ctx, p + ".cv2", t)
        return mlca(ctx, p + ".attention", t2, out, add=x)
    return conv_bn_act(ctx, p + ".cv2", t, out=out, add=x)


def c3k(ctx, p, x, out, attention=False):
    """nn/modules/block.py:256-270 C3.forward / :742-750 C3k / :1596-1600 C3k_MLCA (n = 2 bottlenecks)"""
    c_ = ctx.P.conv_bn(p + ".cv1").cout
    cat = ctx.act(x.n, x.h, x.w, 2 * c_)
    done = None
    if ctx.parallel_levels:  # cv2(x) is independent of the bottleneck chain: a parallel graph branch
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        br = ctx.side_streams(4)[3]
        br.wait_event(fork)
        with torch.cuda.stream(br):
            conv_bn_act(ctx, p + ".cv2", x, out=cat.slice(c_, c_))
            done = torch.cuda.Event()
            done.record(br)
    else:
        conv_bn_act(ctx, p + ".cv2", x, out=cat.slice(c_, c_))
    a = conv_bn_act(ctx, p + ".cv1", x)
    a = bottleneck(ctx, p + ".m.0", a, attention=attention)
    bottleneck(ctx, p + ".m.1", a, out=cat.slice(0, c_), attention=attention)
    if done is not None:
        torch.cuda.current_stream().wait_event(done)
    return conv_bn_act(ctx, p + ".cv3", cat, out=out)


def c3k2(ctx, p, x, use_c3k=False, attention=False):
    """nn/modules/block.py:232-247 C2f.forward / :731-739 C3k2 / :1602-1605 C3k2_MLCA (n = 1)"""
    c = ctx.P.conv_bn(p + ".cv1").cout // 2
    cat = ctx.act(x.n, x.h, x.w, 3 * c)
    conv_bn_act(ctx, p + ".cv1", x, out=cat.slice(0, 2 * c))
    y1 = cat.slice(c, c)
    if use_c3k:
        c3k(ctx, p + ".m.0", y1, cat.slice(2 * c, c), attention)
    else:
        bottleneck(ctx, p + ".m.0", y1, out=cat.slice(2 * c, c), attention=attention)
    return conv_bn_act(ctx, p + ".cv2", cat)


def sppf(ctx, p, x):
    """nn/modules/block.py:177-196 SPPF"""
    c = ctx.P.conv_bn(p + ".cv1").cout
    cat = ctx.act(x.n, x.h, x.w, 4 * c)
    y0 = conv_bn_act(ctx, p + ".cv1", x, out=cat.slice(0, c))
    ops.sppf_pool(y0, cat.slice(c, c), cat.slice(2 * c, c), cat.slice(3 * c, c))
    return conv_bn_act(ctx, p + ".cv2", cat)


def ela_hsfpn(ctx, p, x, flag=True, out=None, factors=False):
    """nn/modules/block.py:1408-1424 ELA_HSFPN: both 1-D branches share Conv1d(k7) + GroupNorm(16) + sigmoid.  factors=True (with flag=False):
    return the two 1-D gates (gh (n, h, 1, c), gw (n, w, 1, c)) instead of their broadcast product -- the consumer is the yaml's Multiply, which
    the lateral 1x1 convolution applies in its epilogue (yad_epilogue.gate_h / gate_w), so the (n, h, w, c) gate map is never written or read."""
    P = ctx.P
    n, h, w, c = x.n, x.h, x.w, x.c
    cw = P.conv(p + ".conv1x1.0.weight", p + ".conv1x1.0.bias")
    gamma, beta = P.f32(p + ".conv1x1.1.weight"), P.f32(p + ".conv1x1.1.bias")
    if out is None and not factors:
        out = ctx.act(n, h, w, c)

    def branch(means):
        stats = ctx.f64(means.n, 16, 2)
        t = conv(ctx, means, cw, gn_stats=stats, gn_groups=16)
        g = ctx.act(t.n, t.h, t.w, t.c)
        return ops.group_norm(t, g, stats, 16, gamma, beta, GN_EPS, ACT_SIGMOID, stats_ready=True)

    if h == w:  # one batched pass over (2n, L, 1, c)
        means = ctx.act(2 * n, h, 1, c)
        ops.rowcol_mean(x, means.images(0, n), means.images(n, n))
        g = branch(means)
        gh, gw = g.images(0, n), g.images(n, n)
    else:
        rows, cols = ctx.act(n, h, 1, c), ctx.act(n, w, 1, c)
        ops.rowcol_mean(x, rows, cols)
        gh, gw = branch(rows), branch(cols)
    if factors:
        assert not flag and out is None
        return gh, gw
    return ops.rowcol_gate(x if flag else None, gh, gw, out)


def mona(ctx, p, x, out=None):
    """nn/modules/mona.py:36-64 Mona.forward (eval: dropout is the identity), five launches:
      x1 = LayerNorm2d(x) * gamma + x * gammax                        yad_ln_mix
      p1 = project1(x1)                                               1x1 conv C -> 64
      s  = (conv1 + conv2 + conv3)(p1) / 3 + p1                       ONE 7x7 depthwise conv: the 3x3 / 5x5 / 7x7 kernels of MonaOp (:12-34) share
                                                                      their centre, so their mean is a single zero-padded 7x7 kernel (+1 at the
                                                                      centre tap for the identity), folded once on the host in fp32
      g  = gelu(s + projector(s))                                     1x1 conv 64 -> 64 with weights W + I, GELU in the epilogue
      y  = x + project2(g)                                            1x1 conv 64 -> C, residual in the epilogue"""
    P = ctx.P
    n, h, w, c = x.n, x.h, x.w, x.c

    def merged_dw():
        k7 = P.sd[p + ".adapter_conv.conv3.weight"].float().clone()            # (64, 1, 7, 7)
        k7[:, :, 1:6, 1:6] += P.sd[p + ".adapter_conv.conv2.weight"].float()
        k7[:, :, 2:5, 2:5] += P.sd[p + ".adapter_conv.conv1.weight"].float()
        k7 /= 3.0
        k7[:, :, 3, 3] += 1.0
        b = (P.sd[p + ".adapter_conv.conv1.bias"].float() + P.sd[p + ".adapter_conv.conv2.bias"].float()
             + P.sd[p + ".adapter_conv.conv3.bias"].float()) / 3.0
        return k7.reshape(k7.shape[0], 49).t().contiguous().to(P.device), b.contiguous().to(P.device)

    def projector_plus_identity():
        wp = P.sd[p + ".adapter_conv.projector.weight"].float().clone()        # (64, 64, 1, 1)
        wp[:, :, 0, 0] += torch.eye(wp.shape[0])
        return wp

    x1 = ops.ln_mix(x, P.f32(p + ".norm.weight"), P.f32(p + ".norm.bias"), P.misc(p + ".gamma", lambda: P.sd[p + ".gamma"].float().reshape(-1).contiguous().to(P.device)),
                    P.misc(p + ".gammax", lambda: P.sd[p + ".gammax"].float().reshape(-1).contiguous().to(P.device)), 1e-5, ctx.act(n, h, w, c))
    p1 = conv(ctx, x1, P.conv(p + ".project1.weight", p + ".project1.bias"))
    dw, db = P.misc(p + ".adapter_dw", merged_dw)
    s = ops.dwconv(p1, dw, ctx.act(n, h, w, p1.c), bias=db, k=7)
    cwp = P.conv_raw(p + ".adapter_conv.projector+I", P.misc(p + ".adapter_pw", projector_plus_identity), P.sd[p + ".adapter_conv.projector.bias"].float())
    g = conv(ctx, s, cwp, act=ops.ACT_GELU)
    return conv(ctx, g, P.conv(p + ".project2.weight", p + ".project2.bias"), out=out, add=x)


def dynamic_tanh(ctx, p, x):
    """nn/modules/block.py:1624-1641 DynamicTanh(channels_last=False): tanh(alpha * x) * weight + bias = the one-branch case of yad_adt_apply
    (importance 1, 0, 0 and alphas alpha, 0, 0)"""
    P = ctx.P
    imp = P.misc(f"dyt.imp.{x.n}", lambda: torch.tensor([[1.0, 0.0, 0.0]] * x.n, dtype=torch.float32).to(P.device))
    alphas = P.misc(p + ".al3", lambda: torch.tensor([P.scalar(p + ".alpha"), 0.0, 0.0], dtype=torch.float32).to(P.device))
    return ops.adt_apply(x, imp, alphas, P.f32(p + ".weight"), P.f32(p + ".bias"), ctx.act(x.n, x.h, x.w, x.c))


def attention_tssa(ctx, p, x, heads, add=None, out=None):
    """nn/modules/block.py:1646-1683 AttentionTSSA on the (n, h*w, c) tokens of an NHWC map: qkv Linear (no bias) -> yad_attention_tssa -> to_out
    Linear (+ bias), with the block's residual in the last epilogue"""
    P = ctx.P
    wq = conv(ctx, x, P.conv(p + ".qkv.weight"))
    temp = P.misc(p + ".temp1d", lambda: P.f32(p + ".temp").reshape(-1).contiguous())
    att = ops.attention_tssa(wq, temp, heads, ctx.act(x.n, x.h, x.w, x.c))
    return conv(ctx, att, P.conv(p + ".to_out.0.weight", p + ".to_out.0.bias"), out=out, add=add)


def tssa_dyt_mona_edffn(ctx, p, x, heads, out=None):
    """nn/modules/block.py:1685-1703 TSSAlock_DYT_Mona_EDFFN.forward (shortcut=True):
    x += attn(dyt1(x)); x = mona1(x); x += ffn(dyt2(x)); x = mona2(x)"""
    t = attention_tssa(ctx, p + ".attn", dynamic_tanh(ctx, p + ".dyt1", x), heads, add=x)
    t = mona(ctx, p + ".mona1", t)
    o, m = edffn(ctx, p + ".ffn", dynamic_tanh(ctx, p + ".dyt2", t), t, None)
    t = ops.patch_filter(o, m, ctx.act(t.n, t.h, t.w, t.c), alpha=1.0, add=t)
    return mona(ctx, p + ".mona2", t, out=out)


def c2tssa_dyt_mona_edffn(ctx, p, x, n=1):
    """nn/modules/block.py:1705-1709 C2TSSA_DYT_Mona_EDFFN + C2PSA.forward :1045-1049 (split, n blocks on b, concat-free cv2)"""
    c = ctx.P.conv_bn(p + ".cv1").cout // 2
    ab = conv_bn_act(ctx, p + ".cv1", x)
    b = ab.slice(c, c)
    cur = b
    for i in range(n):
        cur = tssa_dyt_mona_edffn(ctx, f"{p}.m.{i}", cur, heads=c // 64, out=b if i == n - 1 else None)
    return conv_bn_act(ctx, p + ".cv2", ab)


def psa_attention(ctx, p, x, heads, add=None, out=None):
    """nn/modules/block.py:874-925 Attention(dim, num_heads, attn_ratio=0.5).forward on an NHWC map, head_dim 64 (C2PSA builds num_heads = c // 64):
    qkv 1x1 (+ folded BN) -> softmax(q^T k * key_dim^-0.5) v per head on yad_mha -> + pe(v) (depthwise 3x3 + BN) -> proj 1x1 (+ the block's residual
    in its epilogue).  The reference interleaves [q | k | v] per head with key_dim = head_dim / 2; the qkv rows are re-ordered at weight-packing
    time into the [Q | K | V] thirds yad_mha reads, q / k zero-padded to 64 channels per head (zero channels add nothing to q . k) and the q rows
    scaled by sqrt(64 / key_dim) so that the kernel's 1 / sqrt(64) becomes key_dim^-0.5."""
    P = ctx.P
    n, h, w, c = x.n, x.h, x.w, x.c
    hd = c // heads
    assert hd == 64, "psa_attention: head_dim 64 only (yad_mha)"

    def fold_qkv():
        wq, bq = P.sd[p + ".qkv.conv.weight"].float(), None
        if p + ".qkv.bn.weight" in P.sd:
            wq, bq = fold_bn(wq, P.sd[p + ".qkv.bn.weight"].float(), P.sd[p + ".qkv.bn.bias"].float(), P.sd[p + ".qkv.bn.running_mean"].float(),
                             P.sd[p + ".qkv.bn.running_var"].float(), P.sd.get(p + ".qkv.conv.bias"))
        else:
            bq = P.sd[p + ".qkv.conv.bias"].float()
        kd = (wq.shape[0] // heads - hd) // 2
        per = 2 * kd + hd
        wn, bn = torch.zeros(3 * c, c, 1, 1), torch.zeros(3 * c)
        qs = (hd / kd) ** 0.5
        for i in range(heads):
            src = i * per
            wn[i * hd:i * hd + kd], bn[i * hd:i * hd + kd] = wq[src:src + kd] * qs, bq[src:src + kd] * qs
            wn[c + i * hd:c + i * hd + kd], bn[c + i * hd:c + i * hd + kd] = wq[src + kd:src + 2 * kd], bq[src + kd:src + 2 * kd]
            wn[2 * c + i * hd:2 * c + (i + 1) * hd], bn[2 * c + i * hd:2 * c + (i + 1) * hd] = wq[src + 2 * kd:src + per], bq[src + 2 * kd:src + per]
        return wn, bn

    wn, bn = P.misc(p + ".qkv#mha", fold_qkv)
    qkv = conv(ctx, x, P.conv_raw(p + ".qkv#mha", wn, bn))
    att = ops.mha(qkv, heads, ctx.act(n, h, w, c))

    def fold_pe():
        wp, bp = P.sd[p + ".pe.conv.weight"].float(), None
        if p + ".pe.bn.weight" in P.sd:
            wp, bp = fold_bn(wp, P.sd[p + ".pe.bn.weight"].float(), P.sd[p + ".pe.bn.bias"].float(), P.sd[p + ".pe.bn.running_mean"].float(),
                             P.sd[p + ".pe.bn.running_var"].float(), P.sd.get(p + ".pe.conv.bias"))
        else:
            bp = P.sd[p + ".pe.conv.bias"].float()
        return wp.reshape(c, 9).t().contiguous().to(P.device), bp.contiguous().to(P.device)

    wpe, bpe = P.misc(p + ".pe#dw", fold_pe)
    t = ops.dwconv(qkv.slice(2 * c, c), wpe, ctx.act(n, h, w, c), bias=bpe, k=3, add=att)  # pe(v) + attention output
    return conv(ctx, t, P.conv_bn(p + ".proj"), out=out, add=add)


def psablock(ctx, p, x, heads, out=None):
    """nn/modules/block.py:928-964 PSABlock.forward (shortcut=True): x = x + attn(x); x = x + ffn(x)"""
    t = psa_attention(ctx, p + ".attn", x, heads, add=x)
    f = conv(ctx, t, ctx.P.conv_bn(p + ".ffn.0"), act=ACT_SILU)
    return conv(ctx, f, ctx.P.conv_bn(p + ".ffn.1"), out=out, add=t)


def c2psa(ctx, p, x, n=1):
    """nn/modules/block.py:1010-1049 C2PSA.forward: cv1 -> split (a, b) -> n PSABlocks on b -> cv2 over [a | b] (no concat buffer: b is written back
    into its half of cv1's output)"""
    c = ctx.P.conv_bn(p + ".cv1").cout // 2
    ab = conv_bn_act(ctx, p + ".cv1", x)
    b = ab.slice(c, c)
    cur = b
    for i in range(n):
        cur = psablock(ctx, f"{p}.m.{i}", cur, heads=c // 64, out=b if i == n - 1 else None)
    return conv_bn_act(ctx, p + ".cv2", ab)


def simple_feature_processor(ctx, p, x, alpha=1.0):
    """nn/modules/block.py:2080-2096 SimpleFeatureProcessor: GroupNorm(c // 32 groups) -> depthwise 3x3 (+ bias) -> GELU -> 1x1 (+ bias); alpha scales
    the output in the last epilogue"""
    P = ctx.P
    n, h, w, c = x.n, x.h, x.w, x.c
    g = max(1, c // 32)
    t = ops.group_norm(x, ctx.act(n, h, w, c), ctx.f64(n, g, 2), g, P.f32(p + ".norm.weight"), P.f32(p + ".norm.bias"), GN_EPS, ACT_NONE)
    dw, db, _ = P.dw(p + ".conv_dw.weight", p + ".conv_dw.bias")
    t = ops.dwconv(t, dw, ctx.act(n, h, w, c), bias=db, k=3, act=ops.ACT_GELU)
    return conv(ctx, t, P.conv(p + ".conv_pw.weight", p + ".conv_pw.bias"), alpha=alpha)


def progressive_tssa_fusion0(ctx, p, x, out=None):
    """nn/modules/block.py:2147-2202 ProgressiveTSSA_Fusion0.forward (shortcut=True):
    x = x + SE(pre_attn_block(x)) * residual_weight1;  x = x + ffn(pre_ffn_block(x)) * residual_weight2.
    SEBlock (:2049-2064) = t * sigmoid(W2 relu(W1 mean(t))): residual_weight1 rides on t through the 1x1 epilogue (and 1 / residual_weight1 on W1, so
    that the gate still sees mean(t)); gate on yad_gap + yad_gate_mlp, gate * t + x in one yad_mlca_apply (local_size 1 = one gate per image)."""
    P = ctx.P
    n, h, w, c = x.n, x.h, x.w, x.c
    rw1, rw2 = P.scalar(p + ".residual_weight1"), P.scalar(p + ".residual_weight2")
    t = simple_feature_processor(ctx, p + ".pre_attn_block", x, alpha=rw1)
    avg = ops.gap(t, ctx.f32(n, c))
    w1 = P.misc(p + ".se1", lambda: (P.sd[p + ".attn.fc.0.weight"].float().reshape(-1, c) / (rw1 if rw1 != 0.0 else 1.0)).contiguous().to(P.device))
    w2 = P.misc(p + ".se2", lambda: P.sd[p + ".attn.fc.2.weight"].float().reshape(c, -1).contiguous().to(P.device))
    zb = P.misc(f"zeros.{c}", lambda: torch.zeros(c, dtype=torch.float32, device=P.device))
    gate = ops.gate_mlp(avg, w1, zb, w2, zb, ctx.f32(n, c), kind=0)
    y = ops.mlca_apply(t, gate, 1, ctx.act(n, h, w, c), add=x)
    u = simple_feature_processor(ctx, p + ".pre_ffn_block", y)
    f = conv(ctx, u, P.conv(p + ".ffn.cv1.weight"), act=ops.ACT_GELU)
    return conv(ctx, f, P.conv(p + ".ffn.cv2.weight"), out=out, alpha=rw2, add=y)


def c2sfa(ctx, p, x, n=1):
    """nn/modules/block.py:2358-2373 C2SFA + C2PSA.forward :1045-1049 (split, n ProgressiveTSSA_Fusion0 blocks on b, concat-free cv2)"""
    c = ctx.P.conv_bn(p + ".cv1").cout // 2
    ab = conv_bn_act(ctx, p + ".cv1", x)
    b = ab.slice(c, c)
    cur = b
    for i in range(n):
        cur = progressive_tssa_fusion0(ctx, f"{p}.m.{i}", cur, out=b if i == n - 1 else None)
    return conv_bn_act(ctx, p + ".cv2", ab)


def fusion_bifpn(ctx, p, xs):
    """nn/modules/block.py:1532-1535 Fusion('bifpn') for two inputs"""
    w = torch.relu(ctx.P.sd[p + ".fusion_weight"].float())
    w = (w / (w.sum() + 1e-4)).tolist()
    assert len(xs) == 2
    return ops.eltwise(0, xs[0], xs[1], ctx.act(xs[0].n, xs[0].h, xs[0].w, xs[0].c), alpha=w[0], beta=w[1])


# --------------------------------------------------------------------------------------------------------------------
# layer 10: C2ProgressiveTSSA_Fusion
# --------------------------------------------------------------------------------------------------------------------
def progressive_feature_fusion(ctx, p, x):
    """nn/modules/block.py:2579-2630 ProgressiveFeatureFusion"""
    P = ctx.P
    n, h, w, c = x.n, x.h, x.w, x.c
    outs, cur = [], x
    for i in range(3):
        q = f"{p}.stages.{i}"
        tmp = done = None
        if i < 2:  # stage_fusion over cat([cur, o]) as two accumulating 1x1 convs (no concat buffer); the half on `cur` does not depend on this
            # stage's own chain: under graph capture it runs on a side stream (a parallel branch), one launch less on the critical path per stage
            wk, bk = f"{p}.stage_fusion.{i}.weight", f"{p}.stage_fusion.{i}.bias"
            wfull = P.sd[wk].float()
            c1 = P.conv_raw(wk + "#a", wfull[:, :c], P.sd[bk].float())
            c2 = P.conv_raw(wk + "#b", wfull[:, c:], None)
            if ctx.parallel_levels:
                main = torch.cuda.current_stream()
                fork = torch.cuda.Event()
                fork.record(main)
                br = ctx.side_streams(8)[7]
                br.wait_event(fork)
                with torch.cuda.stream(br):
                    tmp = conv(ctx, cur, c1)
                    done = torch.cuda.Event()
                    done.record(br)
            else:
                tmp = conv(ctx, cur, c1)
        dw3, b3, _ = P.dw(q + ".conv.weight", q + ".conv.bias")
        sc, sh = P.bn_affine(q + ".norm")
        t = ops.dwconv(cur, dw3, ctx.act(n, h, w, c), bias=b3, scale=sc, shift=sh, k=3, act=ops.ACT_GELU)
        dw7, b7, _ = P.dw(q + ".spatial_mix.weight", q + ".spatial_mix.bias")
        sm = ops.dwconv(t, dw7, ctx.act(n, h, w, c), bias=b7, k=7, add=cur)  # spatial_mix(t) + cur
        o = conv(ctx, t, P.conv(q + ".channel_mix.weight", q + ".channel_mix.bias"), add=sm)  # + channel_mix(t)
        outs.append(o)
        if i < 2:
            if done is not None:
                torch.cuda.current_stream().wait_event(done)
            cur = conv(ctx, o, c2, add=tmp)
    sa = P.vec(p + ".stage_attention")
    return ops.eltwise(3, outs[0], outs[1], ctx.act(n, h, w, c), c3=outs[2], d4=x, alpha=sa[0], beta=sa[1], gamma=sa[2])


def adaptive_dynamic_tanh(ctx, p, x):
    """nn/modules/block.py:2493-2577 AdaptiveDynamicTanh"""
    P = ctx.P
    g = ops.gap(x, ctx.f32(x.n, x.c))
    w1 = P.misc(p + ".ig1", lambda: P.f32(p + ".importance_gate.1.weight").reshape(-1, x.c).contiguous())
    w2 = P.misc(p + ".ig3", lambda: P.f32(p + ".importance_gate.3.weight").reshape(3, -1).contiguous())
    imp = ops.gate_mlp(g, w1, P.f32(p + ".importance_gate.1.bias"), w2, P.f32(p + ".importance_gate.3.bias"), ctx.f32(x.n, 3), kind=1)
    alphas = P.misc(p + ".al", lambda: P.f32(p + ".alphas").reshape(-1).contiguous())
    return ops.adt_apply(x, imp, alphas, P.f32(p + ".weight"), P.f32(p + ".bias"), ctx.act(x.n, x.h, x.w, x.c))


def cross_scale_attention_tssa(ctx, p, x, identity, rw, heads=2, scales=(1, 2, 4)):
    """nn/modules/block.py:2417-2491 CrossScaleAttentionTSSA, fused with `identity + attn * residual_weight1` (block.py:2680-2683).
    out_proj (of nn.MultiheadAttention), the mean over scales and to_out are all linear, so they run as ONE 1x1 conv after the mean."""
    P = ctx.P
    n, h, w, c = x.n, x.h, x.w, x.c
    T = h * w
    st = ctx.act(n, len(scales) * T, 1, c)
    temps = P.misc(p + ".temps", lambda: P.f32(p + ".temps").reshape(len(scales), heads).contiguous())
    def scale_branch(i, s):
        xs = x if s == 1 else ops.pool_upsample(x, s, ctx.act(n, h, w, c))
        qkv = conv(ctx, xs, P.conv(f"{p}.qkv_projections.{i}.weight"))
        ops.tssa(qkv, temps[i], heads, st, i * T)

    if ctx.parallel_levels and len(scales) > 1:  # the scales are independent (disjoint token ranges of `st`): parallel graph branches
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        side = ctx.side_streams(2 + len(scales) - 1)[2:]
        joins = []
        for i in range(1, len(scales)):
            side[i - 1].wait_event(fork)
            with torch.cuda.stream(side[i - 1]):
                scale_branch(i, scales[i])
                ev = torch.cuda.Event()
                ev.record(side[i - 1])
                joins.append(ev)
        scale_branch(0, scales[0])
        for ev in joins:
            cur.wait_event(ev)
    else:
        for i, s in enumerate(scales):
            scale_branch(i, s)
    q = p + ".cross_scale_fusion"
    qkv2 = conv(ctx, st, P.conv(q + ".in_proj_weight", q + ".in_proj_bias"))
    ao = ops.mha(qkv2, heads, ctx.act(n, len(scales) * T, 1, c))
    am = ops.group_mean(ao, len(scales), ctx.act(n, h, w, c))

    def fold():
        wo, bo = P.sd[q + ".out_proj.weight"].double(), P.sd[q + ".out_proj.bias"].double()
        wt, bt = P.sd[p + ".to_out.0.weight"].double(), P.sd[p + ".to_out.0.bias"].double()
        return (wt @ wo).float()[:, :, None, None], (wt @ bo + bt).float()

    wf, bf = P.misc(p + ".fold", fold)
    return conv(ctx, am, P.conv_raw(p + ".out_fold", wf, bf), alpha=rw, add=identity)


def edffn(ctx, p, x, residual, rw):
    """nn/modules/block.py:2376-2415 EDFFN, fused with `x + ffn * residual_weight2` (block.py:2694-2697); writes into `residual`'s shape"""
    P = ctx.P
    n, h, w = x.n, x.h, x.w
    t = conv(ctx, x, P.conv(p + ".project_in.weight"))
    dw, _, _ = P.dw(p + ".dwconv.weight")
    g = ops.dwconv(t, dw, ctx.act(n, h, w, t.c // 2), k=3, gate_split=t.c // 2)
    o = conv(ctx, g, P.conv(p + ".project_out.weight"))
    m = P.misc(p + ".spectral", lambda: edffn_spectral_matrix(P.sd[p + ".fft"]).to(ctx.device))
    return o, m


def progressive_tssa_fusion(ctx, p, x, out):
    """nn/modules/block.py:2632-2698 ProgressiveTSSA_Fusion.forward (shortcut=True); result written into `out`"""
    P = ctx.P
    t = progressive_feature_fusion(ctx, p + ".progressive_fusion1", x)
    t = adaptive_dynamic_tanh(ctx, p + ".dyt1", t)
    t = cross_scale_attention_tssa(ctx, p + ".attn", t, identity=x, rw=P.scalar(p + ".residual_weight1"))
    t = progressive_feature_fusion(ctx, p + ".progressive_fusion2", t)
    f = adaptive_dynamic_tanh(ctx, p + ".dyt2", t)
    o, m = edffn(ctx, p + ".ffn", f, t, None)
    return ops.patch_filter(o, m, out, alpha=P.scalar(p + ".residual_weight2"), add=t)


def c2ptssa(ctx, p, x):
    """nn/modules/block.py:2700-2710 C2ProgressiveTSSA_Fusion + C2PSA.forward :1045-1049"""
    c = ctx.P.conv_bn(p + ".cv1").cout // 2
    ab = conv_bn_act(ctx, p + ".cv1", x)
    b = ab.slice(c, c)
    progressive_tssa_fusion(ctx, p + ".m.0", b, out=b)  # b is dead once the attention residual has consumed it
    return conv_bn_act(ctx, p + ".cv2", ab)


# --------------------------------------------------------------------------------------------------------------------
# AYHead1
# --------------------------------------------------------------------------------------------------------------------
def task_decomposition(ctx, p, feat, avg, out):
    """nn/modules/head.py:626-669 TaskDecomposition (stacked_convs = 1): per-image scalar gate on a 1x1 conv, then GN + SiLU"""
    P = ctx.P
    w1 = P.misc(p + ".la1", lambda: P.f32(p + ".la_conv1.weight").reshape(-1, feat.c).contiguous())
    w2 = P.misc(p + ".la2", lambda: P.f32(p + ".la_conv2.weight").reshape(1, -1).contiguous())
    gate = ops.gate_mlp(avg, w1, P.f32(p + ".la_conv1.bias"), w2, P.f32(p + ".la_conv2.bias"), ctx.f32(feat.n, 1), kind=0)
    return conv_gn_act(ctx, p + ".reduction_conv", feat, out=out, img_scale=gate)


def coord_att(ctx, p, x, out=None):
    """nn/modules/head.py:671-707 CoordAtt (conv1 + bn1 folded, h-swish, per-axis 1x1 + sigmoid gates)"""
    P = ctx.P
    n, h, w, c = x.n, x.h, x.w, x.c

    def fold():
        return fold_bn(P.sd[p + ".conv1.weight"].float(), P.sd[p + ".bn1.weight"].float(), P.sd[p + ".bn1.bias"].float(),
                       P.sd[p + ".bn1.running_mean"].float(), P.sd[p + ".bn1.running_var"].float(), P.sd[p + ".conv1.bias"].float())

    wf, bf = P.misc(p + ".fold", fold)
    rows, cols = ctx.act(n, h, 1, c), ctx.act(n, w, 1, c)
    ops.rowcol_mean(x, rows, cols)
    mip, oup = wf.shape[0], P.sd[p + ".conv_h.weight"].shape[0]
    if mip in (8, 16, 32) and oup % 8 == 0:
        # one launch for conv1 (+ bn1) -> h-swish -> conv_h / conv_w -> sigmoid on both axes (fp32 weights and hidden activations)
        dev = lambda t: t.float().contiguous().to(ctx.device)  # noqa: E731
        w1, b1 = P.misc(p + ".mlp1", lambda: (dev(wf.reshape(mip, c)), dev(bf)))
        wh, ww = P.misc(p + ".mlp2", lambda: (dev(P.sd[p + ".conv_h.weight"].reshape(oup, mip)), dev(P.sd[p + ".conv_w.weight"].reshape(oup, mip))))
        gh, gw = ops.coordatt_mlp(rows, cols, w1, b1, wh, P.f32(p + ".conv_h.bias"), ww, P.f32(p + ".conv_w.bias"),
                                  ctx.act(n, h, 1, oup), ctx.act(n, w, 1, oup))
    else:
        c1 = P.conv_raw(p + ".conv1f", wf, bf)
        gh = conv(ctx, conv(ctx, rows, c1, act=ACT_HARDSWISH), P.conv(p + ".conv_h.weight", p + ".conv_h.bias"), act=ACT_SIGMOID)
        gw = conv(ctx, conv(ctx, cols, c1, act=ACT_HARDSWISH), P.conv(p + ".conv_w.weight", p + ".conv_w.bias"), act=ACT_SIGMOID)
    return ops.rowcol_gate(x, gh, gw, out if out is not None else ctx.act(n, h, w, c))


def ayhead_level(ctx, p, x, i):
    """nn/modules/head.py:1131-1175: one pyramid level of AYHead1.forward -> raw (n, h, w, 4*reg_max + nc)"""
    P = ctx.P
    n, h, w = x.n, x.h, x.w
    ad = conv_gn_act(ctx, f"{p}.stems.{i}", x)
    feat = conv_gn_act(ctx, p + ".share_conv.1", conv_gn_act(ctx, p + ".share_conv.0", ad))
    fc = feat.c
    avg = ops.gap(feat, ctx.f32(n, fc))
    crc, crr = ctx.act(n, h, w, 2 * fc), ctx.act(n, h, w, 2 * fc)  # [cls | reg_to_cls(reg)] and [reg | cls_to_reg(cls)]
    cls = task_decomposition(ctx, p + ".cls_decomp", feat, avg, crc.slice(0, fc))
    reg = task_decomposition(ctx, p + ".reg_decomp", feat, avg, crr.slice(0, fc))
    # CrossTaskInteraction head.py:1319-1333
    q = p + ".cross_task"
    conv(ctx, cls, P.conv(q + ".cls_to_reg.weight", q + ".cls_to_reg.bias"), out=crr.slice(fc, fc))
    conv(ctx, reg, P.conv(q + ".reg_to_cls.weight", q + ".reg_to_cls.bias"), out=crc.slice(fc, fc))
    cls2 = conv(ctx, crc, P.conv(q + ".cls_gate.0.weight", q + ".cls_gate.0.bias"), act=ACT_SIGMOID, mul=crc.slice(fc, fc), add=cls)
    reg2 = conv(ctx, crr, P.conv(q + ".reg_gate.0.weight", q + ".reg_gate.0.bias"), act=ACT_SIGMOID, mul=crr.slice(fc, fc), add=reg)
    # the classification tail (ResidualBlockGN + cls_prob) and the regression tail (DyDCNv2 + CoordAtt) are independent: the former runs on a
    # side stream of this level (a parallel branch of the captured graph)
    branch = None
    if ctx.parallel_levels:
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        branch = ctx.side_streams(7 + i)[4 + i]
        branch.wait_event(fork)

    def cls_tail():
        # ResidualBlockGN head.py:1031-1047
        ce = conv_gn_act(ctx, p + ".rep_block_cls.conv2", conv_gn_act(ctx, p + ".rep_block_cls.conv1", cls2), add=cls2)
        # cls_prob head.py:1168-1169
        c1 = conv(ctx, feat, P.conv(p + ".cls_prob_conv.0.weight", p + ".cls_prob_conv.0.bias"), act=ACT_RELU)
        return ce, conv(ctx, c1, P.conv(p + ".cls_prob_conv.2.weight", p + ".cls_prob_conv.2.bias"), act=ACT_SIGMOID)  # channel 0 of 8

    if branch is not None:
        with torch.cuda.stream(branch):
            cls_e, cp = cls_tail()
            cls_done = torch.cuda.Event()
            cls_done.record(branch)
    else:
        cls_e, cp = cls_tail()
    # DyDCNv2 head.py:751-782 (offsets / mask from `feat`, head.py:1155-1159) + GroupNorm(16) + CoordAtt
    om = conv(ctx, feat, P.conv(p + ".spatial_conv_offset.weight", p + ".spatial_conv_offset.bias"))
    dstats = ctx.f64(n, 16, 2)
    dw = P.conv(p + ".DyDCNV2.conv.weight")
    if ctx.dcn_col:
        # sampled, mask-weighted column tensor once (elementwise gather at full occupancy), then a plain 1x1 GEMM on the TMA-fed tcgen05 path:
        # [co][9][ci] weights are [co][1][9*ci] for the tap-major column tensor.  Faster in isolation than the gathered-operand deformable mode
        # (0.47 vs 0.82 ms at 80x80, batch 64) but not inside the step, where the pyramid levels overlap on parallel graph branches and the 9x
        # wider temporary costs bandwidth (8.00 vs 7.95 ms): off by default, the training path uses it (its backward needs the column tensor).
        col = ops.deform_col(reg2, om, ctx.act(n, h, w, 9 * fc))
        ra = ops.conv2d(col, dw.w, ctx.act(n, h, w, dw.cout), impl=ctx.conv_impl, gn_stats=dstats, gn_groups=16)
    else:
        ra = conv(ctx, reg2, dw, mode=ops.CONV_DEFORM, offmask=om, gn_stats=dstats, gn_groups=16)
    ra = ops.group_norm(ra, ctx.act(n, h, w, fc), dstats, 16, P.f32(p + ".DyDCNV2.norm.weight"), P.f32(p + ".DyDCNV2.norm.bias"),
                        GN_EPS, ACT_NONE, stats_ready=True)
    reg_e = coord_att(ctx, p + ".coord_attention_reg", ra)
    if branch is not None:
        torch.cuda.current_stream().wait_event(cls_done)
    cv2, cv3 = P.conv(p + ".cv2.weight", p + ".cv2.bias"), P.conv(p + ".cv3.weight", p + ".cv3.bias")
    out = ctx.act(n, h, w, cv2.cout + cv3.cout)
    conv(ctx, reg_e, cv2, out=out.slice(0, cv2.cout), alpha=P.scalar(f"{p}.scale.{i}.scale"))
    conv(ctx, cls_e, cv3, out=out.slice(cv2.cout, cv3.cout), pix_scale=cp)  # cv3(cls * cls_prob): the 1x1 conv commutes with a per-pixel scalar
    return out


def ayhead(ctx, p, xs, strides=(8, 16, 32), nc=80, reg_max=16, decode=True):
    """nn/modules/head.py:1127-1204 AYHead1.forward -> (y (B, 4+nc, N) fp32, [raw level outputs])"""
    # the pyramid levels are independent until decode: run P4 / P5 (small grids) on side streams next to P3 -- fork / join with events,
    # which a CUDA-graph capture records as parallel branches
    if ctx.parallel_levels and len(xs) > 1:
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        side = ctx.side_streams(len(xs) - 1)
        outs, joins = [None] * len(xs), []
        for i in range(1, len(xs)):
            side[i - 1].wait_event(fork)
            with torch.cuda.stream(side[i - 1]):
                outs[i] = ayhead_level(ctx, p, xs[i], i)
                ev = torch.cuda.Event()
                ev.record(side[i - 1])
                joins.append(ev)
        outs[0] = ayhead_level(ctx, p, xs[0], 0)
        for ev in joins:
            cur.wait_event(ev)
    else:
        outs = [ayhead_level(ctx, p, x, i) for i, x in enumerate(xs)]
    if not decode:
        return None, outs
    n_anchors = sum(o.h * o.w for o in outs)
    y = ctx.f32(outs[0].n, 4 + nc, n_anchors)
    proj = ctx.P.misc(p + ".proj", lambda: ctx.P.f32(p + ".dfl.conv.weight").reshape(-1).contiguous())
    ops.decode(outs, strides, nc, reg_max, proj, y)
    return y, outs


# --------------------------------------------------------------------------------------------------------------------
# whole model (z-yaml/yolo11-701-YOLO-AD-Refine.yaml at scale n; nn/tasks.py:141-168 _predict_once)
# --------------------------------------------------------------------------------------------------------------------
def stem(ctx, p, img):
    """Layer 0: Conv(3 -> 16, k3, s2) + BN + SiLU straight from the NCHW image.  uint8 images carry the predictor's /255 (engine/predictor.py:129-133)
    folded into the weights.  Falls back to layout conversion + the generic conv for any other first layer."""
    P = ctx.P
    n, cin, H, W = img.shape
    w0 = P.sd[p + ".conv.weight"]
    if tuple(w0.shape[1:]) == (cin, 3, 3) and w0.shape[0] == 16 and cin <= 4:
        u8 = img.dtype == torch.uint8

        def fold():
            wf, bf = w0.float(), None
            if p + ".bn.weight" in P.sd:
                wf, bf = fold_bn(wf, P.sd[p + ".bn.weight"].float(), P.sd[p + ".bn.bias"].float(), P.sd[p + ".bn.running_mean"].float(),
                                 P.sd[p + ".bn.running_var"].float(), P.sd.get(p + ".conv.bias"))
            else:
                bf = P.sd[p + ".conv.bias"].float()
            if u8:
                wf = wf / 255.0
            return (wf.permute(0, 2, 3, 1).reshape(16, -1).contiguous().to(ctx.device), bf.contiguous().to(ctx.device))

        wq, bq = P.misc(p + (".stem_u8" if u8 else ".stem_f32"), fold)
        return ops.stem_conv(img, wq, bq, ctx.act(n, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 16))
    x = ops.u8_to_nhwc(img, ctx.act(n, H, W, 8)) if img.dtype == torch.uint8 else ops.nchw_to_nhwc(img, ctx.act(n, H, W, 8))
    return conv_bn_act(ctx, p, x, 2)


def forward_model(ctx, img, decode=True, keep_layers=False):
    """img: fp32 (n, 3, H, W) in [0, 1], or uint8 (n, 3, H, W) in [0, 255], on the device.  Returns (y, raw_levels[, layer outputs])."""
    n, _, H, W = img.shape
    assert H % 32 == 0 and W % 32 == 0, "image size must be a multiple of the maximum stride 32"
    L = {}
    L[0] = stem(ctx, "model.0", img)
    L[1] = conv_bn_act(ctx, "model.1", L[0], 2)
    L[2] = c3k2(ctx, "model.2", L[1])
    L[3] = conv_bn_act(ctx, "model.3", L[2], 2)
    L[4] = c3k2(ctx, "model.4", L[3])
    # the neck's lateral gates ELA_HSFPN(L4) / ELA_HSFPN(L6) (yaml layers 21 / 14) depend only on the backbone maps: they run on side streams
    # (graph branches) next to the latency-bound 20x20 tail of the backbone and are joined where the neck consumes them
    side = ctx.side_streams(2) if ctx.parallel_levels else None
    joins = {}

    def fork(idx, stream, fn):
        if side is None:
            L[idx] = fn()
            return
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        stream.wait_event(ev)
        with torch.cuda.stream(stream):
            L[idx] = fn()
            done = torch.cuda.Event()
            done.record(stream)
        joins[idx] = done

    def join(idx):
        if idx in joins:
            torch.cuda.current_stream().wait_event(joins.pop(idx))

    fork(21, side[0] if side else None, lambda: ela_hsfpn(ctx, "model.21", L[4], True))
    L[5] = conv_bn_act(ctx, "model.5", L[4], 2)
    L[6] = c3k2(ctx, "model.6", L[5], True, True)
    fork(14, side[1] if side else None, lambda: ela_hsfpn(ctx, "model.14", L[6], True))
    L[7] = conv_bn_act(ctx, "model.7", L[6], 2)
    L[8] = c3k2(ctx, "model.8", L[7], True, True)
    L[9] = sppf(ctx, "model.9", L[8])
    L[10] = c2ptssa(ctx, "model.10", L[9])
    P = ctx.P
    # neck (HS-FPN): lateral 1x1 * gate + upsampled, fused into the 1x1 conv epilogue (yaml layers 15-18 and 22-25)
    L[11] = ela_hsfpn(ctx, "model.11", L[10], True)
    L[12] = conv(ctx, L[11], P.conv("model.12.weight", "model.12.bias"))
    L[13] = conv(ctx, L[12], P.conv("model.13.weight", "model.13.bias", stride=2, transposed=True), mode=ops.CONV_TRANSPOSED)
    # ELA_HSFPN(flag=False) (16 / 23) hands its two 1-D gates to the lateral 1x1 convolution, whose epilogue does Multiply (17 / 24) + Add (18 / 25);
    # keep_layers (the per-layer parity tests) materialises the gate map the yaml's layer 16 / 23 would output
    g16 = ela_hsfpn(ctx, "model.16", L[13], False, factors=True)
    if keep_layers:
        L[16] = ops.rowcol_gate(None, g16[0], g16[1], ctx.act(L[13].n, L[13].h, L[13].w, L[13].c))
    join(14)
    L[18] = conv(ctx, L[14], P.conv("model.15.weight", "model.15.bias"), gate=g16, add=L[13])  # Multiply (17) + Add (18)
    L[19] = c3k2(ctx, "model.19", L[18], False, True)
    L[20] = conv(ctx, L[19], P.conv("model.20.weight", "model.20.bias", stride=2, transposed=True), mode=ops.CONV_TRANSPOSED)
    g23 = ela_hsfpn(ctx, "model.23", L[20], False, factors=True)
    if keep_layers:
        L[23] = ops.rowcol_gate(None, g23[0], g23[1], ctx.act(L[20].n, L[20].h, L[20].w, L[20].c))
    join(21)
    L[25] = conv(ctx, L[21], P.conv("model.22.weight", "model.22.bias"), gate=g23, add=L[20])  # Multiply (24) + Add (25)
    L[26] = c3k2(ctx, "model.26", L[25], False, True)
    L[27] = conv_bn_act(ctx, "model.27", L[26], 2)
    L[28] = fusion_bifpn(ctx, "model.28", [L[27], L[19]])
    L[29] = c3k2(ctx, "model.29", L[28])
    L[30] = conv_bn_act(ctx, "model.30", L[29], 2)
    L[31] = fusion_bifpn(ctx, "model.31", [L[30], L[12]])
    L[32] = c3k2(ctx, "model.32", L[31])
    y, outs = ayhead(ctx, "model.33", [L[26], L[29], L[32]], decode=decode)
    return (y, outs, L) if keep_layers else (y, outs)
