"""Training path of the YOLO-AD-Refine model: the train()-mode forward (batch-statistics BatchNorm, unfolded
weights), the detection loss and the complete backward pass, expressed as calls into libyad.so.

The reference's backward is PyTorch autograd over its nn.Modules (engine/trainer.py:382-397 `self.loss, self.loss_items = self.model(batch)`
-> `scaler.scale(self.loss).backward()`).  Here every forward primitive appends its hand-written backward (a closure over libyad kernels) to
a tape; `Graph.backward()` replays the tape in reverse.  This file contains no arithmetic: PyTorch allocates buffers and owns the streams.

Gradient buffers mirror the activation buffers (same storage offsets, same channel windows), so the concatenation-free layout of the forward
carries over: the gradient of a channel window of a `cat` buffer is the same window of the gradient buffer.  A window that has not been
written yet is written (acc = 0), afterwards accumulated into (acc = 1).
"""
import math

import torch

from . import ops
from .ops import ACT_GELU, ACT_HARDSWISH, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SILU, CONV_NORMAL, CONV_TRANSPOSED, Act
from .weights import BN_EPS, GN_EPS, gn_groups

BN_MOMENTUM = 0.03  # utils/torch_utils.py:426-436 initialize_weights


def _subtract(r, m):
    """rectangle r minus rectangle m; rectangles are (row0, row1, ch0, ch1) windows of one gradient buffer"""
    r0, r1, c0, c1 = r
    q0, q1, d0, d1 = m
    if q0 >= r1 or q1 <= r0 or d0 >= c1 or d1 <= c0:
        return [r]
    out = []
    if q0 > r0:
        out.append((r0, q0, c0, c1))
    if q1 < r1:
        out.append((q1, r1, c0, c1))
    a0, a1 = max(r0, q0), min(r1, q1)
    if d0 > c0:
        out.append((a0, a1, c0, d0))
    if d1 < c1:
        out.append((a0, a1, d1, c1))
    return out


class Graph:
    """one training step's tape + buffer helpers"""

    def __init__(self, tp, conv_impl=0, update_bn=True):
        self.tp, self.dtype, self.device, self.conv_impl, self.update_bn = tp, tp.dtype, tp.device, conv_impl, update_bn
        self.tape = []
        self._g = {}  # storage ptr -> (flat gradient tensor, owner buffer kept alive)
        self._w = {}  # storage ptr -> list of written windows (row0, row1, c0, c1)

    # ---- buffers -------------------------------------------------------------------------------------------------------
    def act(self, n, h, w, c, ld=None):
        return Act.empty(n, h, w, c, self.dtype, self.device, ld)

    def like(self, a):
        return Act.empty(a.n, a.h, a.w, a.c, self.dtype, self.device)

    def f32(self, *shape, zero=False):
        return (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=self.device)

    def f64(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.device)

    # ---- gradients ------------------------------------------------------------------------------------------------------
    def grad(self, a):
        """gradient view of an activation view, for WRITING (uninitialised memory: pair it with mark(), which says whether to accumulate)"""
        st = a.buf.untyped_storage()
        key = st.data_ptr()
        if key not in self._g:
            self._g[key] = (torch.empty(st.nbytes() // a.buf.element_size(), dtype=a.dtype, device=a.device), a.buf)
        flat = self._g[key][0]
        off = a.buf.storage_offset()
        return Act(flat[off:off + a.buf.numel()].view(a.buf.shape), a.c0, a.c)

    def gin(self, a):
        """gradient view of an activation view, for READING: whatever part of the window no consumer has written is zero (an output nobody
        used) and is zero-filled now."""
        ga = self.grad(a)
        key, win = self._window(a)
        rem = self._uncovered(key, win)
        if rem:
            self._zero(key, a.ld, rem)
            self._w.setdefault(key, []).extend(rem)
        return ga

    def _window(self, a):
        ld = a.ld
        r0 = a.buf.storage_offset() // ld
        return a.buf.untyped_storage().data_ptr(), (r0, r0 + a.buf.numel() // ld, a.c0, a.c0 + a.c)

    def _uncovered(self, key, win):
        rem = [win]
        for m in self._w.get(key, ()):
            rem = [piece for r in rem for piece in _subtract(r, m)]
            if not rem:
                break
        return rem

    def _zero(self, key, ld, rects):
        view = self._g[key][0].view(-1, ld)
        for r0, r1, c0, c1 in rects:
            view[r0:r1, c0:c1].zero_()

    def mark(self, a):
        """Declare a write to the gradient window of activation `a`.  Returns 0 when nothing of the window holds data yet (the caller writes),
        1 when it does (the caller accumulates); a partly written window has its unwritten remainder zero-filled first, so accumulating is
        always safe.  Gradient buffers are never memset wholesale."""
        key, win = self._window(a)
        rem = self._uncovered(key, win)
        self._w.setdefault(key, []).append(win)
        if rem == [win]:
            return 0
        if rem:
            self.grad(a)  # make sure the buffer exists
            self._zero(key, a.ld, rem)
        return 1

    def accumulate(self, a, src, alpha=1.0, alpha_dev=None):
        """grad(a) (+)= alpha * src"""
        ga = self.grad(a)
        acc = self.mark(a)
        ops.eltwise_dev(5, src, None, ga, c3=ga if acc else None, alpha=alpha, pa=alpha_dev)

    def backward(self):
        for fn in reversed(self.tape):
            fn()
        self.tape = []


# ------------------------------------------------------------------------------------------------------------------------
# primitives: forward + tape entry
# ------------------------------------------------------------------------------------------------------------------------
def _bias(g, key):
    """(parameter, gradient) fp32 device vectors of a bias, zero-padded to a multiple of 8 when needed"""
    tp = g.tp
    if math.prod(tp.shape[key]) % 8 == 0:
        return tp.p(key).reshape(-1), tp.g(key).reshape(-1)
    return tp.f32(key, "pad")


def conv(g, x, wkey, bkey=None, stride=1, mode=CONV_NORMAL, act=ACT_NONE, out=None, add=None, img_scale=None, gn_stats=None, groups=0,
         need_dx=True):
    """dense convolution + bias (+ activation | + add | x per-image gate); weights in torch layout under `wkey` (Conv2d, Linear, Conv1d as k x 1,
    ConvTranspose2d when mode = TRANSPOSED).  img_scale = (gate fp32 [n], dgate fp32 [n] accumulated).
    Backward = autograd of F.conv2d / F.conv_transpose2d / F.linear: bias gradient (column sums), weight gradient (yad_conv_wgrad), input
    gradient (yad_conv2d on dy with the permuted weight)."""
    tp = g.tp
    transposed = mode == CONV_TRANSPOSED
    W = tp.conv(wkey, "convT_fwd" if transposed else "fwd")
    kh, kw = W.kh, W.kw
    ph, pw = kh // 2, kw // 2
    assert x.c == W.cin, (wkey, x.c, W.cin)
    if transposed:
        ho, wo = 2 * x.h, 2 * x.w
    else:
        ho, wo = (x.h + 2 * ph - kh) // stride + 1, (x.w + 2 * pw - kw) // stride + 1
    y = out if out is not None else g.act(x.n, ho, wo, W.cout)
    assert (y.n, y.h, y.w, y.c) == (x.n, ho, wo, W.cout), (wkey, (y.n, y.h, y.w, y.c), (x.n, ho, wo, W.cout))
    assert not (act != ACT_NONE and (add is not None or img_scale is not None))
    b, gb = _bias(g, bkey) if bkey else (None, None)
    gate = img_scale[0] if img_scale is not None else None
    ops.conv2d(x, W.w, y, bias=b, kh=kh, kw=kw, stride=2 if transposed else stride, pad_h=ph, pad_w=pw, act=act, mode=mode, add=add, img_scale=gate,
               impl=g.conv_impl, gn_stats=gn_stats, gn_groups=groups)

    def bwd():
        dy = g.gin(y)
        if add is not None:
            g.accumulate(add, dy)
        if act != ACT_NONE:
            t = g.like(y)
            ops.act_bwd(y, dy, act, t, 0)
            dy = t
        if gate is not None:  # y = gate[n] * conv(x): d gate[n] = <dy, y> / gate[n]; d conv = gate[n] dy
            ops.dot(dy, y, img_scale[1], per_image=True, img_div=gate)
            dy = ops.scale_img(dy, gate, g.like(y))
        if gb is not None:
            ops.colsum(dy, gb)
        if transposed:  # weight gradient of ConvTranspose2d = wgrad of the stride-2 conv that maps y back to x, operands swapped
            Wt = tp.conv(wkey, "convT_dgrad")
            ops.conv_wgrad(dy, x, Wt.gw, kh, kw, 2, ph, pw, impl=g.conv_impl)
        else:
            ops.conv_wgrad(x, dy, W.gw, kh, kw, stride, ph, pw, impl=g.conv_impl)
        if need_dx:
            dx = g.grad(x)
            acc = g.mark(x)
            if transposed:
                ops.conv2d(dy, Wt.w, dx, kh=kh, kw=kw, stride=2, pad_h=ph, pad_w=pw, add=dx if acc else None, impl=g.conv_impl)
            elif stride == 1:
                Wd = tp.conv(wkey, "dgrad")
                ops.conv2d(dy, Wd.w, dx, kh=kh, kw=kw, stride=1, pad_h=kh - 1 - ph, pad_w=kw - 1 - pw, add=dx if acc else None, impl=g.conv_impl)
            else:
                Wd = tp.conv(wkey, "dgrad_t")
                ops.conv2d(dy, Wd.w, dx, kh=kh, kw=kw, stride=2, pad_h=ph, pad_w=pw, mode=CONV_TRANSPOSED, add=dx if acc else None,
                           impl=g.conv_impl)

    if transposed:
        tp.conv(wkey, "convT_dgrad")  # register the layouts before the weights are packed
    elif need_dx:
        tp.conv(wkey, "dgrad" if stride == 1 else "dgrad_t")
    g.tape.append(bwd)
    return y


def norm_act(g, t, wkey, bkey, groups, eps, act, out=None, add=None, stats=None, bn=None):
    """y = act(normalise(t) * gamma + beta) (+ add).  groups > 0: GroupNorm per image (nn/modules/head.py:1276); bn = key prefix: BatchNorm2d with
    batch statistics (nn.Module.train(); conv.py:50), i.e. the whole batch viewed as one image with one channel per group, plus the running
    statistics update.  stats: double (n, groups, 2) already filled by the producing conv's epilogue (GroupNorm only)."""
    tp = g.tp
    gamma, beta = tp.p(wkey), tp.p(bkey)
    y = out if out is not None else g.like(t)
    if bn is not None:
        tv, yv = t.reshape(1, t.n * t.h, t.w), y.reshape(1, y.n * y.h, y.w)
        addv = add.reshape(1, add.n * add.h, add.w) if add is not None else None
        groups = t.c
    else:
        tv, yv, addv = t, y, add
    if stats is None:
        stats = g.f64(tv.n, groups, 2)
        ops.group_norm(tv, yv, stats, groups, gamma, beta, eps, act, addv)
    else:
        ops.group_norm(tv, yv, stats, groups, gamma, beta, eps, act, addv, stats_ready=True)
    if bn is not None and g.update_bn:
        ops.bn_running_update(stats, t.c, t.n * t.h * t.w, BN_MOMENTUM, tp.buf(bn + ".running_mean"), tp.buf(bn + ".running_var"))

    def bwd():
        dy = g.gin(y)
        if add is not None:
            g.accumulate(add, dy)
        dt = g.grad(t)
        acc = g.mark(t)
        dyv, dtv = (dy.reshape(1, y.n * y.h, y.w), dt.reshape(1, t.n * t.h, t.w)) if bn is not None else (dy, dt)
        ops.norm_bwd(tv, dyv, stats, groups, gamma, beta, eps, act, g.f64(tv.n, groups, 2), tp.g(wkey), tp.g(bkey), dtv, acc)

    g.tape.append(bwd)
    return y


def conv_bn_act(g, p, x, stride=1, out=None, add=None, need_dx=True):
    """nn/modules/conv.py:36-51 Conv.forward in train(): conv(bias=False) -> BatchNorm2d(batch statistics) -> SiLU (+ residual)"""
    co, _, kh, _ = g.tp.shape[p + ".conv.weight"]
    if kh == 3 and x.c <= 32 and co <= 32:  # small-channel 3x3: the single-pass conv_small kernel (no fused statistics) + yad_gn_stats
        t = conv(g, x, p + ".conv.weight", stride=stride, need_dx=need_dx)
        return norm_act(g, t, p + ".bn.weight", p + ".bn.bias", 0, BN_EPS, ACT_SILU, out=out, add=add, bn=p + ".bn")
    # per-channel batch statistics accumulated in the convolution's epilogue (gn_groups = 0): saves one full read of the conv output
    stats = g.f64(1, (co + 7) // 8 * 8, 2)
    t = conv(g, x, p + ".conv.weight", stride=stride, need_dx=need_dx, gn_stats=stats, groups=0)
    return norm_act(g, t, p + ".bn.weight", p + ".bn.bias", 0, BN_EPS, ACT_SILU, out=out, add=add, stats=stats, bn=p + ".bn")


def conv_gn_act(g, p, x, out=None, add=None, act=ACT_SILU, img_scale=None):
    """nn/modules/head.py:1265-1279 Conv_GN"""
    groups = gn_groups(g.tp.shape[p + ".conv.weight"][0])
    stats = g.f64(x.n, groups, 2)
    t = conv(g, x, p + ".conv.weight", img_scale=img_scale, gn_stats=stats, groups=groups)
    return norm_act(g, t, p + ".gn.weight", p + ".gn.bias", groups, GN_EPS, act, out=out, add=add, stats=stats)


def dwconv(g, x, wkey, bkey, k, out=None, add=None):
    """depthwise k x k conv + bias (+ add).  Backward: taps gradient, bias gradient, input gradient = the same kernel with flipped taps."""
    tp = g.tp
    w, gw = tp.f32(wkey, "dw")
    wflip, _ = tp.f32(wkey, "dw_flip")
    b, gb = (tp.p(bkey), tp.g(bkey)) if bkey else (None, None)
    y = out if out is not None else g.like(x)
    ops.dwconv(x, w, y, bias=b, k=k, add=add)

    def bwd():
        dy = g.gin(y)
        if add is not None:
            g.accumulate(add, dy)
        if gb is not None:
            ops.colsum(dy, gb)
        ops.dwconv_wgrad(x, dy, k, gw)
        dx = g.grad(x)
        acc = g.mark(x)
        ops.dwconv(dy, wflip, dx, k=k, add=dx if acc else None)

    g.tape.append(bwd)
    return y


def mul_add(g, a, b, c3=None, out=None):
    """y = a * b (+ c3)"""
    y = out if out is not None else g.like(a)
    ops.eltwise_dev(4, a, b, y, c3=c3)

    def bwd():
        dy = g.gin(y)
        if c3 is not None:
            g.accumulate(c3, dy)
        for u, v in ((a, b), (b, a)):
            du = g.grad(u)
            acc = g.mark(u)
            ops.eltwise_dev(4, dy, v, du, c3=du if acc else None)

    g.tape.append(bwd)
    return y


def scale_add(g, a, skey, c3=None, out=None, index=0):
    """y = s * a (+ c3) with s a learnable scalar (element `index` of parameter `skey`): residual_weight1/2 (block.py:2683,2697), Scale (head.py)"""
    tp = g.tp
    s, gs = tp.p(skey).reshape(-1)[index:], tp.g(skey).reshape(-1)[index:]
    y = out if out is not None else g.like(a)
    ops.eltwise_dev(5, a, None, y, c3=c3, pa=s)

    def bwd():
        dy = g.gin(y)
        if c3 is not None:
            g.accumulate(c3, dy)
        ops.dot(dy, a, gs)
        g.accumulate(a, dy, alpha_dev=s)

    g.tape.append(bwd)
    return y


def copy(g, a, out):
    ops.eltwise_dev(5, a, None, out)
    g.tape.append(lambda: g.accumulate(a, g.gin(out)))
    return out


# ------------------------------------------------------------------------------------------------------------------------
# blocks
# ------------------------------------------------------------------------------------------------------------------------
def mlca(g, p, x, out, add=None):
    """nn/modules/block.py:1540-1584 MLCA (+ residual)"""
    tp = g.tp
    wg, wl = tp.p(p + ".conv.weight"), tp.p(p + ".conv_local.weight")
    k = wg.numel()
    local, att = g.f32(x.n, 25, x.c), g.f32(x.n, 25, x.c)
    ops.mlca(x, out, wg, wl, k, local, att, 5, 0.5, add)

    def bwd():
        dy = g.gin(out)
        if add is not None:
            g.accumulate(add, dy)
        dx = g.grad(x)
        acc = g.mark(x)
        ops.mlca_bwd(x, dy, local, att, wg, wl, k, g.f32(x.n, 25, x.c), g.f32(x.n, 25, x.c), g.f32(5, x.c), tp.g(p + ".conv.weight"),
                     tp.g(p + ".conv_local.weight"), dx, acc)

    g.tape.append(bwd)
    return out


def bottleneck(g, p, x, out=None, attention=False):
    """nn/modules/block.py:341-354 Bottleneck / :1586-1594 Bottleneck_MLCA"""
    t = conv_bn_act(g, p + ".cv1", x)
    if out is None:
        out = g.like(x)
    if attention:
        t2 = conv_bn_act(g, p + ".cv2", t)
        return mlca(g, p + ".attention", t2, out, add=x)
    return conv_bn_act(g, p + ".cv2", t, out=out, add=x)


def c3k(g, p, x, out, attention=False):
    """nn/modules/block.py:256-270 C3.forward / :742-750 C3k / :1596-1600 C3k_MLCA"""
    c_ = g.tp.shape[p + ".cv1.conv.weight"][0]
    cat = g.act(x.n, x.h, x.w, 2 * c_)
    a = conv_bn_act(g, p + ".cv1", x)
    a = bottleneck(g, p + ".m.0", a, attention=attention)
    bottleneck(g, p + ".m.1", a, out=cat.slice(0, c_), attention=attention)
    conv_bn_act(g, p + ".cv2", x, out=cat.slice(c_, c_))
    return conv_bn_act(g, p + ".cv3", cat, out=out)


def c3k2(g, p, x, use_c3k=False, attention=False):
    """nn/modules/block.py:232-247 C2f.forward / :731-739 C3k2 / :1602-1605 C3k2_MLCA"""
    c = g.tp.shape[p + ".cv1.conv.weight"][0] // 2
    cat = g.act(x.n, x.h, x.w, 3 * c)
    conv_bn_act(g, p + ".cv1", x, out=cat.slice(0, 2 * c))
    y1 = cat.slice(c, c)
    if use_c3k:
        c3k(g, p + ".m.0", y1, cat.slice(2 * c, c), attention)
    else:
        bottleneck(g, p + ".m.0", y1, out=cat.slice(2 * c, c), attention=attention)
    return conv_bn_act(g, p + ".cv2", cat)


def sppf(g, p, x):
    """nn/modules/block.py:177-196 SPPF; backward routes through the three chained 5x5 max-pools (first maximum in window order)"""
    c = g.tp.shape[p + ".cv1.conv.weight"][0]
    cat = g.act(x.n, x.h, x.w, 4 * c)
    y = [cat.slice(i * c, c) for i in range(4)]
    conv_bn_act(g, p + ".cv1", x, out=y[0])
    ops.sppf_pool(y[0], y[1], y[2], y[3])

    def bwd():
        carry = None
        for i in (2, 1, 0):  # gradient of pool output i+1 -> pool input i
            nxt = g.f32(x.n, x.h, x.w, c, zero=True)
            ops.maxpool5_bwd(y[i], g.gin(y[i + 1]), carry, nxt)
            carry = nxt
        ops.cast_acc(carry, g.grad(y[0]), g.mark(y[0]))

    g.tape.append(bwd)
    return conv_bn_act(g, p + ".cv2", cat)


def ela_hsfpn(g, p, x, flag=True):
    """nn/modules/block.py:1408-1424 ELA_HSFPN"""
    n, h, w, c = x.n, x.h, x.w, x.c
    wk, bk, gk, bek = p + ".conv1x1.0.weight", p + ".conv1x1.0.bias", p + ".conv1x1.1.weight", p + ".conv1x1.1.bias"
    out = g.act(n, h, w, c)

    def branch(means):
        stats = g.f64(means.n, 16, 2)
        t = conv(g, means, wk, bk, gn_stats=stats, groups=16)
        return norm_act(g, t, gk, bek, 16, GN_EPS, ACT_SIGMOID, stats=stats)

    if h == w:
        means = g.act(2 * n, h, 1, c)
        rows, cols = means.images(0, n), means.images(n, n)
    else:
        rows, cols = g.act(n, h, 1, c), g.act(n, w, 1, c)
    ops.rowcol_mean(x, rows, cols)

    def bwd_mean():
        dx = g.grad(x)
        acc = g.mark(x)
        ops.bcast_add(dx, acc, row=g.gin(rows), s_row=1.0 / w, col=g.gin(cols), s_col=1.0 / h)

    g.tape.append(bwd_mean)
    if h == w:
        gates = branch(means)
        gh, gw = gates.images(0, n), gates.images(n, n)
    else:
        gh, gw = branch(rows), branch(cols)
    ops.rowcol_gate(x if flag else None, gh, gw, out)

    def bwd_gate():
        dy = g.gin(out)
        g.mark(gh), g.mark(gw)
        if flag:
            dx = g.grad(x)
            acc = g.mark(x)
            ops.rowcol_gate_bwd(x, gh, gw, dy, dx, acc, g.grad(gh), g.grad(gw))
        else:
            ops.rowcol_gate_bwd(None, gh, gw, dy, None, 0, g.grad(gh), g.grad(gw))

    g.tape.append(bwd_gate)
    return out


def fusion_bifpn(g, p, xs):
    """nn/modules/block.py:1532-1535 Fusion('bifpn'), two inputs"""
    tp = g.tp
    key = p + ".fusion_weight"
    w = g.f32(8)
    ops.fusion_weights(tp.p(key), 2, w=w)
    y = g.like(xs[0])
    ops.eltwise_dev(0, xs[0], xs[1], y, pa=w[0:], pb=w[1:])

    def bwd():
        dy = g.gin(y)
        dw = g.f32(8, zero=True)
        for i in range(2):
            ops.dot(dy, xs[i], dw[i:])
            g.accumulate(xs[i], dy, alpha_dev=w[i:])
        ops.fusion_weights(tp.p(key), 2, dw=dw, dp=tp.g(key))

    g.tape.append(bwd)
    return y


# ---- layer 10 ----------------------------------------------------------------------------------------------------------------
def progressive_feature_fusion(g, p, x):
    """nn/modules/block.py:2579-2630 ProgressiveFeatureFusion"""
    n, h, w, c = x.n, x.h, x.w, x.c
    outs, cur = [], x
    for i in range(3):
        q = f"{p}.stages.{i}"
        u = dwconv(g, cur, q + ".conv.weight", q + ".conv.bias", 3)
        t = norm_act(g, u, q + ".norm.weight", q + ".norm.bias", 0, BN_EPS, ACT_GELU, bn=q + ".norm")
        sm = dwconv(g, t, q + ".spatial_mix.weight", q + ".spatial_mix.bias", 7, add=cur)
        if i < 2:  # stage_fusion reads cat([cur, out]): both are written into one buffer
            cat = g.act(n, h, w, 2 * c)
            copy(g, cur, cat.slice(0, c))
            o = conv(g, t, q + ".channel_mix.weight", q + ".channel_mix.bias", add=sm, out=cat.slice(c, c))
            cur = conv(g, cat, f"{p}.stage_fusion.{i}.weight", f"{p}.stage_fusion.{i}.bias")
        else:
            o = conv(g, t, q + ".channel_mix.weight", q + ".channel_mix.bias", add=sm)
        outs.append(o)
    y = scale_add(g, outs[0], p + ".stage_attention", c3=x, index=0)
    y = scale_add(g, outs[1], p + ".stage_attention", c3=y, index=1)
    return scale_add(g, outs[2], p + ".stage_attention", c3=y, index=2)


def adaptive_dynamic_tanh(g, p, x):
    """nn/modules/block.py:2493-2577 AdaptiveDynamicTanh"""
    tp = g.tp
    k1, b1, k2, b2 = (p + ".importance_gate.1.weight", p + ".importance_gate.1.bias", p + ".importance_gate.3.weight",
                      p + ".importance_gate.3.bias")
    w1 = tp.p(k1).reshape(-1, x.c)
    w2 = tp.p(k2).reshape(3, -1)
    avg = ops.gap(x, g.f32(x.n, x.c))
    imp = ops.gate_mlp(avg, w1, tp.p(b1), w2, tp.p(b2), g.f32(x.n, 3), kind=1)
    y = ops.adt_apply(x, imp, tp.p(p + ".alphas").reshape(-1), tp.p(p + ".weight"), tp.p(p + ".bias"), g.like(x))

    def bwd():
        dx = g.grad(x)
        acc = g.mark(x)
        dimp, davg = g.f32(x.n, 3), g.f32(x.n, x.c)
        ops.adt_bwd(x, g.gin(y), imp, tp.p(p + ".alphas").reshape(-1), tp.p(p + ".weight"), dx, acc, dimp, tp.g(p + ".alphas"), tp.g(p + ".weight"),
                    tp.g(p + ".bias"))
        ops.gate_mlp_bwd(avg, w1, tp.p(b1), w2, tp.p(b2), 1, dimp, davg, tp.g(k1), tp.g(b1), tp.g(k2), tp.g(b2))
        ops.bcast_add(dx, 1, img=davg, s_img=1.0 / (x.h * x.w))

    g.tape.append(bwd)
    return y


def pool_upsample(g, x, s):
    y = ops.pool_upsample(x, s, g.like(x))

    def bwd():
        dx = g.grad(x)
        acc = g.mark(x)
        ops.pool_upsample_bwd(g.gin(y), s, g.f32(x.n, x.h // s, x.w // s, x.c), dx, acc)

    g.tape.append(bwd)
    return y


def cross_scale_attention_tssa(g, p, x, identity, rw_key, heads=2, scales=(1, 2, 4)):
    """nn/modules/block.py:2417-2491 CrossScaleAttentionTSSA followed by `identity + attn * residual_weight1` (block.py:2680-2683)"""
    tp = g.tp
    n, h, w, c = x.n, x.h, x.w, x.c
    T = h * w
    S = len(scales)
    st = g.act(n, S * T, 1, c)
    temps, dtemps = tp.p(p + ".temps").reshape(S, heads), tp.g(p + ".temps").reshape(S, heads)
    for i, s in enumerate(scales):
        xs = x if s == 1 else pool_upsample(g, x, s)
        qkv = conv(g, xs, f"{p}.qkv_projections.{i}.weight")
        ops.tssa(qkv, temps[i], heads, st, i * T)

        def bwd(qkv=qkv, i=i):
            g.mark(qkv)
            ops.tssa_bwd(qkv, temps[i], heads, g.gin(st), i * T, g.grad(qkv), dtemps[i])

        g.tape.append(bwd)
    q = p + ".cross_scale_fusion"
    qkv2 = conv(g, st, q + ".in_proj_weight", q + ".in_proj_bias")
    ao = ops.mha(qkv2, heads, g.act(n, S * T, 1, c))

    def bwd_mha():
        g.mark(qkv2)
        ops.mha_bwd(qkv2, heads, ao, g.gin(ao), g.grad(qkv2), g.f32(n * heads, S * T, 2))

    g.tape.append(bwd_mha)
    po = conv(g, ao, q + ".out_proj.weight", q + ".out_proj.bias")
    am = ops.group_mean(po, S, g.act(n, h, w, c))

    def bwd_mean():
        ops.group_mean_bwd(g.gin(am), S, g.grad(po), g.mark(po))

    g.tape.append(bwd_mean)
    to = conv(g, am, p + ".to_out.0.weight", p + ".to_out.0.bias")
    return scale_add(g, to, rw_key, c3=identity)


def edffn_spectral_basis(device):
    """The 8x8-patch rfft2 -> * W -> irfft2 of EDFFN (block.py:2405-2409) is linear in W: returns the constant basis B fp32 [64 out * 64 in][40]
    with filter matrix m[o][i][ch] = sum_f B[o*64+i][f] * W[ch][f] (built once with torch.fft on the host: constant generation, not data path)."""
    eye = torch.eye(64).view(64, 8, 8)
    f = torch.fft.rfft2(eye)  # (64 in, 8, 5)
    basis = torch.zeros(64, 64, 40)
    for k in range(40):
        wk = torch.zeros(8, 5)
        wk.view(-1)[k] = 1.0
        out = torch.fft.irfft2(f * wk, s=(8, 8))  # (64 in, 8, 8)
        basis[:, :, k] = out.reshape(64, 64).t()  # (out, in)
    return basis.reshape(4096, 40).contiguous().to(device)


def edffn(g, p, x, residual, rw_key, out=None):
    """nn/modules/block.py:2376-2415 EDFFN followed by `x + ffn * residual_weight2` (block.py:2694-2697)"""
    tp = g.tp
    n, h, w = x.n, x.h, x.w
    t = conv(g, x, p + ".project_in.weight")
    d = dwconv(g, t, p + ".dwconv.weight", None, 3)
    hc = d.c // 2
    d1, d2 = d.slice(0, hc), d.slice(hc, hc)
    gt = g.act(n, h, w, hc)
    ops.eltwise_dev(7, d1, d2, gt)

    def bwd_gate():
        g.mark(d)
        ops.gelu_gate_bwd(d1, d2, g.gin(gt), g.grad(d1), g.grad(d2), 0)

    g.tape.append(bwd_gate)
    o = conv(g, gt, p + ".project_out.weight")
    c = o.c
    if not hasattr(tp, "_edffn_basis"):
        tp._edffn_basis = edffn_spectral_basis(g.device)
    basis = tp._edffn_basis
    m = g.f32(4096, c)
    ops.small_gemm(basis, tp.p(p + ".fft").reshape(c, 40), m, 4096, c, 40)
    f = ops.patch_filter(o, m, g.like(o))

    def bwd_filter():
        dxf, dm = g.f32(n, h, w, c), g.f32(4096, c, zero=True)
        ops.patch_filter_bwd(o, g.gin(f), m, 1.0, dxf, dm)
        ops.cast_acc(dxf, g.grad(o), g.mark(o))
        ops.small_gemm(dm, basis, tp.g(p + ".fft").reshape(c, 40), c, 40, 4096, trans_a=True, acc=True)

    g.tape.append(bwd_filter)
    return scale_add(g, f, rw_key, c3=residual, out=out)


def progressive_tssa_fusion(g, p, x, out=None):
    """nn/modules/block.py:2632-2698 ProgressiveTSSA_Fusion.forward (shortcut=True)"""
    t = progressive_feature_fusion(g, p + ".progressive_fusion1", x)
    t = adaptive_dynamic_tanh(g, p + ".dyt1", t)
    t = cross_scale_attention_tssa(g, p + ".attn", t, identity=x, rw_key=p + ".residual_weight1")
    t = progressive_feature_fusion(g, p + ".progressive_fusion2", t)
    f = adaptive_dynamic_tanh(g, p + ".dyt2", t)
    return edffn(g, p + ".ffn", f, t, p + ".residual_weight2", out=out)


def c2ptssa(g, p, x):
    """nn/modules/block.py:2700-2710 C2ProgressiveTSSA_Fusion + C2PSA.forward :1045-1049"""
    c = g.tp.shape[p + ".cv1.conv.weight"][0] // 2
    ab = conv_bn_act(g, p + ".cv1", x)
    cat = g.act(x.n, x.h, x.w, 2 * c)
    copy(g, ab.slice(0, c), cat.slice(0, c))
    progressive_tssa_fusion(g, p + ".m.0", ab.slice(c, c), out=cat.slice(c, c))
    return conv_bn_act(g, p + ".cv2", cat)


# ---- head --------------------------------------------------------------------------------------------------------------------
def task_decomposition(g, p, feat, avg, davg, out):
    """nn/modules/head.py:626-669 TaskDecomposition (stacked_convs = 1); davg: fp32 (n, c) gradient buffer of `avg` owned by this consumer"""
    tp = g.tp
    k1, b1, k2, b2 = p + ".la_conv1.weight", p + ".la_conv1.bias", p + ".la_conv2.weight", p + ".la_conv2.bias"
    w1, w2 = tp.p(k1).reshape(-1, feat.c), tp.p(k2).reshape(1, -1)
    bb2, gb2 = tp.p(b2), tp.g(b2)
    gate = ops.gate_mlp(avg, w1, tp.p(b1), w2, bb2, g.f32(feat.n, 1), kind=0)
    dgate = g.f32(feat.n, 1, zero=True)

    def bwd():
        ops.gate_mlp_bwd(avg, w1, tp.p(b1), w2, bb2, 0, dgate, davg, tp.g(k1), tp.g(b1), tp.g(k2), gb2)

    g.tape.append(bwd)
    return conv_gn_act(g, p + ".reduction_conv", feat, out=out, img_scale=(gate, dgate))


def coord_att(g, p, x):
    """nn/modules/head.py:671-707 CoordAtt in train(): bn1 normalises with the statistics of the concatenated (h + w) axis of the whole batch"""
    n, h, w, c = x.n, x.h, x.w, x.c
    buf = torch.empty((1, n * (h + w), 1, c), dtype=g.dtype, device=g.device)
    means = Act(buf)
    rows, cols = Act(buf[:, :n * h].view(n, h, 1, c)), Act(buf[:, n * h:].view(n, w, 1, c))
    ops.rowcol_mean(x, rows, cols)

    def bwd_mean():
        ops.bcast_add(g.grad(x), g.mark(x), row=g.gin(rows), s_row=1.0 / w, col=g.gin(cols), s_col=1.0 / h)

    g.tape.append(bwd_mean)
    t = conv(g, means, p + ".conv1.weight", p + ".conv1.bias")
    a = norm_act(g, t, p + ".bn1.weight", p + ".bn1.bias", 0, BN_EPS, ACT_HARDSWISH, bn=p + ".bn1")
    mip = a.c
    ah, aw = Act(a.buf[:, :n * h].view(n, h, 1, a.ld), 0, mip), Act(a.buf[:, n * h:].view(n, w, 1, a.ld), 0, mip)
    gh = conv(g, ah, p + ".conv_h.weight", p + ".conv_h.bias", act=ACT_SIGMOID)
    gw = conv(g, aw, p + ".conv_w.weight", p + ".conv_w.bias", act=ACT_SIGMOID)
    y = ops.rowcol_gate(x, gh, gw, g.like(x))

    def bwd_gate():
        g.mark(gh), g.mark(gw)
        dx = g.grad(x)
        acc = g.mark(x)
        ops.rowcol_gate_bwd(x, gh, gw, g.gin(y), dx, acc, g.grad(gh), g.grad(gw))

    g.tape.append(bwd_gate)
    return y


def deform_conv_gn(g, p, x, om):
    """nn/modules/head.py:751-782 DyDCNv2: modulated deformable 3x3 conv (offsets / mask logits in `om`) + GroupNorm(16).  Training path: the
    sampled, mask-weighted column tensor is materialised once; conv, dgrad and wgrad are then 1x1 GEMMs on it."""
    tp = g.tp
    n, h, w, c = x.n, x.h, x.w, x.c
    wkey = p + ".conv.weight"
    W = tp.conv(wkey, "fwd")      # [co][9][ci] == [co][1][9*ci]
    Wd = tp.conv(wkey, "col_dgrad")  # [9*ci][1][co]
    col = ops.deform_col(x, om, g.act(n, h, w, 9 * c))
    stats = g.f64(n, 16, 2)
    t = g.act(n, h, w, W.cout)
    ops.conv2d(col, W.w, t, impl=g.conv_impl, gn_stats=stats, gn_groups=16)

    def bwd():
        dt = g.gin(t)
        ops.conv_wgrad(col, dt, W.gw, impl=g.conv_impl)
        dcol = ops.conv2d(dt, Wd.w, g.act(n, h, w, 9 * c), impl=g.conv_impl)
        dxf = g.f32(n, h, w, c)
        g.mark(om)
        ops.deform_col_bwd(x, om, dcol, dxf, g.grad(om))
        ops.cast_acc(dxf, g.grad(x), g.mark(x))

    g.tape.append(bwd)
    return norm_act(g, t, p + ".norm.weight", p + ".norm.bias", 16, GN_EPS, ACT_NONE, stats=stats)


def ayhead_level(g, p, x, i):
    """nn/modules/head.py:1131-1175: one pyramid level of AYHead1.forward in train() -> raw (n, h, w, 4*reg_max + nc)"""
    n, h, w = x.n, x.h, x.w
    ad = conv_gn_act(g, f"{p}.stems.{i}", x)
    feat = conv_gn_act(g, p + ".share_conv.1", conv_gn_act(g, p + ".share_conv.0", ad))
    fc = feat.c
    avg = ops.gap(feat, g.f32(n, fc))
    davg_c, davg_r = g.f32(n, fc), g.f32(n, fc)

    def bwd_gap():
        df = g.grad(feat)
        ops.bcast_add(df, g.mark(feat), img=davg_c, s_img=1.0 / (h * w))
        ops.bcast_add(df, 1, img=davg_r, s_img=1.0 / (h * w))

    g.tape.append(bwd_gap)
    crc, crr = g.act(n, h, w, 2 * fc), g.act(n, h, w, 2 * fc)  # [cls | reg_to_cls(reg)] and [reg | cls_to_reg(cls)]
    cls = task_decomposition(g, p + ".cls_decomp", feat, avg, davg_c, crc.slice(0, fc))
    reg = task_decomposition(g, p + ".reg_decomp", feat, avg, davg_r, crr.slice(0, fc))
    # CrossTaskInteraction head.py:1319-1333
    q = p + ".cross_task"
    c2r = conv(g, cls, q + ".cls_to_reg.weight", q + ".cls_to_reg.bias", out=crr.slice(fc, fc))
    r2c = conv(g, reg, q + ".reg_to_cls.weight", q + ".reg_to_cls.bias", out=crc.slice(fc, fc))
    cg = conv(g, crc, q + ".cls_gate.0.weight", q + ".cls_gate.0.bias", act=ACT_SIGMOID)
    rg = conv(g, crr, q + ".reg_gate.0.weight", q + ".reg_gate.0.bias", act=ACT_SIGMOID)
    cls2 = mul_add(g, r2c, cg, c3=cls)
    reg2 = mul_add(g, c2r, rg, c3=reg)
    # ResidualBlockGN head.py:1031-1047
    cls_e = conv_gn_act(g, p + ".rep_block_cls.conv2", conv_gn_act(g, p + ".rep_block_cls.conv1", cls2), add=cls2)
    # DyDCNv2 (offsets / mask from `feat`, head.py:1155-1159) + CoordAtt
    om = conv(g, feat, p + ".spatial_conv_offset.weight", p + ".spatial_conv_offset.bias")
    ra = deform_conv_gn(g, p + ".DyDCNV2", reg2, om)
    reg_e = coord_att(g, p + ".coord_attention_reg", ra)
    # cls_prob head.py:1168-1169
    cp = conv(g, feat, p + ".cls_prob_conv.0.weight", p + ".cls_prob_conv.0.bias", act=ACT_RELU)
    cp = conv(g, cp, p + ".cls_prob_conv.2.weight", p + ".cls_prob_conv.2.bias", act=ACT_SIGMOID)  # channel 0 of 8
    z = g.like(cls_e)
    ops.eltwise_dev(2, cls_e, cp, z)

    def bwd_gate():
        dz = g.gin(z)
        dc = g.grad(cls_e)
        acc = g.mark(cls_e)
        ops.eltwise_dev(8 if acc else 2, dz, cp, dc, c3=dc if acc else None)
        g.mark(cp)
        ops.dot_pixel(dz, cls_e, g.grad(cp))

    g.tape.append(bwd_gate)
    reg_ch, nc = g.tp.shape[p + ".cv2.weight"][0], g.tp.shape[p + ".cv3.weight"][0]
    ncp = (nc + 7) // 8 * 8  # class counts that are not multiples of 8 (custom datasets): cv3's packed weight has pad8(nc) rows, the padding channels
    out = g.act(n, h, w, reg_ch + ncp)  # stay zero, yad_head_pack / yad_head_unpack hand the loss the first nc (as functional.ayhead_level does)
    r = conv(g, reg_e, p + ".cv2.weight", p + ".cv2.bias")
    scale_add(g, r, f"{p}.scale.{i}.scale", out=out.slice(0, reg_ch))
    conv(g, z, p + ".cv3.weight", p + ".cv3.bias", out=out.slice(reg_ch, ncp))
    return out


# ---- whole model ---------------------------------------------------------------------------------------------------------------
def forward_model(g, img):
    """train()-mode forward of z-yaml/yolo11-701-YOLO-AD-Refine.yaml (nn/tasks.py:141-168); img fp32 (n, 3, H, W) in [0, 1] or uint8 in [0, 255],
    on the device.
    Returns the three raw head outputs (head.py:1178-1179) as NHWC views (n, h, w, 4*reg_max + nc)."""
    n, _, H, W = img.shape
    assert H % 32 == 0 and W % 32 == 0, "image size must be a multiple of the maximum stride 32"
    L = {}
    # uint8 batches carry the trainer's device-side `.float() / 255` (models/yolo/detect/train.py:57-59 preprocess_batch) fused into the layout pass
    x0 = ops.u8_to_nhwc(img, g.act(n, H, W, 8)) if img.dtype == torch.uint8 else ops.nchw_to_nhwc(img, g.act(n, H, W, 8))
    L[0] = conv_bn_act(g, "model.0", x0, 2, need_dx=False)
    L[1] = conv_bn_act(g, "model.1", L[0], 2)
    L[2] = c3k2(g, "model.2", L[1])
    L[3] = conv_bn_act(g, "model.3", L[2], 2)
    L[4] = c3k2(g, "model.4", L[3])
    L[5] = conv_bn_act(g, "model.5", L[4], 2)
    L[6] = c3k2(g, "model.6", L[5], True, True)
    L[7] = conv_bn_act(g, "model.7", L[6], 2)
    L[8] = c3k2(g, "model.8", L[7], True, True)
    L[9] = sppf(g, "model.9", L[8])
    L[10] = c2ptssa(g, "model.10", L[9])
    L[11] = ela_hsfpn(g, "model.11", L[10], True)
    L[12] = conv(g, L[11], "model.12.weight", "model.12.bias")
    L[13] = conv(g, L[12], "model.13.weight", "model.13.bias", mode=CONV_TRANSPOSED)
    L[14] = ela_hsfpn(g, "model.14", L[6], True)
    L[15] = conv(g, L[14], "model.15.weight", "model.15.bias")
    L[16] = ela_hsfpn(g, "model.16", L[13], False)
    L[18] = mul_add(g, L[15], L[16], c3=L[13])  # Multiply (17) + Add (18)
    L[19] = c3k2(g, "model.19", L[18], False, True)
    L[20] = conv(g, L[19], "model.20.weight", "model.20.bias", mode=CONV_TRANSPOSED)
    L[21] = ela_hsfpn(g, "model.21", L[4], True)
    L[22] = conv(g, L[21], "model.22.weight", "model.22.bias")
    L[23] = ela_hsfpn(g, "model.23", L[20], False)
    L[25] = mul_add(g, L[22], L[23], c3=L[20])  # Multiply (24) + Add (25)
    L[26] = c3k2(g, "model.26", L[25], False, True)
    L[27] = conv_bn_act(g, "model.27", L[26], 2)
    L[28] = fusion_bifpn(g, "model.28", [L[27], L[19]])
    L[29] = c3k2(g, "model.29", L[28])
    L[30] = conv_bn_act(g, "model.30", L[29], 2)
    L[31] = fusion_bifpn(g, "model.31", [L[30], L[12]])
    L[32] = c3k2(g, "model.32", L[31])
    outs = [ayhead_level(g, "model.33", x, i) for i, x in enumerate((L[26], L[29], L[32]))]
    return outs, L
