"""Thin Python wrappers over the C ABI (include/yad.h).  PyTorch is only plumbing here: it owns device memory and streams.

`Act` is an NHWC activation view: a torch buffer of shape (n, h, w, ld) plus a channel window [c0, c0+c) -- the `yad_tensor` of the
C side.  A channel window of a wider buffer is how concatenations are eliminated (C2f / C3 / SPPF concat, cross-task concat).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_HARDSWISH, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SILU, BF16, CONV_DEFORM, CONV_NORMAL,  # noqa: F401
                   CONV_TRANSPOSED, F32, YadConvDesc, YadEpilogue, YadTensor, check)

_DT = {torch.float32: F32, torch.bfloat16: BF16}

# number of libyad kernel launches issued through this module (bench.py reports it as gpu_launches)
LAUNCHES = 0
# per-launch multiplicity of the multi-kernel entry points
_MULTI = {"yad_nms": 6, "yad_tal_assign": 2, "yad_gn_stats": 1, "yad_mlca_att": 2, "yad_norm_bwd": 2, "yad_mha_bwd": 2, "yad_mlca_bwd": 4}


# when set to a dict, every entry point is bracketed by CUDA events on the launching stream: name -> [(start, end, meta), ...]
PROFILE = None


def _count(name):
    global LAUNCHES
    LAUNCHES += _MULTI.get(name, 1)


def _call(name, *args, meta=None):
    """count, optionally time with CUDA events on the current stream, call the C entry point and raise on a non-zero status"""
    _count(name)
    fn = getattr(lib(), name)
    if PROFILE is None:
        check(fn(*args), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(fn(*args), name)
    e.record()
    PROFILE.setdefault(name, []).append((s, e, meta))


def lib():
    return _lib.load()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t):
    return _DT[t.dtype if isinstance(t, torch.Tensor) else t]


class Act:
    """NHWC activation view (n, h, w, c) with pixel stride ld inside `buf` (n, h, w, ld)."""
    __slots__ = ("buf", "c0", "c", "_yt")

    def __init__(self, buf, c0=0, c=None):
        assert buf.dim() == 4 and buf.is_contiguous(), "Act buffers are contiguous (n, h, w, ld)"
        self.buf, self.c0 = buf, c0
        self.c = buf.shape[3] - c0 if c is None else c
        assert self.c0 % 8 == 0 and self.c % 8 == 0 and buf.shape[3] % 8 == 0, (self.c0, self.c, buf.shape)
        self._yt = None

    n = property(lambda s: s.buf.shape[0])
    h = property(lambda s: s.buf.shape[1])
    w = property(lambda s: s.buf.shape[2])
    ld = property(lambda s: s.buf.shape[3])
    dtype = property(lambda s: s.buf.dtype)
    device = property(lambda s: s.buf.device)

    @property
    def ptr(self):
        return self.buf.data_ptr() + self.c0 * self.buf.element_size()

    def yt(self):
        if self._yt is None:
            self._yt = YadTensor(self.ptr, self.n, self.h, self.w, self.c, self.ld)
        return C.byref(self._yt)

    def slice(self, c0, c):
        assert 0 <= c0 and c0 + c <= self.c
        return Act(self.buf, self.c0 + c0, c)

    def images(self, i0, cnt):
        """sub-batch view (images are contiguous)"""
        return Act(self.buf[i0:i0 + cnt], self.c0, self.c)

    def reshape(self, n, h, w):
        """same memory, different (n, h, w) factorisation of the pixel axis (only for full-width views)"""
        assert n * h * w == self.n * self.h * self.w
        return Act(self.buf.view(n, h, w, self.ld), self.c0, self.c)

    def torch(self):
        """(n, h, w, c) torch view"""
        return self.buf[..., self.c0:self.c0 + self.c]

    def nchw(self):
        """(n, c, h, w) torch view with channels-last strides (no copy)"""
        return self.torch().permute(0, 3, 1, 2)

    @staticmethod
    def empty(n, h, w, c, dtype, device, ld=None):
        return Act(torch.empty((n, h, w, ld or c), dtype=dtype, device=device), 0, c)

    @staticmethod
    def from_nchw(x, dtype=None):
        """copy an (n, c, h, w) torch tensor into a fresh NHWC buffer (channels padded to a multiple of 8 with zeros)"""
        n, c, h, w = x.shape
        cp = (c + 7) // 8 * 8
        buf = torch.zeros((n, h, w, cp), dtype=dtype or x.dtype, device=x.device)
        buf[..., :c] = x.permute(0, 2, 3, 1)
        return Act(buf, 0, cp)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _ap(a):
    return (None, 0) if a is None else (C.c_void_p(a.ptr), a.ld)


def conv2d(x, w, y, bias=None, kh=1, kw=1, stride=1, pad_h=0, pad_w=0, act=ACT_NONE, alpha=1.0, img_scale=None, pix_scale=None,
           mul=None, add=None, mode=CONV_NORMAL, offmask=None, impl=0, gn_stats=None, gn_groups=0, gate=None):
    """y = epilogue(conv(x, w)); w packed [cout][kh*kw][cin] in the activation dtype.  gate = (gh, gw): Act views (n, out_h, 1, c) / (n, out_w, 1, c)
    whose product multiplies the output (ELA_HSFPN flag=False + Multiply without the materialised gate map; 1x1 stride-1 convolutions)."""
    d = YadConvDesc(mode, kh, kw, stride, pad_h, pad_w, None if offmask is None else offmask.ptr, 0 if offmask is None else offmask.ld, impl)
    mp, mld = _ap(mul)
    adp, ald = _ap(add)
    psp, psld = _ap(pix_scale)
    e = YadEpilogue(None if bias is None else bias.data_ptr(), None if img_scale is None else img_scale.data_ptr(),
                    psp, psld, act, alpha, mp, mld, adp, ald, None if gn_stats is None else gn_stats.data_ptr(), gn_groups)
    if gate is not None:
        gh, gw = gate
        assert (gh.n, gh.h, gh.w, gh.c) == (y.n, y.h, 1, y.c) and (gw.n, gw.h, gw.w, gw.c) == (y.n, y.w, 1, y.c) and gh.ld == gw.ld, "conv2d: gate shapes"
        e.gate_h, e.gate_w, e.gate_ld = gh.ptr, gw.ptr, gh.ld
    flops = 2.0 * y.n * y.h * y.w * y.c * kh * kw * x.c if mode != CONV_TRANSPOSED else 2.0 * x.n * x.h * x.w * y.c * kh * kw * x.c
    meta = None
    if PROFILE is not None:
        esz = x.buf.element_size()
        byts = esz * (x.n * x.h * x.w * x.c + y.n * y.h * y.w * y.c * (1 + (mul is not None) + (add is not None)) + w.numel()
                      + (0 if gate is None else y.n * (y.h + y.w) * y.c))
        meta = dict(flops=flops, bytes=byts, shape=f"{x.c}->{y.c} k{kh}x{kw} s{stride} m{mode} in{x.h}x{x.w} out{y.h}x{y.w} n{x.n}")
    _call("yad_conv2d", x.yt(), _p(w), C.byref(d), C.byref(e), y.yt(), dt(x.dtype), stream_ptr(), meta=meta)
    return y


def dwconv(x, w, y, bias=None, scale=None, shift=None, k=3, act=ACT_NONE, gate_split=0, add=None):
    adp, ald = _ap(add)
    _call("yad_dwconv", x.yt(), _p(w), _p(bias), _p(scale), _p(shift), k, act, gate_split, adp, ald, y.yt(), dt(x.dtype), stream_ptr())
    return y


def group_norm(x, y, stats, groups, gamma, beta, eps=1e-5, act=ACT_NONE, add=None, stats_ready=False):
    """stats: double (n, groups, 2) scratch; stats_ready: already filled by the producing conv's fused epilogue"""
    if not stats_ready:
        _call("yad_gn_stats", x.yt(), groups, _p(stats), dt(x.dtype), stream_ptr(), meta=_m(x, 1) if PROFILE is not None else None)
    adp, ald = _ap(add)
    _call("yad_gn_apply", x.yt(), _p(stats), groups, _p(gamma), _p(beta), eps, act, adp, ald, y.yt(), dt(x.dtype), stream_ptr(),
          meta=_m(x, 2 + (add is not None)) if PROFILE is not None else None)
    return y


def ln_mix(x, weight, bias, gamma, gammax, eps, y):
    """y = LayerNorm_c(x) * gamma + x * gammax (Mona, nn/modules/mona.py:55)"""
    for t in (weight, bias, gamma, gammax):
        assert t.dtype == torch.float32 and t.is_cuda and t.numel() == x.c
    _call("yad_ln_mix", x.yt(), _p(weight), _p(bias), _p(gamma), _p(gammax), float(eps), y.yt(), dt(x.dtype), stream_ptr(),
          meta=_m(x, 2) if PROFILE is not None else None)
    return y


def attention_tssa(w, temp, heads, out):
    """AttentionTSSA between its Linear layers (nn/modules/block.py:1666-1682): w, out (n, h, w, c) token maps; temp fp32 [heads]"""
    assert temp.dtype == torch.float32 and temp.is_cuda and temp.numel() == heads
    _call("yad_attention_tssa", w.yt(), _p(temp), heads, out.yt(), dt(w.dtype), stream_ptr(), meta=_m(w, 5) if PROFILE is not None else None)
    return out


def sppf_pool(x, y1, y2, y3):
    _call("yad_sppf_pool", x.yt(), y1.yt(), y2.yt(), y3.yt(), dt(x.dtype), stream_ptr())


def gap(x, out):
    _call("yad_gap", x.yt(), _p(out), dt(x.dtype), stream_ptr())
    return out


def rowcol_mean(x, rows, cols):
    _call("yad_rowcol_mean", x.yt(), rows.yt(), cols.yt(), dt(x.dtype), stream_ptr())


def rowcol_gate(x, gh, gw, y):
    _call("yad_rowcol_gate", None if x is None else x.yt(), gh.yt(), gw.yt(), y.yt(), dt(y.dtype), stream_ptr())
    return y


def coordatt_mlp(rows, cols, w1, b1, wh, bh, ww, bw, gh, gw):
    """CoordAtt's conv1 (+ folded bn1) -> hardswish -> conv_h / conv_w -> sigmoid on the pooled rows / columns (head.py:694-703); fp32 weights"""
    for t in (w1, b1, wh, bh, ww, bw):
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
    mip = b1.numel()
    assert w1.numel() == mip * rows.c and wh.numel() == gh.c * mip and ww.numel() == gw.c * mip and bh.numel() == gh.c and bw.numel() == gw.c
    _call("yad_coordatt_mlp", rows.yt(), cols.yt(), _p(w1), _p(b1), mip, _p(wh), _p(bh), _p(ww), _p(bw), gh.yt(), gw.yt(), dt(rows.dtype), stream_ptr())
    return gh, gw


def pool_upsample(x, s, y):
    _call("yad_pool_upsample", x.yt(), s, y.yt(), dt(x.dtype), stream_ptr())
    return y


def mlca(x, y, w_global, w_local, ksize, local, att, local_size=5, local_weight=0.5, add=None):
    """y = x * MLCA_attention(x) (+ add); local/att: fp32 (n, ls*ls, c) scratch"""
    L = lib()
    _call("yad_mlca_pool", x.yt(), _p(local), local_size, dt(x.dtype), stream_ptr())
    scratch = torch.empty((x.n, x.c), dtype=torch.float32, device=att.device)
    _call("yad_mlca_att", _p(local), _p(w_global), _p(w_local), ksize, local_weight, x.n, x.c, local_size, _p(att), _p(scratch), stream_ptr())
    adp, ald = _ap(add)
    _call("yad_mlca_apply", x.yt(), _p(att), local_size, adp, ald, y.yt(), dt(x.dtype), stream_ptr())
    return y


def mlca_apply(x, att, local_size, y, add=None):
    """y = x * adaptive_avg_pool(att fp32 [n][local_size^2][c] -> (h, w)) (+ add); local_size 1 = one gate per (image, channel) (SEBlock)"""
    assert att.dtype == torch.float32 and att.numel() == x.n * local_size * local_size * x.c
    adp, ald = _ap(add)
    _call("yad_mlca_apply", x.yt(), _p(att), local_size, adp, ald, y.yt(), dt(x.dtype), stream_ptr())
    return y


def gate_mlp(g, w1, b1, w2, b2, out, kind):
    n, c = g.shape
    hidden, nout = w1.shape[0], w2.shape[0]
    _call("yad_gate_mlp", _p(g), _p(w1), _p(b1), _p(w2), _p(b2), n, c, hidden, nout, kind, _p(out), stream_ptr())
    return out


def adt_apply(x, imp, alphas, weight, bias, y):
    _call("yad_adt_apply", x.yt(), _p(imp), _p(alphas), _p(weight), _p(bias), y.yt(), dt(x.dtype), stream_ptr())
    return y


def eltwise(op, a, b, y, c3=None, d4=None, alpha=1.0, beta=1.0, gamma=1.0):
    bp, bld = _ap(b)
    cp, cld = _ap(c3)
    dp, dld = _ap(d4)
    _call("yad_eltwise", op, a.yt(), bp, bld, cp, cld, dp, dld, alpha, beta, gamma, y.yt(), dt(a.dtype), stream_ptr())
    return y


def tssa(qkv, temps, heads, out, tok_offset):
    _call("yad_tssa", qkv.yt(), _p(temps), heads, out.yt(), tok_offset, dt(qkv.dtype), stream_ptr())


def mha(qkv, heads, out):
    _call("yad_mha", qkv.yt(), heads, out.yt(), dt(qkv.dtype), stream_ptr())
    return out


def group_mean(x, s, y):
    _call("yad_group_mean", x.yt(), s, y.yt(), dt(x.dtype), stream_ptr())
    return y


def patch_filter(x, m, y, alpha=1.0, add=None):
    adp, ald = _ap(add)
    _call("yad_patch_filter", x.yt(), _p(m), alpha, adp, ald, y.yt(), dt(x.dtype), stream_ptr())
    return y


def u8_to_nhwc(img_u8, y, scale=1.0 / 255.0):
    """img_u8: uint8 (n, c, h, w) contiguous on the device -> y = img * scale, NHWC (engine/predictor.py:129-133: H2D of uint8, then /255)"""
    assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous()
    _call("yad_u8_to_nhwc", _p(img_u8), img_u8.shape[1], y.yt(), scale, dt(y.dtype), stream_ptr())
    return y


def stem_conv(img, wgt, bias, y, act=ACT_SILU):
    """img: (n, cin, h, w) contiguous uint8 or fp32 on the device; wgt fp32 (16, 9*cin); y: NHWC (n, h/2, w/2, 16)"""
    assert img.is_contiguous() and img.dtype in (torch.uint8, torch.float32)
    n, cin, h, w = img.shape
    _call("yad_stem_conv", _p(img), int(img.dtype == torch.uint8), n, h, w, cin, _p(wgt), _p(bias), act, y.yt(), dt(y.dtype), stream_ptr())
    return y


def nchw_to_nhwc(img, y):
    """img: fp32 (n, c, h, w) contiguous"""
    assert img.dtype == torch.float32 and img.is_contiguous()
    _call("yad_nchw_to_nhwc", _p(img), img.shape[1], y.yt(), dt(y.dtype), stream_ptr())
    return y


def decode(levels, strides, nc, reg_max, proj, y):
    """levels: list of Act (n, h, w, 4*reg_max+nc) raw head outputs, or list of (B, no, H, W) torch tensors.  y: fp32 (B, 4+nc, N)."""
    nl = len(levels)
    ptrs, sb, sc, sa, hs, ws = (C.c_void_p * nl)(), (C.c_int64 * nl)(), (C.c_int64 * nl)(), (C.c_int64 * nl)(), (C.c_int32 * nl)(), (C.c_int32 * nl)()
    st = (C.c_float * nl)(*[float(s) for s in strides])
    dtype = None
    for i, lv in enumerate(levels):
        if isinstance(lv, Act):
            ptrs[i], sb[i], sc[i], sa[i], hs[i], ws[i] = lv.ptr, lv.h * lv.w * lv.ld, 1, lv.ld, lv.h, lv.w
            batch, d = lv.n, lv.dtype
        else:  # (B, no, H, W) tensor with arbitrary strides, as long as (H, W) flatten to one anchor stride
            b_, no_, h_, w_ = lv.shape
            assert lv.stride(2) == w_ * lv.stride(3), "decode: level must have a uniform anchor stride"
            ptrs[i], sb[i], sc[i], sa[i], hs[i], ws[i] = lv.data_ptr(), lv.stride(0), lv.stride(1), lv.stride(3), h_, w_
            batch, d = b_, lv.dtype
        assert dtype in (None, d)
        dtype = d
    _call("yad_decode", ptrs, sb, sc, sa, hs, ws, st, nl, batch, nc, reg_max, _p(proj), _p(y), dt(dtype), stream_ptr())
    return y


def nms_workspace_bytes(batch, n_anchors, nc, multi_label, max_nms):
    return int(lib().yad_nms_workspace_bytes(batch, n_anchors, nc, int(multi_label), max_nms))


def nms(pred, conf_thres, iou_thres, classes_mask, agnostic, multi_label, max_det, max_nms, max_wh, out, out_idx, out_count, workspace):
    b, ch, n = pred.shape
    assert pred.dtype == torch.float32 and pred.is_contiguous()
    _call("yad_nms", _p(pred), b, ch - 4, n, conf_thres, iou_thres, _p(classes_mask), int(agnostic), int(multi_label), max_det, max_nms,
                        max_wh, _p(out), _p(out_idx), _p(out_count), _p(workspace), stream_ptr())


def letterbox(desc, batch, out, out_h, out_w, pad_value=114, swap_rb=True):
    """desc: uint8 device tensor holding `batch` yad_image_desc records; out: contiguous uint8 (>= batch, 3, out_h, out_w) device tensor"""
    assert desc.is_cuda and desc.dtype == torch.uint8 and desc.numel() >= batch * 48 and out.is_cuda and out.dtype == torch.uint8
    _call("yad_letterbox", _p(desc), batch, _p(out), out_h, out_w, pad_value, int(swap_rb), stream_ptr())


def scale_boxes(det, count, desc, row_ld=None):
    """det: fp32 (B, max_det, row) device tensor (rows x1, y1, x2, y2, ...), scaled and clipped in place for rows < count[b] (count None: all)"""
    assert det.is_cuda and det.dtype == torch.float32 and det.dim() == 3
    b, max_det, row = det.shape
    row_ld = row_ld or det.stride(1)
    assert (b <= 1 or det.stride(0) == max_det * row_ld) and det.stride(2) == 1
    assert desc.is_cuda and desc.numel() >= b * 48 and (count is None or (count.dtype == torch.int32 and count.numel() >= b))
    _call("yad_scale_boxes", _p(det), row_ld, _p(count), b, max_det, _p(desc), stream_ptr())


def val_labels(bboxes_xywhn, batch_idx, img_hw, desc, out):
    """bboxes_xywhn fp32 (m, 4), batch_idx int32 (m), desc: yad_image_desc table (uint8 device tensor) -> out fp32 (m, 4) native-image xyxy"""
    m = bboxes_xywhn.shape[0]
    assert bboxes_xywhn.is_cuda and bboxes_xywhn.dtype == torch.float32 and bboxes_xywhn.is_contiguous() and out.is_contiguous()
    assert batch_idx.is_cuda and batch_idx.dtype == torch.int32 and batch_idx.numel() == m and out.shape == (m, 4) and out.dtype == torch.float32
    _call("yad_val_labels", _p(bboxes_xywhn), _p(batch_idx), m, int(img_hw[0]), int(img_hw[1]), _p(desc), _p(out), stream_ptr())


def val_match(det, count, gt_xyxy, gt_cls, gt_offset, max_labels_per_image, iouv, correct, stat_conf=None, stat_cls=None):
    """det fp32 (B, max_det, >= 6), count int32 (B) or None, gt_* grouped by image through gt_offset int32 (B + 1), iouv fp32 (niou)
    -> correct uint8 (B, max_det, niou)"""
    b, max_det, _ = det.shape
    assert det.is_cuda and det.dtype == torch.float32 and det.stride(2) == 1 and (b <= 1 or det.stride(0) == max_det * det.stride(1))
    assert gt_offset.dtype == torch.int32 and gt_offset.numel() == b + 1 and iouv.dtype == torch.float32 and correct.dtype == torch.uint8
    assert correct.is_contiguous() and correct.shape == (b, max_det, iouv.numel())
    assert gt_xyxy.dtype == torch.float32 and gt_cls.dtype == torch.float32 and gt_xyxy.is_contiguous() and gt_cls.is_contiguous()
    assert count is None or (count.dtype == torch.int32 and count.numel() >= b)
    for t in (stat_conf, stat_cls):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == b * max_det)
    _call("yad_val_match", _p(det), det.stride(1), _p(count), b, max_det, _p(gt_xyxy), _p(gt_cls), _p(gt_offset), int(max_labels_per_image), _p(iouv),
          iouv.numel(), _p(correct), _p(stat_conf), _p(stat_cls), stream_ptr())


def val_ap_workspace_bytes(n, nc, niou):
    r = int(lib().yad_val_ap_workspace_bytes(n, nc, niou))
    if r < 0:
        raise RuntimeError("yad_val_ap_workspace_bytes failed")
    return r


def val_ap(tp, conf, pred_cls, target_cls, nc, eps, ap, p_curve, r_curve, f1_curve, nt, summary, f1_index, workspace):
    n, niou = tp.shape
    assert tp.dtype == torch.uint8 and tp.is_contiguous() and conf.dtype == torch.float32 and pred_cls.dtype == torch.float32
    assert conf.numel() == n and pred_cls.numel() == n and target_cls.dtype == torch.float32
    assert ap.dtype == torch.float64 and ap.shape == (nc, niou) and all(t.dtype == torch.float64 and t.shape == (nc, 1000) and t.is_contiguous()
                                                                       for t in (p_curve, r_curve, f1_curve))
    assert nt.dtype == torch.int32 and nt.numel() == nc and summary.dtype == torch.float64 and summary.shape == (nc, 5) and f1_index.dtype == torch.int32
    _call("yad_val_ap", _p(tp), _p(conf), _p(pred_cls), n, _p(target_cls), target_cls.numel(), nc, niou, float(eps), _p(ap), _p(p_curve), _p(r_curve),
          _p(f1_curve), _p(nt), _p(summary), _p(f1_index), _p(workspace), stream_ptr())


# ----------------------------------------------------------------------------------------------------------------------
# training path (include/yad.h, "Training path"): thin wrappers, same conventions as above
# ----------------------------------------------------------------------------------------------------------------------
def _fp(t):
    """device pointer of an fp32 tensor / tensor view (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def _m(a, passes=1):
    """profile metadata of a memory-bound op on view `a`: `passes` = algorithmic number of full-tensor reads + writes"""
    if PROFILE is None:
        return None
    return dict(flops=0.0, bytes=passes * a.n * a.h * a.w * a.c * a.buf.element_size(), shape=f"{a.n}x{a.h}x{a.w}x{a.c}")


def eltwise_dev(op, a, b, y, c3=None, d4=None, alpha=1.0, beta=1.0, gamma=1.0, pa=None, pb=None, pg=None):
    bp, bld = _ap(b)
    cp, cld = _ap(c3)
    dp, dld = _ap(d4)
    _call("yad_eltwise_dev", op, a.yt(), bp, bld, cp, cld, dp, dld, alpha, beta, gamma, _fp(pa), _fp(pb), _fp(pg), y.yt(), dt(a.dtype), stream_ptr(),
          meta=_m(a, 2 + (b is not None) + (c3 is not None)))
    return y


def conv_wgrad(x, dy, dw, kh=1, kw=1, stride=1, pad_h=0, pad_w=0, impl=0):
    d = YadConvDesc(CONV_NORMAL, kh, kw, stride, pad_h, pad_w, None, 0, impl)
    meta = None
    if PROFILE is not None:
        meta = dict(flops=2.0 * dy.n * dy.h * dy.w * dy.c * kh * kw * x.c, bytes=0, shape=f"wgrad {x.c}->{dy.c} k{kh}x{kw} s{stride} out{dy.h}x{dy.w} n{x.n}")
    _call("yad_conv_wgrad", x.yt(), dy.yt(), C.byref(d), _fp(dw), dt(x.dtype), stream_ptr(), meta=meta)


def dwconv_wgrad(x, dy, k, dw):
    _call("yad_dwconv_wgrad", x.yt(), dy.yt(), k, _fp(dw), dt(x.dtype), stream_ptr(), meta=_m(x, 2))


def colsum(a, out, b=None):
    bp, bld = _ap(b)
    _call("yad_colsum", a.yt(), bp, bld, _fp(out), dt(a.dtype), stream_ptr(), meta=_m(a, 1 + (b is not None)))


def dot(a, b, out, scale=1.0, per_image=False, img_div=None):
    bp, bld = _ap(b)
    _call("yad_dot", a.yt(), bp, bld, int(per_image), scale, _fp(img_div), _fp(out), dt(a.dtype), stream_ptr())


def dot_pixel(a, b, y):
    bp, bld = _ap(b)
    _call("yad_dot_pixel", a.yt(), bp, bld, y.yt(), dt(a.dtype), stream_ptr())


def norm_bwd(x, dy, stats, groups, gamma, beta, eps, act, sums, dgamma, dbeta, dx, acc):
    _call("yad_norm_bwd", x.yt(), dy.yt(), _fp(stats), groups, _fp(gamma), _fp(beta), eps, act, _fp(sums), _fp(dgamma), _fp(dbeta), dx.yt(), int(acc),
          dt(x.dtype), stream_ptr(), meta=_m(x, 3))


def bn_running_update(stats, c, count, momentum, rmean, rvar):
    _call("yad_bn_running_update", _fp(stats), c, float(count), momentum, _fp(rmean), _fp(rvar), stream_ptr())


def act_bwd(y, dy, act, dx, acc):
    _call("yad_act_bwd", y.yt(), dy.yt(), act, dx.yt(), int(acc), dt(y.dtype), stream_ptr())


def bcast_add(dx, acc, img=None, s_img=0.0, row=None, s_row=0.0, col=None, s_col=0.0):
    _call("yad_bcast_add", dx.yt(), _fp(img), s_img, None if row is None else row.yt(), s_row, None if col is None else col.yt(), s_col, int(acc),
          dt(dx.dtype), stream_ptr())


def rowcol_gate_bwd(x, gh, gw, dy, dx, acc, dgh, dgw):
    _call("yad_rowcol_gate_bwd", None if x is None else x.yt(), gh.yt(), gw.yt(), dy.yt(), None if dx is None else dx.yt(), int(acc), dgh.yt(),
          dgw.yt(), dt(dy.dtype), stream_ptr())


def mlca_bwd(x, dy, local, att, wg, wl, ksize, datt, dlocal, dG, dwg, dwl, dx, acc, local_size=5, local_weight=0.5):
    _call("yad_mlca_bwd", x.yt(), dy.yt(), _fp(local), _fp(att), _fp(wg), _fp(wl), ksize, local_weight, local_size, _fp(datt), _fp(dlocal), _fp(dG),
          _fp(dwg), _fp(dwl), dx.yt(), int(acc), dt(x.dtype), stream_ptr())


def maxpool5_bwd(x, dy_t, dy_f, dx_f):
    p, ld = _ap(dy_t)
    _call("yad_maxpool5_bwd", x.yt(), p, ld, _fp(dy_f), _fp(dx_f), dt(x.dtype), stream_ptr())


def cast_acc(src, y, acc, scale=1.0):
    _call("yad_cast_acc", _fp(src), scale, y.yt(), int(acc), dt(y.dtype), stream_ptr())


def pool_upsample_bwd(dy, s, dpool, dx, acc):
    _call("yad_pool_upsample_bwd", dy.yt(), s, _fp(dpool), dx.yt(), int(acc), dt(dy.dtype), stream_ptr())


def gate_mlp_bwd(g, w1, b1, w2, b2, kind, dout, dg, dw1, db1, dw2, db2):
    n, c = g.shape
    hidden, nout = w1.shape[0], w2.shape[0]
    _call("yad_gate_mlp_bwd", _fp(g), _fp(w1), _fp(b1), _fp(w2), _fp(b2), n, c, hidden, nout, kind, _fp(dout), _fp(dg), _fp(dw1), _fp(db1), _fp(dw2),
          _fp(db2), stream_ptr())


def adt_bwd(x, dy, imp, alphas, weight, dx, acc, dimp, dalpha, dweight, dbias):
    _call("yad_adt_bwd", x.yt(), dy.yt(), _fp(imp), _fp(alphas), _fp(weight), dx.yt(), int(acc), _fp(dimp), _fp(dalpha), _fp(dweight), _fp(dbias),
          dt(x.dtype), stream_ptr())


def gelu_gate_bwd(a, b, dy, da, db, acc):
    _call("yad_gelu_gate_bwd", a.yt(), C.c_void_p(b.ptr), b.ld, dy.yt(), da.yt(), C.c_void_p(db.ptr), db.ld, int(acc), dt(a.dtype), stream_ptr())


def scale_img(a, s, y, acc=0):
    _call("yad_scale_img", a.yt(), _fp(s), y.yt(), int(acc), dt(a.dtype), stream_ptr())
    return y


def group_mean_bwd(dy, s, dx, acc=0):
    _call("yad_group_mean_bwd", dy.yt(), s, dx.yt(), int(acc), dt(dy.dtype), stream_ptr())


def patch_filter_bwd(x, dy, m, alpha, dx_f, dm):
    _call("yad_patch_filter_bwd", x.yt(), dy.yt(), _fp(m), alpha, _fp(dx_f), _fp(dm), dt(x.dtype), stream_ptr())


def tssa_bwd(qkv, temps, heads, dout, tok_offset, dqkv, dtemps):
    _call("yad_tssa_bwd", qkv.yt(), _fp(temps), heads, dout.yt(), tok_offset, dqkv.yt(), _fp(dtemps), dt(qkv.dtype), stream_ptr())


def mha_bwd(qkv, heads, out, dout, dqkv, lse_d):
    _call("yad_mha_bwd", qkv.yt(), heads, out.yt(), dout.yt(), dqkv.yt(), _fp(lse_d), dt(qkv.dtype), stream_ptr())


def deform_col(x, offmask, col):
    _call("yad_deform_col", x.yt(), offmask.yt(), col.yt(), dt(x.dtype), stream_ptr())
    return col


def deform_col_bwd(x, offmask, dcol, dx_f, doffmask):
    _call("yad_deform_col_bwd", x.yt(), offmask.yt(), dcol.yt(), _fp(dx_f), doffmask.yt(), dt(x.dtype), stream_ptr())


def head_pack(level, anchor0, n_anchors, reg_ch, nc, distri, logits):
    _call("yad_head_pack", level.yt(), anchor0, n_anchors, reg_ch, nc, _fp(distri), _fp(logits), dt(level.dtype), stream_ptr())


def head_unpack(gd, gl, scale, anchor0, n_anchors, reg_ch, nc, level):
    _call("yad_head_unpack", _fp(gd), _fp(gl), scale, anchor0, n_anchors, reg_ch, nc, level.yt(), dt(level.dtype), stream_ptr())


def fusion_weights(p, k, w=None, dw=None, dp=None):
    _call("yad_fusion_weights", _fp(p), k, _fp(w), _fp(dw), _fp(dp), stream_ptr())


def small_gemm(a, b, c, m, n, k, trans_a=False, acc=False):
    _call("yad_small_gemm", _fp(a), _fp(b), _fp(c), m, n, k, int(trans_a), int(acc), stream_ptr())
