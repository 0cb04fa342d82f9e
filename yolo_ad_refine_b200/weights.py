"""Host-side weight preparation: reference state-dict (key names of ultralytics' DetectionModel for
z-yaml/yolo11-701-YOLO-AD-Refine.yaml) -> kernel-layout device tensors.

 * Conv+BN folding follows utils/torch_utils.py:243-270 (fuse_conv_and_bn) with eps = 1e-3 (torch_utils.py:426-436).
 * conv weights are packed [cout][kh*kw][cin] (taps row-major) in the activation dtype; cin / cout are zero-padded to multiples
   of 8 (layer 0: cin 3 -> 8; offset/mask conv: 27 -> 32; cls_prob: 1 -> 8).
 * everything that is applied in fp32 inside a kernel (biases, GN/BN affine, depthwise taps, gates) stays fp32.
This is parameter plumbing done once per model load; it runs in PyTorch on the device.
"""
import math

import torch

BN_EPS = 1e-3
GN_EPS = 1e-5


def _pad8(n):
    return (n + 7) // 8 * 8


class ConvW:
    """packed dense-conv weights"""
    __slots__ = ("w", "b", "kh", "kw", "cin", "cout", "stride", "pad_h", "pad_w")

    def __init__(self, w, b, kh, kw, stride=1, pad_h=None, pad_w=None):
        self.w, self.b, self.kh, self.kw, self.stride = w, b, kh, kw, stride
        self.cout, self.cin = w.shape[0], w.shape[2]
        self.pad_h = kh // 2 if pad_h is None else pad_h
        self.pad_w = kw // 2 if pad_w is None else pad_w


def pack_conv(w, b, dtype, device, stride=1, transposed=False):
    """w: torch conv weight (cout, cin, kh, kw) [or (cin, cout, kh, kw) when transposed]; b: (cout,) or None."""
    w = w.detach().float()
    if transposed:
        w = w.permute(1, 0, 2, 3)
    cout, cin, kh, kw = w.shape
    cop, cip = _pad8(cout), _pad8(cin)
    p = torch.zeros(cop, kh * kw, cip, dtype=torch.float32)
    p[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    bias = None
    if b is not None:
        bias = torch.zeros(cop, dtype=torch.float32)
        bias[:cout] = b.detach().float()
        bias = bias.to(device)
    return ConvW(p.to(device=device, dtype=dtype).contiguous(), bias, kh, kw, stride)


def fold_bn(w, bn_w, bn_b, bn_m, bn_v, conv_b=None, eps=BN_EPS):
    scale = bn_w / torch.sqrt(bn_v + eps)
    wf = w * scale.view(-1, *([1] * (w.dim() - 1)))
    b0 = torch.zeros_like(bn_m) if conv_b is None else conv_b
    return wf, (b0 - bn_m) * scale + bn_b


def gn_groups(c2, num_groups=16):
    """nn/modules/head.py:1270-1274 (the second, effective Conv_GN definition)"""
    g = min(num_groups, c2)
    if c2 % g != 0:
        g = max(i for i in range(1, g + 1) if c2 % i == 0)
    return g


def edffn_spectral_matrix(fft_param):
    """rfft2 * W -> irfft2 on 8x8 patches (nn/modules/block.py:2405-2409) is linear per channel: returns M (64 out, 64 in, C) fp32
    with out_patch.flatten() = M[:, :, ch] @ in_patch.flatten(), built with torch.fft itself on the host."""
    fp = fft_param.detach().float().cpu()
    c = fp.shape[0]
    eye = torch.eye(64).view(64, 8, 8)
    f = torch.fft.rfft2(eye)
    out = torch.fft.irfft2(f.unsqueeze(0) * fp.view(c, 1, 8, 5), s=(8, 8))  # (C, in, 8, 8)
    return out.reshape(c, 64, 64).permute(2, 1, 0).contiguous()  # (out, in, C)


class Prepared:
    """Namespace of prepared parameters, addressed by reference module path (e.g. P['model.2.cv1'])."""

    def __init__(self, sd, dtype, device):
        self.sd = {k: v.detach().cpu() for k, v in sd.items()}  # host copy: scalars / folding are host-side, one-off
        self.dtype, self.device = dtype, device
        self.cache = {}

    def f32(self, key):
        k = ("f32", key)
        if k not in self.cache:
            self.cache[k] = self.sd[key].float().contiguous().to(self.device)
        return self.cache[k]

    def has(self, key):
        return key in self.sd

    def conv_bn(self, p, stride=1):
        """conv.Conv: conv(bias=False) + BN folded.  Handles both the unfused (.bn.*) and the fused (.conv.bias) state dict."""
        k = ("conv_bn", p)
        if k not in self.cache:
            w = self.sd[p + ".conv.weight"].float()
            if p + ".bn.weight" in self.sd:
                w, b = fold_bn(w, self.sd[p + ".bn.weight"].float(), self.sd[p + ".bn.bias"].float(), self.sd[p + ".bn.running_mean"].float(),
                               self.sd[p + ".bn.running_var"].float(), self.sd.get(p + ".conv.bias"))
            else:
                b = self.sd[p + ".conv.bias"].float()
            self.cache[k] = pack_conv(w, b, self.dtype, self.device, stride)
        return self.cache[k]

    def conv(self, wkey, bkey=None, stride=1, transposed=False, scale_rows=None):
        k = ("conv", wkey, bkey, transposed)
        if k not in self.cache:
            w = self.sd[wkey].float()
            if w.dim() == 2:
                w = w[:, :, None, None]
            elif w.dim() == 3:  # Conv1d (cout, cin, k) -> k x 1 kernel over the (n, L, 1, c) view
                w = w[:, :, :, None]
            b = self.sd[bkey].float() if bkey else None
            self.cache[k] = pack_conv(w, b, self.dtype, self.device, stride, transposed)
        return self.cache[k]

    def conv_raw(self, name, w, b, stride=1):
        k = ("raw", name)
        if k not in self.cache:
            self.cache[k] = pack_conv(w, b, self.dtype, self.device, stride)
        return self.cache[k]

    def dw(self, wkey, bkey=None):
        """depthwise weight (c, 1, k, k) -> fp32 [k*k][c]"""
        k = ("dw", wkey)
        if k not in self.cache:
            w = self.sd[wkey].float()
            c, _, kh, kw = w.shape
            self.cache[k] = (w.reshape(c, kh * kw).t().contiguous().to(self.device), self.f32(bkey) if bkey else None, kh)
        return self.cache[k]

    def bn_affine(self, p, eps=BN_EPS):
        k = ("bn", p)
        if k not in self.cache:
            scale = self.sd[p + ".weight"].float() / torch.sqrt(self.sd[p + ".running_var"].float() + eps)
            shift = self.sd[p + ".bias"].float() - self.sd[p + ".running_mean"].float() * scale
            self.cache[k] = (scale.contiguous().to(self.device), shift.contiguous().to(self.device))
        return self.cache[k]

    def scalar(self, key):
        return float(self.sd[key].float().reshape(-1)[0])

    def vec(self, key):
        return [float(v) for v in self.sd[key].float().reshape(-1)]

    def misc(self, name, fn):
        k = ("misc", name)
        if k not in self.cache:
            self.cache[k] = fn()
        return self.cache[k]


def prepare(sd, dtype=torch.bfloat16, device="cuda"):
    return Prepared(sd, dtype, torch.device(device))


__all__ = ["prepare", "Prepared", "ConvW", "pack_conv", "gn_groups", "edffn_spectral_matrix", "BN_EPS", "GN_EPS", "math"]
