"""CPU ORACLE (test infrastructure, NOT product code) -- the Mona adapter restated in plain fp32 PyTorch from a state
dict: Mona.forward (nn/modules/mona.py:52-64, eval mode: dropout = identity), LayerNorm2d (:5-10), MonaOp.forward (:21-34).  Pinned against the live
reference by oracle/gen_golden.py mona (tests/golden/mona.npz).  Only tests/ may import this file."""
import numpy as np
import torch
import torch.nn.functional as F

# name -> (channels, batch, h, w, seed)
MONA_CASES = {
    "c128_20": (128, 2, 20, 20, 70),     # layer-10 geometry of the sibling yamls at 640^2 (scale n: 128 channels at 20x20)
    "c64_ragged": (64, 1, 7, 9, 71),     # smaller than the 7x7 kernel in one direction: every tap class hits the zero padding
    "c256_12": (256, 1, 12, 12, 72),
}


def make_state(c, seed):
    """seeded Mona(c) parameters; every term of the forward carries weight (the reference's gamma init of 1e-6 would hide the LayerNorm branch)"""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    sd = {"project1.weight": r(64, c, 1, 1) / c ** 0.5, "project1.bias": 0.1 * r(64),
          "project2.weight": r(c, 64, 1, 1) / 8.0, "project2.bias": 0.1 * r(c),
          "adapter_conv.projector.weight": r(64, 64, 1, 1) / 8.0, "adapter_conv.projector.bias": 0.1 * r(64),
          "norm.weight": 1.0 + 0.2 * r(c), "norm.bias": 0.1 * r(c),
          "gamma": (0.5 + 0.2 * r(c)).reshape(c, 1, 1), "gammax": (1.0 + 0.2 * r(c)).reshape(c, 1, 1)}
    for i, k in ((1, 3), (2, 5), (3, 7)):
        sd[f"adapter_conv.conv{i}.weight"] = r(64, 1, k, k) / k
        sd[f"adapter_conv.conv{i}.bias"] = 0.1 * r(64)
    return sd


def make_input(c, n, h, w, seed):
    return torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(seed + 1000))


def mona_forward(sd, x, p=""):
    """x (n, c, h, w) fp32 -> (n, c, h, w)"""
    g = lambda k: sd[p + k].float()
    c = x.shape[1]
    ln = F.layer_norm(x.permute(0, 2, 3, 1), (c,), g("norm.weight"), g("norm.bias"), 1e-5).permute(0, 3, 1, 2)   # LayerNorm2d
    t = ln * g("gamma") + x * g("gammax")
    t = F.conv2d(t, g("project1.weight"), g("project1.bias"))
    dw = sum(F.conv2d(t, g(f"adapter_conv.conv{i}.weight"), g(f"adapter_conv.conv{i}.bias"), padding=k // 2, groups=t.shape[1])
             for i, k in ((1, 3), (2, 5), (3, 7)))
    s = dw / 3.0 + t
    s = s + F.conv2d(s, g("adapter_conv.projector.weight"), g("adapter_conv.projector.bias"))
    return x + F.conv2d(F.gelu(s), g("project2.weight"), g("project2.bias"))
