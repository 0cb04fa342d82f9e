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


# ---- the block that carries the adapter in the sibling yamls: C2TSSA_DYT_Mona_EDFFN (nn/modules/block.py:1705-1709) ------------------------------
# name -> (c1 = c2, n blocks, batch, h, w, seed)
BLOCK_CASES = {
    "c256_n1_20": (256, 1, 2, 20, 20, 80),    # layer 10 at 640^2, scale n (depth 0.5 turns the yaml's 2 repeats into 1)
    "c256_n2_ragged": (256, 2, 1, 13, 11, 81),  # two stacked blocks, map sizes that are not multiples of the 8x8 EDFFN patch
    "c128_n1_16": (128, 1, 1, 16, 16, 82),    # one head of 64 channels: the softmax over heads degenerates to Pi = 1
}


def dynamic_tanh(sd, p, x):
    """nn/modules/block.py:1635-1641 (channels_last=False)"""
    return torch.tanh(sd[p + ".alpha"] * x) * sd[p + ".weight"][:, None, None] + sd[p + ".bias"][:, None, None]


def attention_tssa(sd, p, x, heads):
    """nn/modules/block.py:1665-1683 on tokens x (b, n, c)"""
    b, n, c = x.shape
    w = F.linear(x, sd[p + ".qkv.weight"]).reshape(b, n, heads, c // heads).permute(0, 2, 1, 3)      # b h n d
    w_sq = F.normalize(w, dim=-2) ** 2
    pi = torch.softmax(w_sq.sum(-1) * sd[p + ".temp"], dim=1)                                         # over the heads
    dots = torch.matmul((pi / (pi.sum(-1, keepdim=True) + 1e-8)).unsqueeze(-2), w ** 2)
    out = -(w * pi.unsqueeze(-1)) * (1.0 / (1 + dots))
    return F.linear(out.permute(0, 2, 1, 3).reshape(b, n, c), sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def tssa_dyt_mona_edffn(sd, p, x, heads):
    """nn/modules/block.py:1696-1703 (shortcut=True)"""
    from oracle.model import edffn
    b, c, h, w = x.shape
    t = dynamic_tanh(sd, p + ".dyt1", x).flatten(2).permute(0, 2, 1)
    x = x + attention_tssa(sd, p + ".attn", t, heads).permute(0, 2, 1).reshape(b, c, h, w)
    x = mona_forward(sd, x, p + ".mona1.")
    x = x + edffn(sd, p + ".ffn", dynamic_tanh(sd, p + ".dyt2", x))
    return mona_forward(sd, x, p + ".mona2.")


def c2tssa_dyt_mona_edffn(sd, x, n, p="m"):
    """C2PSA.forward (nn/modules/block.py:1045-1049) with the TSSAlock_DYT_Mona_EDFFN stack"""
    from oracle.model import conv_bn_act
    ab = conv_bn_act(sd, p + ".cv1", x)
    c = ab.shape[1] // 2
    a, b = ab[:, :c], ab[:, c:]
    for i in range(n):
        b = tssa_dyt_mona_edffn(sd, f"{p}.m.{i}", b, c // 64)
    return conv_bn_act(sd, p + ".cv2", torch.cat((a, b), 1))


def make_block_state(shapes, seed):
    """seeded state dict for a block whose {key: shape} table comes from the reference module (tests/golden/mona_block_spec.json, written by
    oracle/gen_golden.py mona_block): every parameter gets magnitude, BatchNorm buffers get non-trivial statistics, the Mona gammas are lifted off
    their 1e-6 init"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in shapes.items():
        shape = tuple(shape)
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(shape, dtype=torch.int64)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shape, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(shape, generator=g)
        elif k.endswith(".fft"):
            sd[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif k.endswith((".gamma", ".alpha", ".temp")):
            sd[k] = 0.5 + 0.2 * torch.rand(shape, generator=g)
        elif k.endswith(("bn.weight", "norm.weight", "gammax", "dyt1.weight", "dyt2.weight")):
            sd[k] = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[k] = torch.randn(shape, generator=g) / fan_in ** 0.5
        else:
            sd[k] = 0.1 * torch.randn(shape, generator=g)
    return sd
