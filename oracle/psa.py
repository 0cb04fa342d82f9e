"""CPU ORACLE (test infrastructure, NOT product code) -- the stock position-sensitive attention block of the sibling yamls
restated in plain fp32 PyTorch from a state dict: Attention.forward (nn/modules/block.py:906-925), PSABlock.forward (:960-964) and C2PSA.forward
(:1045-1049).  Pinned against the live reference by oracle/gen_golden.py psa_block (tests/golden/psa_block.npz).  Only tests/ may import this file."""
import torch
import torch.nn.functional as F

from oracle.model import conv_bn_act

# name -> (c1 = c2, n blocks, batch, h, w, seed)
PSA_CASES = {
    "c256_n1_20": (256, 1, 2, 20, 20, 90),     # layer 10 at 640^2, scale n: hidden width 128 = 2 heads of 64, 400 tokens
    "c256_n2_ragged": (256, 2, 1, 13, 11, 91),  # two stacked blocks, a token count (143) that is not a multiple of the attention tiles
    "c128_n1_16": (128, 1, 1, 16, 16, 92),     # one head
    "c512_n1_7": (512, 1, 1, 7, 7, 93),        # four heads, fewer tokens (49) than one key tile
}


def conv_bn(sd, p, x, groups=1):
    """conv.Conv with act=False: Conv2d(bias=False) + BatchNorm2d(eps 1e-3) in eval mode"""
    w = sd[p + ".conv.weight"].float()
    y = F.conv2d(x, w, None, 1, w.shape[-1] // 2, groups=groups)
    g = lambda k: sd[p + ".bn." + k].float()
    return F.batch_norm(y, g("running_mean"), g("running_var"), g("weight"), g("bias"), False, 0.0, 1e-3)


def attention(sd, p, x, heads):
    """nn/modules/block.py:906-925 Attention.forward (attn_ratio 0.5)"""
    b, c, h, w = x.shape
    n = h * w
    hd = c // heads
    kd = int(hd * 0.5)
    qkv = conv_bn(sd, p + ".qkv", x)
    q, k, v = qkv.view(b, heads, 2 * kd + hd, n).split([kd, kd, hd], dim=2)
    attn = ((q.transpose(-2, -1) @ k) * kd ** -0.5).softmax(dim=-1)
    y = (v @ attn.transpose(-2, -1)).view(b, c, h, w) + conv_bn(sd, p + ".pe", v.reshape(b, c, h, w), groups=c)
    return conv_bn(sd, p + ".proj", y)


def psablock(sd, p, x, heads):
    """nn/modules/block.py:960-964 (shortcut=True)"""
    x = x + attention(sd, p + ".attn", x, heads)
    return x + conv_bn(sd, p + ".ffn.1", conv_bn_act(sd, p + ".ffn.0", x))


def c2psa(sd, x, n, p="m"):
    """nn/modules/block.py:1045-1049 C2PSA.forward"""
    ab = conv_bn_act(sd, p + ".cv1", x)
    c = ab.shape[1] // 2
    a, b = ab[:, :c], ab[:, c:]
    for i in range(n):
        b = psablock(sd, f"{p}.m.{i}", b, c // 64)
    return conv_bn_act(sd, p + ".cv2", torch.cat((a, b), 1))
