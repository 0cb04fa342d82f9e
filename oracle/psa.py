"""CPU ORACLE (test infrastructure, NOT product code) -- the stock position-sensitive attention block of the sibling yamls
restated in plain fp32 PyTorch from a state dict: Attention.forward (nn/modules/block.py:906-925), PSABlock.forward (:960-964) and C2PSA.forward
(:1045-1049).  Pinned against the live reference by oracle/gen_golden.py psa_block (tests/golden/psa_block.npz).  Also C2SFA (:2358-2373) with its ProgressiveTSSA_Fusion0 / SimpleFeatureProcessor / SEBlock / StandardFFN
(:2049-2202), pinned by oracle/gen_golden.py sfa_block (tests/golden/sfa_block.npz).  Only tests/ may import this file."""
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


# ---- C2SFA (nn/modules/block.py:2358-2373): the C2PSA shell around ProgressiveTSSA_Fusion0 blocks ------------------------------------------------
# name -> (c1 = c2, n blocks, batch, h, w, seed)
SFA_CASES = {
    "c256_n1_20": (256, 1, 2, 20, 20, 95),      # layer 10 at 640^2, scale n
    "c256_n2_ragged": (256, 2, 1, 13, 11, 96),  # two stacked blocks
    "c128_n1_16": (128, 1, 1, 16, 16, 97),      # hidden width 64: two GroupNorm groups, SE bottleneck of 4 channels
}


def simple_feature_processor(sd, p, x):
    """nn/modules/block.py:2091-2096"""
    c = x.shape[1]
    t = F.group_norm(x, max(1, c // 32), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    t = F.gelu(F.conv2d(t, sd[p + ".conv_dw.weight"], sd[p + ".conv_dw.bias"], padding=1, groups=c))
    return F.conv2d(t, sd[p + ".conv_pw.weight"], sd[p + ".conv_pw.bias"])


def progressive_tssa_fusion0(sd, p, x):
    """nn/modules/block.py:2174-2202 (shortcut=True); SEBlock :2063-2064, StandardFFN :2077-2078"""
    t = simple_feature_processor(sd, p + ".pre_attn_block", x)
    gate = torch.sigmoid(F.conv2d(F.relu(F.conv2d(t.mean((2, 3), keepdim=True), sd[p + ".attn.fc.0.weight"])), sd[p + ".attn.fc.2.weight"]))
    x = x + t * gate * sd[p + ".residual_weight1"]
    u = simple_feature_processor(sd, p + ".pre_ffn_block", x)
    return x + F.conv2d(F.gelu(F.conv2d(u, sd[p + ".ffn.cv1.weight"])), sd[p + ".ffn.cv2.weight"]) * sd[p + ".residual_weight2"]


def c2sfa(sd, x, n, p="m"):
    """C2PSA.forward (nn/modules/block.py:1045-1049) with the ProgressiveTSSA_Fusion0 stack"""
    ab = conv_bn_act(sd, p + ".cv1", x)
    c = ab.shape[1] // 2
    a, b = ab[:, :c], ab[:, c:]
    for i in range(n):
        b = progressive_tssa_fusion0(sd, f"{p}.m.{i}", b)
    return conv_bn_act(sd, p + ".cv2", torch.cat((a, b), 1))


def sfa_state(shapes, seed):
    """oracle.mona.make_block_state with the two residual weights lifted off their 0.1 init (+-0.6 / -0.8: residual_weight1 is negative for odd seeds, which the
    scale folded into the SE bottleneck must survive)"""
    from oracle.mona import make_block_state
    sd = make_block_state(shapes, seed)
    for k in sd:
        if k.endswith("residual_weight1"):
            sd[k] = torch.tensor(-0.6 if seed % 2 else 0.6)
        elif k.endswith("residual_weight2"):
            sd[k] = torch.tensor(-0.8)
    return sd
