"""CPU ORACLE (test infrastructure, NOT product code) -- forward pass of the YOLO-AD-Refine model
(z-yaml/yolo11-701-YOLO-AD-Refine.yaml at scale `n`) restated as plain fp32 torch-CPU functions over a
state dict.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product path (yolo_ad_refine_b200/) never does.

Parity status: PINNED against the live reference (ultralytics fork under /root/reference, imported with
oracle/ref_shims.py) by oracle/gen_golden.py, which also writes the committed fixtures in tests/golden/.
The reference's own tests hold no golden vectors for this path; mmcv's
ModulatedDeformConv2d is absent and un-pinned, so DCN parity is pinned only to torchvision.ops.deform_conv2d
semantics (restated by hand in `deform_conv3x3`).

Every function cites the reference file:line it follows (paths relative to /root/reference/ultralytics).
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # utils/torch_utils.py:426-436 (initialize_weights sets eps on every BatchNorm2d)
GN_EPS = 1e-5  # torch default, nn/modules/head.py:1276


# --------------------------------------------------------------------------------------------------
# leaf ops
# --------------------------------------------------------------------------------------------------
def fold_bn(w, bn_w, bn_b, bn_m, bn_v, conv_b=None, eps=BN_EPS):
    """utils/torch_utils.py:243-270 fuse_conv_and_bn."""
    scale = bn_w / torch.sqrt(bn_v + eps)
    wf = w * scale.view(-1, *([1] * (w.dim() - 1)))
    b0 = torch.zeros_like(bn_m) if conv_b is None else conv_b
    bf = (b0 - bn_m) * scale + bn_b
    return wf, bf


def conv_bn_act(sd, p, x, act=True):
    """nn/modules/conv.py:36-54 Conv (conv bias=False -> BN -> SiLU); kernel/stride read off the weight."""
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    s = sd.get(p + ".__stride__", 1)
    if sd.get("__bn_train__"):  # nn.Module.train(): batch statistics (Conv.forward, conv.py:50)
        y = batch_norm(sd, p + ".bn", F.conv2d(x, w, None, stride=s, padding=k // 2))
        return F.silu(y) if act else y
    wf, bf = fold_bn(w, sd[p + ".bn.weight"], sd[p + ".bn.bias"], sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"])
    y = F.conv2d(x, wf, bf, stride=s, padding=k // 2)
    return F.silu(y) if act else y


BN_MOMENTUM = 0.03  # utils/torch_utils.py:426-436


def batch_norm(sd, p, x):
    """nn.BatchNorm2d (eps 1e-3, momentum 0.03): running statistics in eval, batch statistics when sd['__bn_train__'] is set; the updated
    running statistics are recorded in sd['__bn_updates__'] (a dict) when present."""
    if not sd.get("__bn_train__"):
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)
    upd = sd.get("__bn_updates__") if isinstance(sd.get("__bn_updates__"), dict) else {}
    # a module applied several times per step (the head's shared CoordAtt.bn1, once per pyramid level) updates its buffers each time
    rm = upd.get(p + ".running_mean", sd[p + ".running_mean"]).detach().clone()
    rv = upd.get(p + ".running_var", sd[p + ".running_var"]).detach().clone()
    y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], True, BN_MOMENTUM, BN_EPS)
    if isinstance(sd.get("__bn_updates__"), dict):
        sd["__bn_updates__"][p + ".running_mean"], sd["__bn_updates__"][p + ".running_var"] = rm, rv
    return y


def group_norm(x, w, b, groups, eps=GN_EPS):
    return F.group_norm(x, groups, w, b, eps)


def gn_groups(c2, num_groups=16):
    """nn/modules/head.py:1270-1274 (second, effective Conv_GN definition)."""
    g = min(num_groups, c2)
    if c2 % g != 0:
        g = max(i for i in range(1, g + 1) if c2 % i == 0)
    return g


def conv_gn_act(sd, p, x, act=True):
    """nn/modules/head.py:1265-1279 Conv_GN: conv(bias=False) -> GroupNorm -> SiLU."""
    w = sd[p + ".conv.weight"]
    y = F.conv2d(x, w, None, 1, w.shape[-1] // 2)
    y = group_norm(y, sd[p + ".gn.weight"], sd[p + ".gn.bias"], gn_groups(w.shape[0]))
    return F.silu(y) if act else y


def mlca(sd, p, x, local_size=5, local_weight=0.5):
    """nn/modules/block.py:1540-1584 MLCA."""
    b, c, m, n = x.shape
    local = F.adaptive_avg_pool2d(x, local_size)  # (b,c,5,5)
    glob = local.mean(dim=(2, 3))  # (b,c)
    w_g = sd[p + ".conv.weight"]
    w_l = sd[p + ".conv_local.weight"]
    k = w_g.shape[-1]
    # local: sequence index = (i*5+j)*c + ch  (block.py:1567)
    seq_l = local.reshape(b, c, -1).transpose(1, 2).reshape(b, 1, -1)
    y_l = F.conv1d(seq_l, w_l, padding=(k - 1) // 2)
    y_l = y_l.reshape(b, local_size * local_size, c).transpose(1, 2).reshape(b, c, local_size, local_size)
    y_g = F.conv1d(glob.reshape(b, 1, c), w_g, padding=(k - 1) // 2).reshape(b, c)
    att_l = y_l.sigmoid()
    # block.py:1575-1579: `y_global.view(b, -1).transpose(-1, -2).unsqueeze(-1)` is a 3-D (c, b, 1) tensor, so the adaptive pool that
    # follows treats the BATCH axis as the height: row i of the 5x5 grid = mean of sigmoid(y_g[b0:b1]) over the images
    # b0 = floor(i*b/5) .. b1 = ceil((i+1)*b/5), and the resulting (c, 5, 5) map is broadcast to every image.  For b == 1 this is the
    # plain per-image broadcast; for b > 1 the global branch mixes images -- reference behaviour, reproduced here.
    att_g = F.adaptive_avg_pool2d(y_g.t().unsqueeze(-1).sigmoid(), [local_size, local_size]).unsqueeze(0)  # (1,c,5,5)
    att = F.adaptive_avg_pool2d(att_g * (1 - local_weight) + att_l * local_weight, [m, n])
    return x * att


def bottleneck(sd, p, x, add=True, attention=False):
    """nn/modules/block.py:341-354 Bottleneck; :1586-1594 Bottleneck_MLCA."""
    y = conv_bn_act(sd, p + ".cv2", conv_bn_act(sd, p + ".cv1", x))
    if attention:
        y = mlca(sd, p + ".attention", y)
    return x + y if add else y


def c3k(sd, p, x, attention=False):
    """nn/modules/block.py:256-270 C3 / :742-750 C3k / :1596-1600 C3k_MLCA (n=2 bottlenecks)."""
    a = conv_bn_act(sd, p + ".cv1", x)
    for i in range(2):
        a = bottleneck(sd, f"{p}.m.{i}", a, add=True, attention=attention)
    return conv_bn_act(sd, p + ".cv3", torch.cat((a, conv_bn_act(sd, p + ".cv2", x)), 1))


def c3k2(sd, p, x, use_c3k=False, attention=False):
    """nn/modules/block.py:232-247 C2f.forward / :731-739 C3k2 / :1602-1605 C3k2_MLCA (n=1 after depth scaling)."""
    y = list(conv_bn_act(sd, p + ".cv1", x).chunk(2, 1))
    if use_c3k:
        y.append(c3k(sd, p + ".m.0", y[-1], attention))
    else:
        y.append(bottleneck(sd, p + ".m.0", y[-1], add=True, attention=attention))
    return conv_bn_act(sd, p + ".cv2", torch.cat(y, 1))


def sppf(sd, p, x):
    """nn/modules/block.py:177-196 SPPF."""
    y = [conv_bn_act(sd, p + ".cv1", x)]
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], 5, 1, 2))
    return conv_bn_act(sd, p + ".cv2", torch.cat(y, 1))


def ela_hsfpn(sd, p, x, flag=True):
    """nn/modules/block.py:1408-1424 ELA_HSFPN."""
    b, c, h, w = x.shape
    cw, cb = sd[p + ".conv1x1.0.weight"], sd[p + ".conv1x1.0.bias"]
    gw, gb = sd[p + ".conv1x1.1.weight"], sd[p + ".conv1x1.1.bias"]

    def branch(v):  # v: (b,c,L)
        return torch.sigmoid(F.group_norm(F.conv1d(v, cw, cb, padding=3), 16, gw, gb, GN_EPS))

    x_h = branch(x.mean(3)).reshape(b, c, h, 1)
    x_w = branch(x.mean(2)).reshape(b, c, 1, w)
    return x * x_h * x_w if flag else x_h * x_w


def fusion_bifpn(sd, p, xs):
    """nn/modules/block.py:1532-1535 Fusion('bifpn')."""
    w = F.relu(sd[p + ".fusion_weight"])
    w = w / (w.sum() + 1e-4)
    return sum(w[i] * xs[i] for i in range(len(xs)))


# --------------------------------------------------------------------------------------------------
# layer 10: C2ProgressiveTSSA_Fusion
# --------------------------------------------------------------------------------------------------
def progressive_feature_fusion(sd, p, x):
    """nn/modules/block.py:2579-2630 ProgressiveFeatureFusion."""
    outs, cur = [], x
    c = x.shape[1]
    for i in range(3):
        q = f"{p}.stages.{i}"
        t = F.conv2d(cur, sd[q + ".conv.weight"], sd[q + ".conv.bias"], 1, 1, groups=c)
        t = batch_norm(sd, q + ".norm", t)
        t = F.gelu(t)
        cm = F.conv2d(t, sd[q + ".channel_mix.weight"], sd[q + ".channel_mix.bias"])
        sm = F.conv2d(t, sd[q + ".spatial_mix.weight"], sd[q + ".spatial_mix.bias"], 1, 3, groups=c)
        out = cm + sm + cur
        outs.append(out)
        if i < 2:
            cur = F.conv2d(torch.cat([cur, out], 1), sd[f"{p}.stage_fusion.{i}.weight"], sd[f"{p}.stage_fusion.{i}.bias"])
    sa = sd[p + ".stage_attention"]
    return sum(sa[i] * outs[i] for i in range(3)) + x


def adaptive_dynamic_tanh(sd, p, x):
    """nn/modules/block.py:2493-2577 AdaptiveDynamicTanh (scale_weights is unused by forward, F10)."""
    g = x.mean(dim=(2, 3), keepdim=True)
    g = F.relu(F.conv2d(g, sd[p + ".importance_gate.1.weight"], sd[p + ".importance_gate.1.bias"]))
    imp = F.conv2d(g, sd[p + ".importance_gate.3.weight"], sd[p + ".importance_gate.3.bias"]).softmax(1)  # (b,3,1,1)
    al = sd[p + ".alphas"]
    y = sum(torch.tanh(al[:, i:i + 1] * x) * imp[:, i:i + 1] for i in range(3))
    return y * sd[p + ".weight"][:, None, None] + sd[p + ".bias"][:, None, None]


def multihead_self_attention(x, in_w, in_b, out_w, out_b, heads):
    """torch nn.MultiheadAttention(batch_first=True), eval, q=k=v=x (block.py:2432-2434, 2479-2486)."""
    B, T, C = x.shape
    d = C // heads
    qkv = x @ in_w.t() + in_b
    q, k, v = qkv.split(C, -1)
    q = q.view(B, T, heads, d).transpose(1, 2)
    k = k.view(B, T, heads, d).transpose(1, 2)
    v = v.view(B, T, heads, d).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d), -1)
    o = (att @ v).transpose(1, 2).reshape(B, T, C)
    return o @ out_w.t() + out_b


def cross_scale_attention_tssa(sd, p, x, heads=2, scales=(1, 2, 4)):
    """nn/modules/block.py:2417-2491 CrossScaleAttentionTSSA; returns (B, HW, C)."""
    B, C, H, W = x.shape
    d = C // heads
    feats = []
    for i, s in enumerate(scales):
        if s > 1:
            xs = F.adaptive_avg_pool2d(x, (H // s, W // s))
            xs = F.interpolate(xs, size=(H, W), mode="bilinear", align_corners=False)
        else:
            xs = x
        t = xs.flatten(2).permute(0, 2, 1)  # (B,HW,C)
        qkv = t @ sd[f"{p}.qkv_projections.{i}.weight"].t()
        q, k, v = qkv.chunk(3, -1)
        q = q.view(B, -1, heads, d).transpose(1, 2)  # b h n d
        k = k.view(B, -1, heads, d).transpose(1, 2)
        v = v.view(B, -1, heads, d).transpose(1, 2)
        wn = F.normalize(q, dim=-1)
        pi = torch.softmax((wn ** 2).sum(-1) * sd[p + ".temps"][i], dim=-1)  # (b,h,n)
        dots = pi.unsqueeze(-2) @ (k ** 2)  # (b,h,1,d)
        attn = 1.0 / (1 + dots)
        out = -(v * pi.unsqueeze(-1)) * attn
        feats.append(out.transpose(1, 2).reshape(B, -1, C))
    st = torch.stack(feats, 1).view(B, len(scales) * H * W, C)
    fused = multihead_self_attention(st, sd[p + ".cross_scale_fusion.in_proj_weight"], sd[p + ".cross_scale_fusion.in_proj_bias"],
                                     sd[p + ".cross_scale_fusion.out_proj.weight"], sd[p + ".cross_scale_fusion.out_proj.bias"], heads)
    fused = fused.view(B, len(scales), H * W, C).mean(1)
    return fused @ sd[p + ".to_out.0.weight"].t() + sd[p + ".to_out.0.bias"]


def edffn_spectral_matrix(fft_param):
    """The 8x8-patch rfft2 * W -> irfft2 of block.py:2405-2409 is linear per channel; return it as a
    (C, 64, 64) matrix M with out_patch.flatten() = M @ in_patch.flatten() (built with torch.fft itself)."""
    C = fft_param.shape[0]
    eye = torch.eye(64).view(64, 8, 8)
    f = torch.fft.rfft2(eye)  # (64,8,5)
    out = torch.fft.irfft2(f.unsqueeze(0) * fft_param.view(C, 1, 8, 5), s=(8, 8))  # (C,64in,8,8)
    return out.reshape(C, 64, 64).transpose(1, 2).contiguous()  # (C, out, in)


def edffn(sd, p, x):
    """nn/modules/block.py:2376-2415 EDFFN."""
    x = F.conv2d(x, sd[p + ".project_in.weight"])
    dw = sd[p + ".dwconv.weight"]
    x1, x2 = F.conv2d(x, dw, None, 1, 1, groups=dw.shape[0]).chunk(2, 1)
    x = F.conv2d(F.gelu(x1) * x2, sd[p + ".project_out.weight"])
    b, c, h, w = x.shape
    hn, wn = (8 - h % 8) % 8, (8 - w % 8) % 8
    x = F.pad(x, (0, wn, 0, hn), mode="reflect")
    H, W = x.shape[2:]
    xp = x.view(b, c, H // 8, 8, W // 8, 8).permute(0, 1, 2, 4, 3, 5)  # b c h w p1 p2
    f = torch.fft.rfft2(xp.float()) * sd[p + ".fft"]
    xp = torch.fft.irfft2(f, s=(8, 8))
    x = xp.permute(0, 1, 2, 4, 3, 5).reshape(b, c, H, W)
    return x[:, :, :h, :w]


def progressive_tssa_fusion(sd, p, x):
    """nn/modules/block.py:2632-2698 ProgressiveTSSA_Fusion.forward (shortcut=True)."""
    B, C, H, W = x.shape
    identity = x
    x = progressive_feature_fusion(sd, p + ".progressive_fusion1", x)
    x = adaptive_dynamic_tanh(sd, p + ".dyt1", x)
    a = cross_scale_attention_tssa(sd, p + ".attn", x).permute(0, 2, 1).reshape(B, C, H, W)
    x = identity + a * sd[p + ".residual_weight1"]
    x = progressive_feature_fusion(sd, p + ".progressive_fusion2", x)
    f = edffn(sd, p + ".ffn", adaptive_dynamic_tanh(sd, p + ".dyt2", x))
    return x + f * sd[p + ".residual_weight2"]


def c2ptssa(sd, p, x):
    """nn/modules/block.py:2700-2710 + C2PSA.forward :1045-1049."""
    a, b = conv_bn_act(sd, p + ".cv1", x).chunk(2, 1)
    b = progressive_tssa_fusion(sd, p + ".m.0", b)
    return conv_bn_act(sd, p + ".cv2", torch.cat((a, b), 1))


# --------------------------------------------------------------------------------------------------
# head
# --------------------------------------------------------------------------------------------------
def deform_conv3x3(x, offset, mask, weight):
    """Modulated deformable conv v2, 3x3, stride 1, pad 1, dilation 1, one offset group, no bias
    (nn/modules/head.py:772-779 -> mmcv ModulatedDeformConv2d; restated from torchvision deform_conv2d:
    offset channels = [dy, dx] interleaved per tap (tap = ki*3+kj), bilinear sampling with zero outside
    (-1, H) x (-1, W), mask per tap)."""
    B, C, H, W = x.shape
    ys = torch.arange(H, dtype=x.dtype).view(1, H, 1)
    xs = torch.arange(W, dtype=x.dtype).view(1, 1, W)
    cols = []
    for t in range(9):
        ki, kj = t // 3, t % 3
        py = ys + (ki - 1) + offset[:, 2 * t]
        px = xs + (kj - 1) + offset[:, 2 * t + 1]
        valid = (py > -1) & (px > -1) & (py < H) & (px < W)
        y0 = torch.floor(py)
        x0 = torch.floor(px)
        ly, lx = py - y0, px - x0
        y0, x0 = y0.long(), x0.long()
        val = torch.zeros(B, C, H, W, dtype=x.dtype)
        for dy, dx, wgt in ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx), (1, 0, ly * (1 - lx)), (1, 1, ly * lx)):
            yy, xx = y0 + dy, x0 + dx
            ok = valid & (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).view(B, 1, H * W).expand(B, C, H * W)
            g = x.reshape(B, C, H * W).gather(2, idx).view(B, C, H, W)
            val = val + g * (wgt * ok).unsqueeze(1)
        cols.append(val * mask[:, t].unsqueeze(1))
    col = torch.stack(cols, 2).reshape(B, C * 9, H * W)  # (B, C*9 [c-major, tap-minor], HW)
    return (weight.reshape(weight.shape[0], -1) @ col).view(B, -1, H, W)


def task_decomposition(sd, p, feat, avg):
    """nn/modules/head.py:626-669 TaskDecomposition with stacked_convs=1: per-image scalar gate on a 1x1 conv."""
    w = F.relu(F.conv2d(avg, sd[p + ".la_conv1.weight"], sd[p + ".la_conv1.bias"]))
    w = torch.sigmoid(F.conv2d(w, sd[p + ".la_conv2.weight"], sd[p + ".la_conv2.bias"]))  # (b,1,1,1)
    cw = sd[p + ".reduction_conv.conv.weight"]  # (64,64,1,1)
    b, c, h, ww = feat.shape
    conv_w = w.reshape(b, 1, 1, 1) * cw.reshape(1, cw.shape[0], 1, cw.shape[1])
    y = torch.bmm(conv_w.reshape(b, cw.shape[0], c), feat.reshape(b, c, h * ww)).reshape(b, cw.shape[0], h, ww)
    y = group_norm(y, sd[p + ".reduction_conv.gn.weight"], sd[p + ".reduction_conv.gn.bias"], gn_groups(cw.shape[0]))
    return F.silu(y)


def coord_att(sd, p, x):
    """nn/modules/head.py:671-707 CoordAtt."""
    n, c, h, w = x.shape
    x_h = x.mean(3, keepdim=True)  # (n,c,h,1)
    x_w = x.mean(2, keepdim=True).permute(0, 1, 3, 2)  # (n,c,w,1)
    y = torch.cat([x_h, x_w], 2)
    y = F.conv2d(y, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])
    y = batch_norm(sd, p + ".bn1", y)
    y = F.hardswish(y)
    y_h, y_w = torch.split(y, [h, w], 2)
    a_h = torch.sigmoid(F.conv2d(y_h, sd[p + ".conv_h.weight"], sd[p + ".conv_h.bias"]))
    a_w = torch.sigmoid(F.conv2d(y_w.permute(0, 1, 3, 2), sd[p + ".conv_w.weight"], sd[p + ".conv_w.bias"]))
    return x * a_w * a_h


def ayhead_level(sd, p, x, i):
    """nn/modules/head.py:1131-1175: one level of AYHead1.forward -> (B,144,H,W)."""
    ad = conv_gn_act(sd, f"{p}.stems.{i}", x)
    feat = conv_gn_act(sd, p + ".share_conv.1", conv_gn_act(sd, p + ".share_conv.0", ad))
    avg = feat.mean(dim=(2, 3), keepdim=True)
    cls = task_decomposition(sd, p + ".cls_decomp", feat, avg)
    reg = task_decomposition(sd, p + ".reg_decomp", feat, avg)
    # CrossTaskInteraction head.py:1319-1333
    q = p + ".cross_task"
    c2r = F.conv2d(cls, sd[q + ".cls_to_reg.weight"], sd[q + ".cls_to_reg.bias"])
    r2c = F.conv2d(reg, sd[q + ".reg_to_cls.weight"], sd[q + ".reg_to_cls.bias"])
    cg = torch.sigmoid(F.conv2d(torch.cat([cls, r2c], 1), sd[q + ".cls_gate.0.weight"], sd[q + ".cls_gate.0.bias"]))
    rg = torch.sigmoid(F.conv2d(torch.cat([reg, c2r], 1), sd[q + ".reg_gate.0.weight"], sd[q + ".reg_gate.0.bias"]))
    cls, reg = cls + r2c * cg, reg + c2r * rg
    # ResidualBlockGN head.py:1031-1047
    cls_e = conv_gn_act(sd, p + ".rep_block_cls.conv2", conv_gn_act(sd, p + ".rep_block_cls.conv1", cls)) + cls
    om = F.conv2d(feat, sd[p + ".spatial_conv_offset.weight"], sd[p + ".spatial_conv_offset.bias"], 1, 1)
    offset, mask = om[:, :18], om[:, 18:].sigmoid()
    reg_a = deform_conv3x3(reg, offset, mask, sd[p + ".DyDCNV2.conv.weight"])
    reg_a = group_norm(reg_a, sd[p + ".DyDCNV2.norm.weight"], sd[p + ".DyDCNV2.norm.bias"], 16)
    reg_e = coord_att(sd, p + ".coord_attention_reg", reg_a)
    cp = F.relu(F.conv2d(feat, sd[p + ".cls_prob_conv.0.weight"], sd[p + ".cls_prob_conv.0.bias"]))
    cp = torch.sigmoid(F.conv2d(cp, sd[p + ".cls_prob_conv.2.weight"], sd[p + ".cls_prob_conv.2.bias"], 1, 1))
    reg_out = F.conv2d(reg_e, sd[p + ".cv2.weight"], sd[p + ".cv2.bias"]) * sd[f"{p}.scale.{i}.scale"]
    cls_out = F.conv2d(cls_e * cp, sd[p + ".cv3.weight"], sd[p + ".cv3.bias"])
    return torch.cat((reg_out, cls_out), 1)


def make_anchors(shapes, strides, offset=0.5):
    """utils/tal.py:303-315 make_anchors: levels in order, row-major, (x+.5, y+.5)."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s)))
    return torch.cat(pts), torch.cat(st)


def decode(x_cat, shapes, strides=(8, 16, 32), reg_max=16, proj=None):
    """nn/modules/head.py:1181-1204,1236-1252 + block.py:78-81 DFL + utils/tal.py:318-327 dist2bbox(xywh).
    x_cat: (B, 4*reg_max+nc, N) -> y (B, 4+nc, N)."""
    B, no, N = x_cat.shape
    box, cls = x_cat[:, :4 * reg_max], x_cat[:, 4 * reg_max:]
    anchors, st = make_anchors(shapes, strides)
    prob = box.view(B, 4, reg_max, N).softmax(2)
    proj = torch.arange(reg_max, dtype=torch.float32) if proj is None else proj.reshape(-1).float()  # DFL conv weight
    dist = (prob * proj.view(1, 1, -1, 1)).sum(2)  # (B,4,N) ltrb
    a = anchors.t().unsqueeze(0)  # (1,2,N)
    lt, rb = dist[:, :2], dist[:, 2:]
    x1y1, x2y2 = a - lt, a + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st.t().unsqueeze(0)
    return torch.cat((dbox, cls.sigmoid()), 1)


def ayhead(sd, p, xs, training=False):
    """nn/modules/head.py:1127-1204 AYHead1.forward."""
    outs = [ayhead_level(sd, p, x, i) for i, x in enumerate(xs)]
    if training:
        return outs
    B = outs[0].shape[0]
    x_cat = torch.cat([o.reshape(B, o.shape[1], -1) for o in outs], 2)
    return decode(x_cat, [o.shape[2:] for o in outs], proj=sd.get(p + ".dfl.conv.weight")), outs


# --------------------------------------------------------------------------------------------------
# graph (z-yaml/yolo11-701-YOLO-AD-Refine.yaml resolved at scale n; nn/tasks.py:943-1108, :141-168)
# --------------------------------------------------------------------------------------------------
LAYERS = [  # (from, kind, args)
    (-1, "conv", 2), (-1, "conv", 2), (-1, "c3k2", (False, False)), (-1, "conv", 2), (-1, "c3k2", (False, False)),
    (-1, "conv", 2), (-1, "c3k2", (True, True)), (-1, "conv", 2), (-1, "c3k2", (True, True)), (-1, "sppf", None),
    (-1, "c2ptssa", None),
    (10, "ela", True), (-1, "conv2d", None), (12, "convT", None), (6, "ela", True), (-1, "conv2d", None),
    (13, "ela", False), ([15, 16], "mul", None), ([-1, 13], "add", None), (-1, "c3k2", (False, True)),
    (19, "convT", None), (4, "ela", True), (-1, "conv2d", None), (20, "ela", False), ([22, 23], "mul", None),
    ([-1, 20], "add", None), (-1, "c3k2", (False, True)),
    (26, "conv", 2), ([-1, 19], "fusion", None), (-1, "c3k2", (False, False)),
    (29, "conv", 2), ([-1, 12], "fusion", None), (-1, "c3k2", (False, False)),
    ([26, 29, 32], "head", None),
]


def run_layer(sd, i, kind, args, x, training=False):
    p = f"model.{i}"
    if kind == "conv":
        sd[p + ".__stride__"] = args
        return conv_bn_act(sd, p, x)
    if kind == "c3k2":
        return c3k2(sd, p, x, use_c3k=args[0], attention=args[1])
    if kind == "sppf":
        return sppf(sd, p, x)
    if kind == "c2ptssa":
        return c2ptssa(sd, p, x)
    if kind == "ela":
        return ela_hsfpn(sd, p, x, args)
    if kind == "conv2d":
        return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"])
    if kind == "convT":
        return F.conv_transpose2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=2, padding=1, output_padding=1)
    if kind == "mul":
        return x[0] * x[1]
    if kind == "add":
        return x[0] + x[1]
    if kind == "fusion":
        return fusion_bifpn(sd, p, x)
    if kind == "head":
        return ayhead(sd, p, x, training)
    raise ValueError(kind)


def forward(sd, img, training=False, return_layers=False, grad=False):
    """nn/tasks.py:141-168 BaseModel._predict_once over LAYERS. sd: fp32 CPU state dict; img: (B,3,H,W) 0-1.
    training: the head returns the raw level outputs (head.py:1178-1179); batch-statistics BatchNorm additionally needs
    sd['__bn_train__'] = True.  grad: run with autograd enabled (train_step_grads)."""
    sd = dict(sd)
    ys = []
    x = img
    with torch.set_grad_enabled(grad):
        for i, (f, kind, args) in enumerate(LAYERS):
            if f != -1:
                x = ys[f] if isinstance(f, int) else [x if j == -1 else ys[j] for j in f]
            x = run_layer(sd, i, kind, args, x, training)
            ys.append(x)
    return (x, ys) if return_layers else x


def train_step_grads(sd, img, batch_idx, cls, bboxes, gains=(7.5, 0.5, 1.5)):
    """One training forward + backward (nn/tasks.py:270-297 BaseModel.loss -> utils/loss.py:419-520; model.train(), so BatchNorm uses
    batch statistics).  Returns (loss*B, loss_items, {key: grad}, {bn buffer key: updated running stat}, raw head outputs)."""
    from .tal_loss import detection_loss
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()
              if isinstance(v, torch.Tensor) and v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))
              and not k.endswith("dfl.conv.weight")}
    work = dict(sd)
    work.update(params)
    work["__bn_train__"] = True
    work["__bn_updates__"] = {}
    feats = forward(work, img, training=True, grad=True)
    nc = int(sd["model.33.cv3.weight"].shape[0])  # the class count lives in the head's cv3 (head.py:1094-1097), 80 in the yaml
    loss, items, aux = detection_loss(feats, batch_idx, cls, bboxes, nc=nc, gains=gains)
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else None) for k, p in params.items()}
    return loss.detach(), items, grads, work["__bn_updates__"], [f.detach() for f in feats]
