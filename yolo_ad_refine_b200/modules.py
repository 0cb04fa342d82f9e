"""nn.Module mirrors of the reference blocks named by z-yaml/yolo11-701-YOLO-AD-Refine.yaml -- the drop-in boundary of SURVEY.md section 8b.

Each class keeps the reference's constructor signature, attribute names and state-dict key names (so `parse_model` can instantiate it by
name and reference checkpoints load), but holds parameters only: `forward` hands NHWC views to libyad.so through functional.py.  There is no
PyTorch arithmetic on the forward path and no CPU fallback.  Paths in docstrings are relative to /root/reference/ultralytics.

Training: the reference calls `model(batch)` with the batch dict in training mode (engine/trainer.py:382-384); plugin.install() routes that call to
bridge.TrainBridge, which runs the libyad training graph (training.py) on these modules' parameters and returns a loss whose backward fills their
.grad.  A training-mode forward of a single module on a TENSOR has no libyad autograd node and raises.
"""
import math

import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .ops import Act
from .weights import gn_groups, prepare


def autopad(k, p=None, d=1):
    """nn/modules/conv.py:27-33"""
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (x - 1) + 1 for x in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


def as_act(x):
    """(n, c, h, w) tensor -> Act.  Zero-copy when x is already a channels-last bf16/fp32 view with c % 8 == 0."""
    if isinstance(x, Act):
        return x
    if not x.is_cuda:
        raise RuntimeError("yolo_ad_refine_b200 modules run on CUDA tensors only (no CPU fallback)")
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"yolo_ad_refine_b200 computes in bf16 or fp32; got {x.dtype} (use .bfloat16() instead of .half())")
    nhwc = x.permute(0, 2, 3, 1)
    if nhwc.is_contiguous() and x.shape[1] % 8 == 0:
        return Act(nhwc)
    return Act.from_nchw(x)


class YadModule(nn.Module):
    """Base: lazily prepares kernel-layout weights from the module's own state dict (prefix 'm.')."""

    def _ctx(self, x):
        # the prepared (BN-folded, packed) weights are keyed by dtype, device AND the in-place version counters of every parameter / buffer, so an
        # optimizer step, an EMA copy_ or model.fuse() can never leave inference running on stale weights
        key = (x.dtype, x.device, sum(t._version for t in self.parameters()) + sum(t._version for t in self.buffers()))
        if getattr(self, "_yad_key", None) != key:
            sd = {"m." + k: v for k, v in self.state_dict().items()}
            self._yad_ctx = Fn.Ctx(prepare(sd, x.dtype, x.device))
            self._yad_key = key
        return self._yad_ctx

    def refresh(self):
        """drop the prepared weights (call after loading / changing parameters)"""
        self._yad_key = None

    def _load_from_state_dict(self, *a, **k):
        self._yad_key = None
        return super()._load_from_state_dict(*a, **k)

    def _check_eval(self):
        if self.training:
            raise NotImplementedError("a training-mode forward of one libyad module on a tensor has no autograd node: train through model(batch) "
                                      "(plugin.install() binds DetectionModel.loss to bridge.TrainBridge) or TrainEngine; call .eval() for inference")


# ---------------------------------------------------------------------------------------------------------------------------
# parameter holders (no forward of their own)
# ---------------------------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter holder: the owning yolo_ad_refine_b200 module runs the fused forward")


def _conv_holder(c1, c2, k=1, s=1, p=None, g=1, d=1):
    h = _Holder()
    h.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
    h.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
    return h


class Conv(YadModule):
    """nn/modules/conv.py:36-54 Conv(c1, c2, k=1, s=1, p=None, g=1, d=1, act=True): Conv2d(bias=False) + BatchNorm2d + SiLU"""
    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        assert g == 1 and d == 1, "grouped / dilated Conv is outside the YOLO-AD-Refine path"
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()
        self._stride = s

    def forward(self, x):
        self._check_eval()
        a = as_act(x)
        ctx = self._ctx(a)
        act = ops.ACT_SILU if isinstance(self.act, nn.SiLU) else ops.ACT_NONE
        return Fn.conv(ctx, a, ctx.P.conv_bn("m", self._stride), act=act).nchw()

    forward_fuse = forward


class _Bottleneck(_Holder):
    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5, attention=False):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = _conv_holder(c1, c_, k[0], 1)
        self.cv2 = _conv_holder(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2
        if attention:
            self.attention = _MLCA(c2)


class _MLCA(_Holder):
    """nn/modules/block.py:1540-1557"""

    def __init__(self, in_size, local_size=5, gamma=2, b=1, local_weight=0.5):
        super().__init__()
        t = int(abs(math.log(in_size, 2) + b) / gamma)
        k = t if t % 2 else t + 1
        self.conv = nn.Conv1d(1, 1, kernel_size=k, padding=(k - 1) // 2, bias=False)
        self.conv_local = nn.Conv1d(1, 1, kernel_size=k, padding=(k - 1) // 2, bias=False)


class _C3k(_Holder):
    """nn/modules/block.py:256-270 C3 + :742-750 C3k / :1596-1600 C3k_MLCA"""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, k=3, attention=False):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = _conv_holder(c1, c_, 1, 1)
        self.cv2 = _conv_holder(c1, c_, 1, 1)
        self.cv3 = _conv_holder(2 * c_, c2, 1)
        self.m = nn.Sequential(*(_Bottleneck(c_, c_, shortcut, g, k=(k, k), e=1.0, attention=attention) for _ in range(n)))


class C3k2(YadModule):
    """nn/modules/block.py:731-739 C3k2(c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True) (C2f.forward :243-247)"""
    _attention = False

    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True):
        super().__init__()
        assert n == 1, "the YOLO-AD-Refine yaml resolves every C3k2 to n = 1 at scale n"
        self.c = int(c2 * e)
        self.cv1 = _conv_holder(c1, 2 * self.c, 1, 1)
        self.cv2 = _conv_holder((2 + n) * self.c, c2, 1)
        self.c3k = c3k
        self.m = nn.ModuleList(_C3k(self.c, self.c, 2, shortcut, g, attention=self._attention) if c3k
                               else _Bottleneck(self.c, self.c, shortcut, g, attention=self._attention) for _ in range(n))

    def forward(self, x):
        self._check_eval()
        a = as_act(x)
        return Fn.c3k2(self._ctx(a), "m", a, self.c3k, self._attention).nchw()


class C3k2_MLCA(C3k2):
    """nn/modules/block.py:1602-1605"""
    _attention = True


class SPPF(YadModule):
    """nn/modules/block.py:177-196 SPPF(c1, c2, k=5)"""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        assert k == 5
        c_ = c1 // 2
        self.cv1 = _conv_holder(c1, c_, 1, 1)
        self.cv2 = _conv_holder(c_ * 4, c2, 1, 1)

    def forward(self, x):
        self._check_eval()
        a = as_act(x)
        return Fn.sppf(self._ctx(a), "m", a).nchw()


class _PFF(_Holder):
    """nn/modules/block.py:2579-2604 ProgressiveFeatureFusion"""

    def __init__(self, c, num_stages=3):
        super().__init__()
        self.stage_attention = nn.Parameter(torch.ones(num_stages) / num_stages)
        self.stages = nn.ModuleList()
        for _ in range(num_stages):
            st = _Holder()
            st.conv = nn.Conv2d(c, c, 3, 1, 1, groups=c)
            st.norm = nn.BatchNorm2d(c, eps=1e-3, momentum=0.03)
            st.channel_mix = nn.Conv2d(c, c, 1)
            st.spatial_mix = nn.Conv2d(c, c, 7, 1, 3, groups=c)
            self.stages.append(st)
        self.stage_fusion = nn.ModuleList(nn.Conv2d(2 * c, c, 1) for _ in range(num_stages - 1))


class _ADT(_Holder):
    """nn/modules/block.py:2493-2540 AdaptiveDynamicTanh"""

    def __init__(self, c, num_scales=3):
        super().__init__()
        self.alphas = nn.Parameter(torch.linspace(0.3, 1.0, num_scales).view(1, num_scales, 1, 1))
        self.scale_weights = nn.Parameter(torch.ones(num_scales) / num_scales)
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.importance_gate = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(c, c // 4, 1), nn.ReLU(inplace=True), nn.Conv2d(c // 4, num_scales, 1),
                                             nn.Softmax(dim=1))


class _CSATSSA(_Holder):
    """nn/modules/block.py:2417-2442 CrossScaleAttentionTSSA"""

    def __init__(self, c, num_heads=2, scales=(1, 2, 4)):
        super().__init__()
        self.temps = nn.Parameter(torch.ones(len(scales), num_heads, 1))
        self.qkv_projections = nn.ModuleList(nn.Linear(c, c * 3, bias=False) for _ in scales)
        self.cross_scale_fusion = nn.MultiheadAttention(c, num_heads, batch_first=True)
        self.to_out = nn.Sequential(nn.Linear(c, c), nn.Dropout(0.0))


class _EDFFN(_Holder):
    """nn/modules/block.py:2376-2392 EDFFN(dim, ffn_expansion_factor=2, bias=False)"""

    def __init__(self, c, factor=2):
        super().__init__()
        hidden = int(c * factor)
        self.fft = nn.Parameter(torch.ones((c, 1, 1, 8, 8 // 2 + 1)))
        self.project_in = nn.Conv2d(c, hidden * 2, 1, bias=False)
        self.dwconv = nn.Conv2d(hidden * 2, hidden * 2, 3, 1, 1, groups=hidden * 2, bias=False)
        self.project_out = nn.Conv2d(hidden, c, 1, bias=False)


class _PTSSA(_Holder):
    """nn/modules/block.py:2632-2660 ProgressiveTSSA_Fusion"""

    def __init__(self, c, num_heads):
        super().__init__()
        self.progressive_fusion1, self.progressive_fusion2 = _PFF(c), _PFF(c)
        self.dyt1, self.dyt2 = _ADT(c), _ADT(c)
        self.attn = _CSATSSA(c, num_heads)
        self.ffn = _EDFFN(c)
        self.residual_weight1 = nn.Parameter(torch.tensor(0.1))
        self.residual_weight2 = nn.Parameter(torch.tensor(0.1))


class C2ProgressiveTSSA_Fusion(YadModule):
    """nn/modules/block.py:2700-2710 (C2PSA base :1010-1049): C2PTSSA(c1, c2, n=1, e=0.5)"""

    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2 and n == 1
        self.c = int(c1 * e)
        assert self.c // 64 == 2, "the attention kernels are built for 2 heads of 64 channels (c = 128)"
        self.cv1 = _conv_holder(c1, 2 * self.c, 1, 1)
        self.cv2 = _conv_holder(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(_PTSSA(self.c, max(1, self.c // 64)) for _ in range(n)))

    def forward(self, x):
        self._check_eval()
        a = as_act(x)
        return Fn.c2ptssa(self._ctx(a), "m", a).nchw()


C2PTSSA = C2ProgressiveTSSA_Fusion


class ELA_HSFPN(YadModule):
    """nn/modules/block.py:1408-1424 ELA_HSFPN(in_planes, flag=True)"""

    def __init__(self, in_planes, flag=True):
        super().__init__()
        self.conv1x1 = nn.Sequential(nn.Conv1d(in_planes, in_planes, 7, padding=3), nn.GroupNorm(16, in_planes), nn.Sigmoid())
        self.flag = flag

    def forward(self, x):
        self._check_eval()
        a = as_act(x)
        return Fn.ela_hsfpn(self._ctx(a), "m", a, self.flag).nchw()


class MonaOp(_Holder):
    """nn/modules/mona.py:12-34 (parameter holder; Mona runs the fused forward)"""

    def __init__(self, in_features):
        super().__init__()
        self.conv1 = nn.Conv2d(in_features, in_features, kernel_size=3, padding=1, groups=in_features)
        self.conv2 = nn.Conv2d(in_features, in_features, kernel_size=5, padding=2, groups=in_features)
        self.conv3 = nn.Conv2d(in_features, in_features, kernel_size=7, padding=3, groups=in_features)
        self.projector = nn.Conv2d(in_features, in_features, kernel_size=1)


class Mona(YadModule):
    """nn/modules/mona.py:36-64 Mona(in_dim): same constructor, attribute and state-dict key names (project1, project2, adapter_conv.*, norm, gamma,
    gammax).  SURVEY.md section 8f rank 3: the adapter of the C2TSSA_DYT_Mona_EDFFN sibling yamls; not part of the 701 yaml."""

    def __init__(self, in_dim):
        super().__init__()
        self.project1 = nn.Conv2d(in_dim, 64, 1)
        self.project2 = nn.Conv2d(64, in_dim, 1)
        self.dropout = nn.Dropout(p=0.1)
        self.adapter_conv = MonaOp(64)
        self.norm = nn.LayerNorm(in_dim)
        self.gamma = nn.Parameter(torch.ones(in_dim, 1, 1) * 1e-6)
        self.gammax = nn.Parameter(torch.ones(in_dim, 1, 1))

    def forward(self, x, hw_shapes=None):
        self._check_eval()
        a = as_act(x)
        return Fn.mona(self._ctx(a), "m", a).nchw()


class DynamicTanh(_Holder):
    """nn/modules/block.py:1624-1641 (parameter holder)"""

    def __init__(self, normalized_shape, channels_last=False, alpha_init_value=0.5):
        super().__init__()
        assert not channels_last
        self.alpha = nn.Parameter(torch.ones(1) * alpha_init_value)
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))


class AttentionTSSA(_Holder):
    """nn/modules/block.py:1646-1663 (parameter holder)"""

    def __init__(self, dim, num_heads=8):
        super().__init__()
        self.heads = num_heads
        self.qkv = nn.Linear(dim, dim, bias=False)
        self.temp = nn.Parameter(torch.ones(num_heads, 1))
        self.to_out = nn.Sequential(nn.Linear(dim, dim), nn.Dropout(0.0))


class TSSAlock_DYT_Mona_EDFFN(_Holder):
    """nn/modules/block.py:1685-1694 (parameter holder; C2TSSA_DYT_Mona_EDFFN runs the fused forward)"""

    def __init__(self, c, attn_ratio=0.5, num_heads=4, shortcut=True):
        super().__init__()
        assert shortcut
        self.ffn = _EDFFN(c)
        self.dyt1, self.dyt2 = DynamicTanh(c), DynamicTanh(c)
        self.mona1, self.mona2 = Mona(c), Mona(c)
        self.attn = AttentionTSSA(c, num_heads=num_heads)


class C2TSSA_DYT_Mona_EDFFN(YadModule):
    """nn/modules/block.py:1705-1709 (C2PSA base :1010-1049): C2TSSA_DYT_Mona_EDFFN(c1, c2, n=1, e=0.5) -- layer 10 of the yolo11-mona / 687 / 689 /
    697 sibling yamls.  Same constructor, attribute and state-dict key names as the reference."""

    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        assert self.c % 64 == 0, "head_dim is 64: the hidden width must be a multiple of 64"
        self.n = n
        self.cv1 = _conv_holder(c1, 2 * self.c, 1, 1)
        self.cv2 = _conv_holder(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(TSSAlock_DYT_Mona_EDFFN(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))

    def forward(self, x):
        self._check_eval()
        a = as_act(x)
        return Fn.c2tssa_dyt_mona_edffn(self._ctx(a), "m", a, self.n).nchw()


class Attention(_Holder):
    """nn/modules/block.py:874-925 Attention(dim, num_heads=8, attn_ratio=0.5) (parameter holder)"""

    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        self.qkv = _conv_holder(dim, dim + 2 * self.key_dim * num_heads, 1)
        self.proj = _conv_holder(dim, dim, 1)
        self.pe = _conv_holder(dim, dim, 3, 1, g=dim)


class PSABlock(_Holder):
    """nn/modules/block.py:928-964 PSABlock(c, attn_ratio=0.5, num_heads=4, shortcut=True) (parameter holder; C2PSA runs the fused forward)"""

    def __init__(self, c, attn_ratio=0.5, num_heads=4, shortcut=True):
        super().__init__()
        assert shortcut
        self.attn = Attention(c, attn_ratio=attn_ratio, num_heads=num_heads)
        self.ffn = nn.Sequential(_conv_holder(c, c * 2, 1), _conv_holder(c * 2, c, 1))
        self.add = shortcut


class C2PSA(YadModule):
    """nn/modules/block.py:1010-1049 C2PSA(c1, c2, n=1, e=0.5): layer 10 of the stock yolo11 yaml and of the sibling ablation yamls that keep the
    PSABlock attention.  Same constructor, attribute and state-dict key names as the reference."""

    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        assert self.c % 64 == 0, "head_dim is 64: the hidden width must be a multiple of 64"
        self.n = n
        self.cv1 = _conv_holder(c1, 2 * self.c, 1, 1)
        self.cv2 = _conv_holder(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))

    def forward(self, x):
        self._check_eval()
        a = as_act(x)
        return Fn.c2psa(self._ctx(a), "m", a, self.n).nchw()


class SEBlock(_Holder):
    """nn/modules/block.py:2049-2064 (parameter holder)"""

    def __init__(self, c1, r=16):
        super().__init__()
        c_ = int(c1 / r)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Conv2d(c1, c_, 1, bias=False), nn.ReLU(inplace=True), nn.Conv2d(c_, c1, 1, bias=False), nn.Sigmoid())


class StandardFFN(_Holder):
    """nn/modules/block.py:2066-2078 (parameter holder)"""

    def __init__(self, c1, expansion=2, bias=False):
        super().__init__()
        assert not bias
        c_ = int(c1 * expansion)
        self.cv1 = nn.Conv2d(c1, c_, 1, bias=bias)
        self.act = nn.GELU()
        self.cv2 = nn.Conv2d(c_, c1, 1, bias=bias)


class SimpleFeatureProcessor(_Holder):
    """nn/modules/block.py:2080-2096 (parameter holder)"""

    def __init__(self, c):
        super().__init__()
        self.norm = nn.GroupNorm(num_groups=max(1, c // 32), num_channels=c)
        self.conv_dw = nn.Conv2d(c, c, 3, padding=1, groups=c)
        self.act = nn.GELU()
        self.conv_pw = nn.Conv2d(c, c, 1)


class ProgressiveTSSA_Fusion0(_Holder):
    """nn/modules/block.py:2147-2172 (parameter holder; C2SFA runs the fused forward)"""

    def __init__(self, c, attn_ratio=0.5, num_heads=4, shortcut=True):
        super().__init__()
        assert shortcut
        self.c, self.add = c, shortcut
        self.pre_attn_block = SimpleFeatureProcessor(c)
        self.attn = SEBlock(c)
        self.pre_ffn_block = SimpleFeatureProcessor(c)
        self.ffn = StandardFFN(c, expansion=2, bias=False)
        self.residual_weight1 = nn.Parameter(torch.tensor(0.1))
        self.residual_weight2 = nn.Parameter(torch.tensor(0.1))


class C2SFA(YadModule):
    """nn/modules/block.py:2358-2373 C2SFA(c1, c2, n=1, e=0.5) (C2PSA base :1010-1049): layer 10 of yolo11-hsfpn+C2SFA.yaml and two more sibling yamls
    (SURVEY.md section 8f rank 3).  Same constructor, attribute and state-dict key names as the reference."""

    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        assert self.c % 32 == 0, "the hidden width must be a multiple of 32 (GroupNorm groups of 32 channels, 8-channel vectors)"
        self.n = n
        self.cv1 = _conv_holder(c1, 2 * self.c, 1, 1)
        self.cv2 = _conv_holder(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(ProgressiveTSSA_Fusion0(self.c, num_heads=max(1, self.c // 64), shortcut=True) for _ in range(n)))

    def forward(self, x):
        self._check_eval()
        a = as_act(x)
        return Fn.c2sfa(self._ctx(a), "m", a, self.n).nchw()


class Multiply(nn.Module):
    """nn/modules/block.py:1442-1447"""

    def forward(self, x):
        a, b = as_act(x[0]), as_act(x[1])
        return ops.eltwise(1, a, b, Act.empty(a.n, a.h, a.w, a.c, a.dtype, a.device)).nchw()


class Add(nn.Module):
    """nn/modules/block.py:1448-1453 (stack + sum)"""

    def forward(self, x):
        acc = as_act(x[0])
        for t in x[1:]:
            acc = ops.eltwise(0, acc, as_act(t), Act.empty(acc.n, acc.h, acc.w, acc.c, acc.dtype, acc.device))
        return acc.nchw()


class Fusion(YadModule):
    """nn/modules/block.py:1500-1537 Fusion(inc_list, fusion='bifpn') -- only the 'bifpn' mode is on the YOLO-AD-Refine path"""

    def __init__(self, inc_list, fusion="bifpn"):
        super().__init__()
        if fusion != "bifpn":
            raise NotImplementedError("only Fusion(..., 'bifpn') is part of the YOLO-AD-Refine hot path")
        self.fusion = fusion
        self.fusion_weight = nn.Parameter(torch.ones(len(inc_list), dtype=torch.float32), requires_grad=True)
        self.epsilon = 1e-4

    def forward(self, x):
        a = [as_act(t) for t in x]
        return Fn.fusion_bifpn(self._ctx(a[0]), "m", a).nchw()


class YadConv2d(YadModule):
    """the yaml's plain nn.Conv2d(c1, c2, 1) laterals (layers 12, 15, 22): same parameter names (weight, bias)"""

    def __init__(self, c1, c2, k=1, s=1, p=0, bias=True):
        super().__init__()
        ref = nn.Conv2d(c1, c2, k, s, p, bias=bias)
        self.weight, self.bias = ref.weight, ref.bias
        self.stride_, self.k = s, k

    def forward(self, x):
        a = as_act(x)
        ctx = self._ctx(a)
        return Fn.conv(ctx, a, ctx.P.conv("m.weight", "m.bias" if self.bias is not None else None, stride=self.stride_)).nchw()


class YadConvTranspose2d(YadModule):
    """the yaml's nn.ConvTranspose2d(c1, c2, 3, 2, 1, 1) upsamplers (layers 13, 20)"""

    def __init__(self, c1, c2, k=3, s=2, p=1, op=1):
        super().__init__()
        assert (k, s, p, op) == (3, 2, 1, 1)
        ref = nn.ConvTranspose2d(c1, c2, k, s, p, op)
        self.weight, self.bias = ref.weight, ref.bias

    def forward(self, x):
        a = as_act(x)
        ctx = self._ctx(a)
        return Fn.conv(ctx, a, ctx.P.conv("m.weight", "m.bias", stride=2, transposed=True), mode=ops.CONV_TRANSPOSED).nchw()


# ---------------------------------------------------------------------------------------------------------------------------
# AYHead1
# ---------------------------------------------------------------------------------------------------------------------------
def _conv_gn_holder(c1, c2, k=1):
    """nn/modules/head.py:1265-1279 (second, effective Conv_GN)"""
    h = _Holder()
    h.conv = nn.Conv2d(c1, c2, k, 1, autopad(k), bias=False)
    h.gn = nn.GroupNorm(gn_groups(c2), c2)
    return h


class _TaskDecomposition(_Holder):
    """nn/modules/head.py:626-650"""

    def __init__(self, feat_channels, stacked_convs=1, la_down_rate=16):
        super().__init__()
        in_ch = feat_channels * stacked_convs
        self.la_conv1 = nn.Conv2d(in_ch, in_ch // la_down_rate, 1)
        self.la_conv2 = nn.Conv2d(in_ch // la_down_rate, stacked_convs, 1, padding=0)
        self.reduction_conv = _conv_gn_holder(in_ch, feat_channels, 1)


class _CoordAtt(_Holder):
    """nn/modules/head.py:671-686"""

    def __init__(self, inp, oup, reduction=32):
        super().__init__()
        mip = max(8, inp // reduction)
        self.conv1 = nn.Conv2d(inp, mip, 1)
        self.bn1 = nn.BatchNorm2d(mip, eps=1e-3, momentum=0.03)
        self.conv_h = nn.Conv2d(mip, oup, 1)
        self.conv_w = nn.Conv2d(mip, oup, 1)


class AYHead(YadModule):
    """nn/modules/head.py:1049-1252 AYHead1 (aliased as AYHead at :1666): AYHead(nc=80, ch=())"""
    dynamic = False
    export = False
    shape = None
    anchors = torch.empty(0)
    strides = torch.empty(0)
    format = None

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc, self.nl, self.reg_max = nc, len(ch), 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        self.ch = ch
        hidc = max(ch) if ch else 512
        t = hidc // 2
        self.stems = nn.ModuleList(_conv_gn_holder(c, hidc, 1) for c in ch)
        self.share_conv = nn.Sequential(_conv_gn_holder(hidc, t, 3), _conv_gn_holder(t, t, 3))
        self.cls_decomp, self.reg_decomp = _TaskDecomposition(t), _TaskDecomposition(t)
        self.rep_block_cls = _Holder()
        self.rep_block_cls.conv1, self.rep_block_cls.conv2 = _conv_gn_holder(t, t, 3), _conv_gn_holder(t, t, 3)
        self.coord_attention_reg = _CoordAtt(t, t)
        self.cross_task = _Holder()
        self.cross_task.cls_to_reg, self.cross_task.reg_to_cls = nn.Conv2d(t, t, 1), nn.Conv2d(t, t, 1)
        self.cross_task.cls_gate = nn.Sequential(nn.Conv2d(2 * t, t, 1), nn.Sigmoid())
        self.cross_task.reg_gate = nn.Sequential(nn.Conv2d(2 * t, t, 1), nn.Sigmoid())
        self.spatial_conv_offset = nn.Conv2d(t, 27, 3, padding=1)
        self.DyDCNV2 = _Holder()
        self.DyDCNV2.conv = _Holder()
        self.DyDCNV2.conv.weight = nn.Parameter(torch.empty(t, t, 3, 3).uniform_(-1.0 / (9 * t) ** 0.5, 1.0 / (9 * t) ** 0.5))
        self.DyDCNV2.norm = nn.GroupNorm(16, t)
        self.cls_prob_conv = nn.Sequential(nn.Conv2d(t, t // 2, 1), nn.ReLU(), nn.Conv2d(t // 2, 1, 3, padding=1), nn.Sigmoid())
        self.cv2 = nn.Conv2d(t, 4 * self.reg_max, 1)
        self.cv3 = nn.Conv2d(t, self.nc, 1)
        self.scale = nn.ModuleList()
        for _ in range(self.nl):
            s = _Holder()
            s.scale = nn.Parameter(torch.tensor(1.0, dtype=torch.float))
            self.scale.append(s)
        self.dfl = _Holder()
        self.dfl.conv = nn.Conv2d(self.reg_max, 1, 1, bias=False).requires_grad_(False)
        self.dfl.conv.weight.data[:] = torch.arange(self.reg_max, dtype=torch.float).view(1, self.reg_max, 1, 1)
        self.initialize_biases()

    def initialize_biases(self):
        """nn/modules/head.py:1206-1225: default strides [8, 16, 32] when unset, YOLO-style bias init"""
        if not torch.is_tensor(self.stride) or self.stride.sum() == 0:
            self.stride = torch.tensor([8.0, 16.0, 32.0][: self.nl])
        self.cv2.bias.data[:] = 1.0
        self.cv3.bias.data[: self.nc] = math.log(5 / self.nc / (640 / float(self.stride.mean())) ** 2)

    bias_init = initialize_biases

    def forward(self, x):
        self._check_eval()
        acts = [as_act(t) for t in x]
        ctx = self._ctx(acts[0])
        y, outs = Fn.ayhead(ctx, "m", acts, strides=[float(s) for s in self.stride], nc=self.nc, reg_max=self.reg_max)
        raw = [o.nchw() for o in outs]
        return y if self.export else (y, raw)

    def decode_bboxes(self, bboxes, anchors=None):
        raise NotImplementedError("decode is fused into yad_decode (DFL + anchors + dist2bbox + stride + sigmoid in one kernel)")


AYHead1 = AYHead


# ---------------------------------------------------------------------------------------------------------------------------
# the model of the 701 yaml at scale n built from the mirrors (nn/tasks.py:943-1108 parse_model result; :141-168 _predict_once)
# ---------------------------------------------------------------------------------------------------------------------------
def build_yolo_ad_refine(nc=80):
    """Returns (nn.Sequential of the 34 layers with the reference's `model.N` numbering, routing table)."""
    L = [
        (-1, Conv(3, 16, 3, 2)), (-1, Conv(16, 32, 3, 2)), (-1, C3k2(32, 64, 1, False, 0.25)), (-1, Conv(64, 64, 3, 2)),
        (-1, C3k2(64, 128, 1, False, 0.25)), (-1, Conv(128, 128, 3, 2)), (-1, C3k2_MLCA(128, 128, 1, True)), (-1, Conv(128, 256, 3, 2)),
        (-1, C3k2_MLCA(256, 256, 1, True)), (-1, SPPF(256, 256, 5)), (-1, C2PTSSA(256, 256, 1)),
        (10, ELA_HSFPN(256)), (-1, YadConv2d(256, 128, 1)), (12, YadConvTranspose2d(128, 128)), (6, ELA_HSFPN(128)), (-1, YadConv2d(128, 128, 1)),
        (13, ELA_HSFPN(128, False)), ([15, 16], Multiply()), ([-1, 13], Add()), (-1, C3k2_MLCA(128, 128, 1, False)),
        (19, YadConvTranspose2d(128, 128)), (4, ELA_HSFPN(128)), (-1, YadConv2d(128, 128, 1)), (20, ELA_HSFPN(128, False)), ([22, 23], Multiply()),
        ([-1, 20], Add()), (-1, C3k2_MLCA(128, 128, 1, False)),
        (26, Conv(128, 128, 3, 2)), ([-1, 19], Fusion([128, 128], "bifpn")), (-1, C3k2(128, 128, 1, False)),
        (29, Conv(128, 128, 3, 2)), ([-1, 12], Fusion([128, 128], "bifpn")), (-1, C3k2(128, 128, 1, False)),
        ([26, 29, 32], AYHead(nc, (128, 128, 128))),
    ]
    seq = nn.Sequential(*[m for _, m in L])
    for i, (f, m) in enumerate(L):
        m.i, m.f = i, f
    return seq


class YoloADRefine(nn.Module):
    """Minimal stand-in for DetectionModel (nn/tasks.py:309-398) built from the mirrors: `.model` Sequential with the reference's key names."""

    def __init__(self, nc=80):
        super().__init__()
        self.model = build_yolo_ad_refine(nc)
        self.stride = self.model[-1].stride
        self.nc = nc

    def forward(self, x):
        """nn/tasks.py:141-168 _predict_once"""
        y = []
        for m in self.model:
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            x = m(x)
            y.append(x)
        return x
