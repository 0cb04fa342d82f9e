"""Bring-up probe for the haloed-patch 3x3 path of conv_v2.cu: one run per tap with the weights of all other taps zeroed, against torch; prints the
error per tap so that a wrong descriptor convention shows which taps it breaks.  Variants through YAD_CONV2_PW / YAD_CONV2_BO (read once per process)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.ops import Act
from yolo_ad_refine_b200.weights import pack_conv

BF, DEV = torch.bfloat16, "cuda"
cin, cout, hw, n = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (64, 64, 16, 1)
g = torch.Generator().manual_seed(1)
x = torch.randn(n, cin, hw, hw, generator=g).to(BF).float()
wfull = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(BF).float()
xa = Act.from_nchw(x.to(DEV), BF)
res = []
for t in list(range(9)) + [-1]:
    w = wfull.clone()
    if t >= 0:
        m = torch.zeros(3, 3)
        m[t // 3, t % 3] = 1
        w = w * m
    ref = F.conv2d(x, w, None, 1, 1)
    cw = pack_conv(w, None, BF, DEV, 1)
    out = Act.empty(n, hw, hw, cw.cout, BF, DEV)
    ops.conv2d(xa, cw.w, out, kh=3, kw=3, pad_h=1, pad_w=1, impl=4)
    got = out.nchw().float().cpu()[:, :cout]
    err = float((got - ref).abs().max() / ref.abs().mean())
    # where is it wrong: per output row / column error profile (first image)
    e = (got - ref).abs()[0].amax(0)
    bad_rows = [i for i in range(hw) if float(e[i].max()) > 0.05 * float(ref.abs().mean())]
    bad_cols = [i for i in range(hw) if float(e[:, i].max()) > 0.05 * float(ref.abs().mean())]
    res.append(err)
    print(f"tap {t:2d}: max err / mean|ref| = {err:.3e}  bad rows {bad_rows[:20]} bad cols {bad_cols[:20]}")
print("PW", os.environ.get("YAD_CONV2_PW"), "BO", os.environ.get("YAD_CONV2_BO"), "OK" if max(res) < 0.05 else "WRONG")
