"""Times yad_patch_filter (EDFFN 8x8 spectral filter) at the layer-10 shape: python tools/probe_patch_filter.py [N H W C]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_ad_refine_b200 import ops  # noqa: E402
from yolo_ad_refine_b200.ops import Act  # noqa: E402
from yolo_ad_refine_b200.weights import edffn_spectral_matrix  # noqa: E402

n, h, w, c = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (64, 20, 20, 128)
dev, dt = "cuda", torch.bfloat16
x = Act(torch.randn(n, h, w, c, device=dev).to(dt))
add = Act(torch.randn(n, h, w, c, device=dev).to(dt))
y = Act.empty(n, h, w, c, dt, dev)
m = edffn_spectral_matrix(1.0 + 0.1 * torch.randn(c, 1, 1, 8, 5)).to(dev)
for _ in range(3):
    ops.patch_filter(x, m, y, alpha=0.5, add=add)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ops.patch_filter(x, m, y, alpha=0.5, add=add)
b.record()
torch.cuda.synchronize()
print(f"patch_filter {n}x{h}x{w}x{c}: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
