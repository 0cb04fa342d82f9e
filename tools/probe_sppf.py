"""Times yad_sppf_pool at the layer-9 shape (bf16): python tools/probe_sppf.py [N H W C]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_ad_refine_b200 import ops  # noqa: E402
from yolo_ad_refine_b200.ops import Act  # noqa: E402

n, h, w, c = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (64, 20, 20, 128)
cat = Act(torch.randn(n, h, w, 4 * c, device="cuda").bfloat16())
args = (cat.slice(0, c), cat.slice(c, c), cat.slice(2 * c, c), cat.slice(3 * c, c))
for _ in range(3):
    ops.sppf_pool(*args)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ops.sppf_pool(*args)
b.record()
torch.cuda.synchronize()
print(f"sppf_pool {n}x{h}x{w}x{c}: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
