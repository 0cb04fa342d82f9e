"""Times yad_stem_conv on a batch of uint8 640x640 images (bf16 output): python tools/probe_stem.py [BATCH]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_ad_refine_b200 import ops  # noqa: E402
from yolo_ad_refine_b200.ops import Act  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
img = torch.randint(0, 256, (n, 3, 640, 640), dtype=torch.uint8, device=dev)
wq = (torch.randn(16, 27, device=dev) / 5 / 255).contiguous()
b = torch.randn(16, device=dev) * 0.1
y = Act.empty(n, 320, 320, 16, torch.bfloat16, dev)
for _ in range(3):
    ops.stem_conv(img, wq, b, y)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ops.stem_conv(img, wq, b, y)
e.record()
torch.cuda.synchronize()
us = a.elapsed_time(e) / 20 * 1e3
nb = n * (3 * 640 * 640 + 320 * 320 * 16 * 2)
print(f"stem_conv u8 {n}x3x640x640 -> bf16 {n}x320x320x16: {us:.1f} us, {nb / us / 1e6:.2f} TB/s algorithmic")
