"""Times yad_dwconv on one shape (bf16, bias + GELU epilogue): python tools/probe_dwconv.py N H W C K [REPS]
(the command the dwconv ncu captures of profiles/ run)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_ad_refine_b200 import ops  # noqa: E402
from yolo_ad_refine_b200.ops import Act  # noqa: E402

n, h, w, c, k = [int(a) for a in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 50
dt, dev = torch.bfloat16, "cuda"
x = Act(torch.randn(n, h, w, c, device=dev).to(dt))
y = Act.empty(n, h, w, c, dt, dev)
wk, bk = torch.randn(k * k, c, device=dev), torch.randn(c, device=dev)
for _ in range(3):
    ops.dwconv(x, wk, y, bias=bk, k=k, act=ops.ACT_GELU)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    ops.dwconv(x, wk, y, bias=bk, k=k, act=ops.ACT_GELU)
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) / reps * 1e3
nb = 2 * n * h * w * c * 2
print(f"dwconv k{k} {n}x{h}x{w}x{c}: {us:.1f} us, {nb / us / 1e6:.2f} TB/s algorithmic")
