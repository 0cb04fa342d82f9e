"""ncu probe: yad_gn_stats / yad_gn_apply / yad_norm_bwd on one BatchNorm-shaped tensor (batch viewed as one image)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_ad_refine_b200 import ops  # noqa: E402
from yolo_ad_refine_b200.ops import Act  # noqa: E402

n, h, w, c = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (1, 10240, 80, 128)))
groups = c if n == 1 else 16
dt = torch.bfloat16
x = Act(torch.randn(n, h, w, c, device="cuda").to(dt))
dy = Act(torch.randn(n, h, w, c, device="cuda").to(dt))
y, dx = Act.empty(n, h, w, c, dt, "cuda"), Act.empty(n, h, w, c, dt, "cuda")
gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
stats, sums = torch.empty(n, groups, 2, dtype=torch.float64, device="cuda"), torch.empty(n, groups, 2, dtype=torch.float64, device="cuda")
dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
for _ in range(3):
    ops.group_norm(x, y, stats, groups, gamma, beta, 1e-3, ops.ACT_SILU)
    ops.norm_bwd(x, dy, stats, groups, gamma, beta, 1e-3, ops.ACT_SILU, sums, dg, db, dx, 0)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    ops.norm_bwd(x, dy, stats, groups, gamma, beta, 1e-3, ops.ACT_SILU, sums, dg, db, dx, 0)
e.record()
torch.cuda.synchronize()
mb = n * h * w * c * 2 / 1e6
print(f"norm_bwd {n}x{h}x{w}x{c}: {s.elapsed_time(e) / 10 * 1e3:.1f} us per call, tensor {mb:.0f} MB -> {5 * mb / (s.elapsed_time(e) / 10) / 1e3:.2f} TB/s (5 passes)")
