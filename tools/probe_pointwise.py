"""Times the memory-bound inference kernels at the shapes the model uses (batch 64, bf16), device-timed over many launches on inputs larger than L2
where the shape allows: yad_gn_apply (head GroupNorm + SiLU), yad_dwconv (k3 / k7), yad_rowcol_mean / gate, yad_mlca_apply, yad_eltwise.
usage: python tools/probe_pointwise.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_ad_refine_b200 import ops  # noqa: E402
from yolo_ad_refine_b200.ops import Act  # noqa: E402

dt, dev = torch.bfloat16, "cuda"


def timed(fn, reps=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def report(name, us, nbytes):
    print(f"{name:44s} {us:8.1f} us  {nbytes / us / 1e6:6.2f} TB/s")


for (n, h, w, c) in ((64, 80, 80, 64), (64, 40, 40, 64), (64, 20, 20, 64), (64, 80, 80, 128), (64, 20, 20, 128)):
    x = Act(torch.randn(n, h, w, c, device=dev).to(dt))
    y = Act.empty(n, h, w, c, dt, dev)
    g = 16
    stats = torch.empty(n, g, 2, dtype=torch.float64, device=dev)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    ops.group_norm(x, y, stats, g, gamma, beta, 1e-5, ops.ACT_SILU)
    nb = n * h * w * c * 2
    report(f"gn_apply {n}x{h}x{w}x{c}", timed(lambda: ops.group_norm(x, y, stats, g, gamma, beta, 1e-5, ops.ACT_SILU, stats_ready=True)), 2 * nb)
    report(f"gn_apply+add {n}x{h}x{w}x{c}", timed(lambda: ops.group_norm(x, y, stats, g, gamma, beta, 1e-5, ops.ACT_SILU, add=x, stats_ready=True)), 3 * nb)
    for k in (3, 7):
        wk, bk = torch.randn(k * k, c, device=dev), torch.randn(c, device=dev)
        report(f"dwconv k{k} {n}x{h}x{w}x{c}", timed(lambda: ops.dwconv(x, wk, y, bias=bk, k=k, act=ops.ACT_GELU)), 2 * nb)
    rows, cols = Act.empty(n, h, 1, c, dt, dev), Act.empty(n, w, 1, c, dt, dev)
    report(f"rowcol_mean {n}x{h}x{w}x{c}", timed(lambda: ops.rowcol_mean(x, rows, cols)), nb)
    report(f"rowcol_gate {n}x{h}x{w}x{c}", timed(lambda: ops.rowcol_gate(x, rows, cols, y)), 2 * nb)
    report(f"eltwise add {n}x{h}x{w}x{c}", timed(lambda: ops.eltwise(0, x, x, y)), 3 * nb)
