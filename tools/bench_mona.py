"""Micro-benchmark of the fused Mona adapter: five launches per call, bf16, CUDA-event timed in a CUDA graph replay
(the way the engine runs blocks).  Algorithmic bytes = x read twice (LayerNorm mix + residual) + y written; the materialised intermediates
(x1, p1, s, g: C + 3 x 64 channels written and read once) are reported as `materialised_bytes`.
usage: python tools/bench_mona.py [C H W BATCH]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from yolo_ad_refine_b200 import functional as Fn
from yolo_ad_refine_b200.modules import Mona
from yolo_ad_refine_b200.ops import Act

c, h, w, n = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (128, 20, 20, 64)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
dev = "cuda"
m = Mona(c).eval().to(dev)
with torch.no_grad():
    for p in m.parameters():
        p.normal_(0, 0.1)
x = Act(torch.randn(n, h, w, c, device=dev).bfloat16())
ctx = m._ctx(x)
out = Act.empty(n, h, w, c, torch.bfloat16, dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        Fn.mona(ctx, "m", x, out=out)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        Fn.mona(ctx, "m", x, out=out)
    for _ in range(5):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    reps = 200
    for _ in range(reps):
        g.replay()
    b.record(s)
    s.synchronize()
ms = a.elapsed_time(b) / reps
alg = n * h * w * c * 2 * 3
mat = n * h * w * (c + 3 * 64) * 2 * 2
print(json.dumps({"metric": "Mona adapter img/s", "value": n / ms * 1e3, "unit": "img/s", "us_per_call": ms * 1e3, "launches": 5,
                  "config": {"workload": f"Mona({c}) on {n}x{c}x{h}x{w} bf16, CUDA-graph replay"},
                  "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": alg / ms / 1e6 / peak,
                               "algorithmic_bytes": alg, "materialised_bytes": mat, "note": "working set fits the 126 MB L2 at 20x20"}}))
