"""Micro-benchmark of the layer-10 blocks of the sibling yamls at the layer-10 geometry (256 channels, 20x20, batch 64,
bf16): C2PSA, C2SFA, C2TSSA_DYT_Mona_EDFFN and the 701 yaml's own C2ProgressiveTSSA_Fusion, each captured in a CUDA graph (the way the engine runs
blocks) and timed with CUDA events on the capturing stream.  Algorithmic bytes = the block's input read + output written once.
usage: python tools/bench_blocks.py [C H W BATCH]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from yolo_ad_refine_b200 import functional as Fn
from yolo_ad_refine_b200 import modules as M
from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.ops import Act

c, h, w, n = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (256, 20, 20, 64)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
dev = "cuda"
BLOCKS = [("C2PSA", lambda: M.C2PSA(c, c, 1), lambda ctx, x: Fn.c2psa(ctx, "m", x, 1)),
          ("C2SFA", lambda: M.C2SFA(c, c, 1), lambda ctx, x: Fn.c2sfa(ctx, "m", x, 1)),
          ("C2TSSA_DYT_Mona_EDFFN", lambda: M.C2TSSA_DYT_Mona_EDFFN(c, c, 1), lambda ctx, x: Fn.c2tssa_dyt_mona_edffn(ctx, "m", x, 1)),
          ("C2ProgressiveTSSA_Fusion", lambda: M.C2ProgressiveTSSA_Fusion(c, c, 1), lambda ctx, x: Fn.c2ptssa(ctx, "m", x))]
for name, make, run in BLOCKS:
    m = make().eval().to(dev)
    g0 = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for k, p in m.state_dict().items():
            if p.dtype.is_floating_point:
                p.copy_((1.0 + 0.1 * torch.randn(p.shape, generator=g0)) if k.endswith(("running_var", "bn.weight", "norm.weight"))
                        else 0.05 * torch.randn(p.shape, generator=g0))
            if k.endswith("running_var"):
                p.abs_()
    m.refresh()
    x = Act(torch.randn(n, h, w, c, device=dev).bfloat16())
    ctx = m._ctx(x)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            run(ctx, x)
        s.synchronize()
        before = ops.LAUNCHES
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            run(ctx, x)
        launches = ops.LAUNCHES - before
        for _ in range(5):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        reps = 100
        for _ in range(reps):
            g.replay()
        b.record(s)
        s.synchronize()
    ms = a.elapsed_time(b) / reps
    alg = n * h * w * c * 2 * 2
    print(json.dumps({"metric": f"{name} img/s", "value": n / ms * 1e3, "unit": "img/s", "us_per_call": ms * 1e3, "launches": launches,
                      "config": {"workload": f"{name}({c}, {c}, 1) on {n}x{c}x{h}x{w} bf16, CUDA-graph replay"},
                      "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": alg / ms / 1e6 / peak,
                                   "algorithmic_bytes": alg, "note": "latency-bound chain of small kernels at 20x20; the working set fits the 126 MB L2"}}))
