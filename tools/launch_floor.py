"""Per-launch floor of dependent small kernels inside a CUDA graph (the regime of the 20x20 / 40x40 layers: layer 10, the P4 / P5 head levels):
a chain of K dependent calls (ping-pong buffers) is captured once and replayed; prints microseconds per call.  With and without programmatic
dependent launch.  usage: python tools/launch_floor.py [HW] [BATCH]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.ops import Act
from yolo_ad_refine_b200.weights import pack_conv

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
K = 40
dev, dt = "cuda", torch.bfloat16


def chain(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            fn(0)
            fn(1)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(K):
                fn(i & 1)
        for _ in range(3):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(20):
            g.replay()
        b.record(s)
        s.synchronize()
    return a.elapsed_time(b) / 20 / K * 1e3


def run(tag):
    res = {}
    for c in (64, 128):
        bufs = [Act(torch.randn(n, hw, hw, c, device=dev).to(dt)) for _ in range(2)]
        res[f"eltwise c{c}"] = chain(lambda i: ops.eltwise(0, bufs[i], bufs[i], bufs[1 - i], alpha=0.5, beta=0.5))
        for k in (1, 3):
            cw = pack_conv(torch.randn(c, c, k, k) / (c * k * k) ** 0.5, torch.randn(c) * 0.1, dt, dev, 1)
            for impl in (2, 5):
                res[f"conv{k}x{k} c{c} impl{impl}"] = chain(lambda i: ops.conv2d(bufs[i], cw.w, bufs[1 - i], bias=cw.b, kh=k, kw=k, pad_h=k // 2, pad_w=k // 2,
                                                                                     act=ops.ACT_SILU, impl=impl))
        dw = torch.randn(9, c, device=dev)
        res[f"dwconv3 c{c}"] = chain(lambda i: ops.dwconv(bufs[i], dw, bufs[1 - i], k=3))
        st = torch.zeros(n, 16, 2, dtype=torch.float64, device=dev)
        st[..., 1] = 1000.0
        gam, bet = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        res[f"gn_apply c{c}"] = chain(lambda i: ops.group_norm(bufs[i], bufs[1 - i], st, 16, gam, bet, 1e-5, ops.ACT_SILU, None, stats_ready=True))
    print(json.dumps({"tag": tag, "hw": hw, "batch": n, "us_per_call": {k: round(v, 2) for k, v in res.items()}}))


run("pdl0")
ops.lib().yad_set_pdl(1)
run("pdl1")
