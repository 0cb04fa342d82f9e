"""Micro-probe of yad_mha (nn.MultiheadAttention core of CrossScaleAttention_TSSA, nn/modules/block.py:2479-2488) at the benchmark shape:
batch 64, T = 3 scales x 20 x 20 = 1200 tokens, 2 heads x 64.  usage: python tools/probe_mha.py [N] [T] [HEADS]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_ad_refine_b200 import ops  # noqa: E402
from yolo_ad_refine_b200.ops import Act  # noqa: E402

n, t, heads = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 1200, 2)))
c = heads * 64
dt = torch.bfloat16
qkv = Act((torch.randn(n, t, 1, 3 * c, device="cuda") * 0.5).to(dt))
out = Act.empty(n, t, 1, c, dt, "cuda")
for _ in range(3):
    ops.mha(qkv, heads, out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    ops.mha(qkv, heads, out)
e.record()
torch.cuda.synchronize()
us = s.elapsed_time(e) / 10 * 1e3
flops = 4.0 * n * heads * t * t * 64
print(f"mha n{n} T{t} heads{heads}: {us:.1f} us  {flops / us / 1e6:.1f} TFLOP/s")
# numerics against torch (fp32 softmax attention on the bf16 inputs)
x = qkv.torch().float().reshape(n, t, 3, heads, 64)
q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(n, t, c)
err = (out.torch().float().reshape(n, t, c) - ref).abs().max().item()
print(f"max abs error vs torch sdpa: {err:.2e}")
