"""Micro-probe for one convolution shape through the C ABI: times impl 1 (SIMT) and impl 2 (tcgen05) with CUDA events.
usage: python tools/conv_probe.py CIN COUT K STRIDE HW [BATCH] [REPS] [IMPLS]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.ops import Act
from yolo_ad_refine_b200.weights import pack_conv

cin, cout, k, s, hw = [int(v) for v in sys.argv[1:6]]
n = int(sys.argv[6]) if len(sys.argv) > 6 else 64
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 10
impls = [int(v) for v in sys.argv[8].split(",")] if len(sys.argv) > 8 else [1, 2]
dev, dt = "cuda", torch.bfloat16
x = Act(torch.randn(n, hw, hw, cin, device=dev).to(dt))
cw = pack_conv(torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5, torch.randn(cout) * 0.1, dt, dev, s)
ho = (hw + 2 * (k // 2) - k) // s + 1
y = Act.empty(n, ho, ho, cw.cout, dt, dev)
if os.environ.get("SLICE") == "1":  # output = a channel window of a wider (concat) buffer
    y = Act.empty(n, ho, ho, cw.cout + 32, dt, dev).slice(16, cw.cout)
add = Act(torch.randn(n, ho, ho, cw.cout, device=dev).to(dt)) if os.environ.get("ADD") == "1" else None
mul = Act(torch.rand(n, ho, ho, cw.cout, device=dev).to(dt)) if os.environ.get("MUL") == "1" else None
gate = (Act(torch.rand(n, ho, 1, cw.cout, device=dev).to(dt)), Act(torch.rand(n, ho, 1, cw.cout, device=dev).to(dt))) if os.environ.get("GATE") == "1" else None
act = {"silu": ops.ACT_SILU, "none": ops.ACT_NONE, "sigmoid": ops.ACT_SIGMOID}[os.environ.get("ACT", "silu")]
flops = 2.0 * n * ho * ho * cw.cout * k * k * cin
byts = 2.0 * (x.buf.numel() + y.buf.numel() * (1 + (add is not None) + (mul is not None)) + cw.w.numel())
deform = os.environ.get("DEFORM") == "1"
om = Act((torch.randn(n, hw, hw, 32, device=dev) * float(os.environ.get("DEFORM_STD", "1.0"))).to(dt)) if deform else None
mode = ops.CONV_DEFORM if deform else ops.CONV_NORMAL
for impl in impls:
    for _ in range(3):
        ops.conv2d(x, cw.w, y, bias=cw.b, kh=k, kw=k, stride=s, pad_h=k // 2, pad_w=k // 2, act=act, impl=impl, mode=mode, offmask=om, add=add, mul=mul, gate=gate)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.conv2d(x, cw.w, y, bias=cw.b, kh=k, kw=k, stride=s, pad_h=k // 2, pad_w=k // 2, act=act, impl=impl, mode=mode, offmask=om, add=add, mul=mul, gate=gate)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"impl {impl}: {cin}->{cout} k{k} s{s} {hw}x{hw} n{n}: {ms*1000:.1f} us  {flops/ms/1e9:.1f} TFLOP/s  {byts/ms/1e6:.0f} GB/s (algorithmic)")
