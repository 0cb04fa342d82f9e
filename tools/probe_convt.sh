# transposed convolution 128 -> 128 of the neck at batch 64: conv2_kernel phases vs conv_tma_kernel phases
python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.ops import Act
from yolo_ad_refine_b200.weights import pack_conv
dt, dev = torch.bfloat16, "cuda"
for hw in (40, 20):
    x = Act(torch.randn(64, hw, hw, 128, device=dev).to(dt))
    cw = pack_conv(torch.randn(128, 128, 3, 3) / 17, torch.randn(128) * 0.1, dt, dev, transposed=True)
    y = Act.empty(64, 2 * hw, 2 * hw, 128, dt, dev)
    for impl in (0, 5):
        f = lambda: ops.conv2d(x, cw.w, y, bias=cw.b, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, mode=ops.CONV_TRANSPOSED, impl=impl)
        for _ in range(3): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): f()
        b.record(); torch.cuda.synchronize()
        print(f"convT 128->128 {hw}->{2*hw} impl {impl}: {a.elapsed_time(b) * 100:.1f} us")
PY
