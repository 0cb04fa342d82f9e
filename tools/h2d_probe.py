"""pinned host -> device copy bandwidth of one batch-64 uint8 image tensor (78.6 MB), alone and beside a running engine"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
x = torch.randint(0, 256, (64, 3, 640, 640), dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    d.copy_(x, non_blocking=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"H2D {x.numel() / 1e6:.1f} MB: {ms:.3f} ms = {x.numel() / ms / 1e6:.1f} GB/s")
