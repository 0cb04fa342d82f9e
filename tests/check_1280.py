"""TEST INFRASTRUCTURE (uses the CPU oracle as the checker; lives under tests/ for that reason, run it by hand on the GPU box).
BASELINE.json configs[4] on one rank: inference at 1280 x 1280, batch 4 per GPU (batch 32 sharded over 8 GPUs) -- parity of the fp32 build
against the CPU oracle on one image, then bf16 throughput of the per-GPU shard."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as om  # noqa: E402
from oracle import postprocess as op  # noqa: E402
from oracle import synth  # noqa: E402
from yolo_ad_refine_b200.engine import RefineEngine  # noqa: E402

sd = synth.make_state_dict(seed=1)
img = torch.from_numpy(synth.make_images(1, 1280, 1280, seed=4))
t0 = time.time()
y_ref, _ = om.forward(sd, img)
print(f"oracle 1280^2 b1: {time.time() - t0:.1f} s")
eng = RefineEngine(sd, batch=1, imgsz=1280, dtype=torch.float32, conv_impl=1, nms_args=dict(conf_thres=0.25, iou_thres=0.7, max_det=300))
y, _ = eng.forward(img)
err = float((y.cpu() - y_ref).abs().max() / y_ref.abs().max())
print("fp32 max rel err of y (B, 84, 33600):", err)
assert err < 1e-3
det = eng.detect(img)
ref = op.non_max_suppression(eng.last_prediction().cpu().numpy(), 0.25, 0.7, max_det=300)[0]  # the y of the detect() step
np.testing.assert_array_equal(det[0].cpu().numpy(), ref)
print("NMS rows bit-exact:", det[0].shape[0])
del eng
eng = RefineEngine(sd, batch=4, imgsz=1280, dtype=torch.bfloat16, nms_args=dict(conf_thres=0.25, iou_thres=0.7, max_det=300), input_u8=True)
eng.fill_inputs(torch.randint(0, 256, (4, 3, 1280, 1280), dtype=torch.uint8).cuda())
for _ in range(5):
    eng.step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    eng.step()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 20
print(f"bf16 1280^2 batch 4: {ms:.2f} ms/step, {4 / ms * 1e3:.0f} img/s per GPU")
