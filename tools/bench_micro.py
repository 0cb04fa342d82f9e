"""Micro-benchmarks of the non-flagship configs in BASELINE.json (device-timed with CUDA events, results as JSON lines):
  config 3: fused DFL decode + batched NMS, batch 256 x 8400 anchors x 80 classes, synthetic logits
  config 4 (loss half): TaskAlignedAssigner + v8DetectionLoss forward and gradients w.r.t. the head outputs, batch 128 at 640^2
usage: python tools/bench_micro.py [--reps 20]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from yolo_ad_refine_b200 import ops, synth
from yolo_ad_refine_b200.loss import detection_loss_raw, preprocess_targets
from yolo_ad_refine_b200.postprocess import nms_raw
from yolo_ad_refine_b200.tal import make_anchors

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = "cuda"
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ---- config 3
B, N = 256, 8400
raw = torch.from_numpy(synth.make_head_logits(B, N, seed=0)).to(dev)
for name, dt_ in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
    lv = [raw[:, :, :6400].reshape(B, 144, 80, 80).permute(0, 2, 3, 1).contiguous().to(dt_),
          raw[:, :, 6400:8000].reshape(B, 144, 40, 40).permute(0, 2, 3, 1).contiguous().to(dt_),
          raw[:, :, 8000:].reshape(B, 144, 20, 20).permute(0, 2, 3, 1).contiguous().to(dt_)]
    levels = [ops.Act(t) for t in lv]
    y = torch.empty(B, 84, N, device=dev)
    proj = torch.arange(16, dtype=torch.float32, device=dev)

    def step():
        ops.decode(levels, (8, 16, 32), 80, 16, proj, y)
        return nms_raw(y, 0.25, 0.7, max_det=300)

    ms = timed(step, args.reps)
    ms_dec = timed(lambda: ops.decode(levels, (8, 16, 32), 80, 16, proj, y), args.reps)
    in_bytes = 144 * N * B * (4 if dt_ == torch.float32 else 2)
    alg = in_bytes + 300 * 6 * 4 * B
    print(json.dumps({"bench": "decode+nms", "config": f"B={B} N={N} nc=80 head logits {name}, conf .25 iou .7 max_det 300", "ms": ms, "decode_ms": ms_dec,
                      "img_per_s": B / ms * 1e3, "algorithmic_GB": alg / 1e9, "achieved_GBps": alg / ms / 1e6, "hbm_peak_GBps": peak,
                      "frac": alg / ms / 1e6 / peak, "decode_only_GBps": (in_bytes + 84 * N * B * 4) / ms_dec / 1e6}))

# ---- config 4, loss half
B = 128
rs = np.random.RandomState(5)
feats = []
for s in (80, 40, 20):
    f = rs.standard_normal((B, 144, s, s)).astype(np.float32)
    f[:, :64] *= 1.5
    f[:, 64:] = -3.0 + 1.5 * f[:, 64:]
    feats.append(torch.from_numpy(f).to(dev))
bi, cl, bb = synth.make_targets(B, seed=3, max_per_img=32, empty_images=(0,))
cat = torch.cat([f.reshape(B, 144, -1) for f in feats], 2)
pd, ps = cat[:, :64].permute(0, 2, 1).contiguous(), cat[:, 64:].permute(0, 2, 1).contiguous()
anc, st = make_anchors([(80, 80), (40, 40), (20, 20)], (8, 16, 32))
anc, st = anc.to(dev), st.to(dev)
gl, gb, mg = preprocess_targets(torch.from_numpy(bi), torch.from_numpy(cl), torch.from_numpy(bb), B, (640, 640), dev)
ms = timed(lambda: detection_loss_raw(pd, ps, anc, st, gl, gb, mg), max(3, args.reps // 2))
alg = B * 8400 * (80 + 64) * 4 * 3  # read logits twice (decode, loss) + write their gradients
print(json.dumps({"bench": "tal+loss fwd+grad", "config": f"B={B} N=8400 nc=80 n_max={gl.shape[1]} fp32 logits", "ms": ms, "img_per_s": B / ms * 1e3,
                  "algorithmic_GB": alg / 1e9, "achieved_GBps": alg / ms / 1e6, "hbm_peak_GBps": peak, "frac": alg / ms / 1e6 / peak}))
