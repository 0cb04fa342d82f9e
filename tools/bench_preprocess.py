"""Micro-benchmark of the device pre-processing path: 64 x 1280x720 BGR frames -> (64, 3, 640, 640) uint8 RGB
(`yad_letterbox`) and scale_boxes of a (64, 300, 6) NMS output (`yad_scale_boxes`).  Device-timed with CUDA events; one JSON line:
  value      images/s of the letterbox kernel alone, sources resident in HBM
  e2e        images/s through DevicePreprocessor.__call__ with the frames in pinned HOST memory (H2D of the raw frames inside the timed region)
  roofline   algorithmic bytes (every source pixel the taps address read once + 3 bytes written per output pixel) / kernel time vs the measured HBM peak
  cpu_baseline   the reference's own host path for the same frames: cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder + stack / BGR->RGB / CHW
                 (LetterBox.__call__, data/augment.py:1588-1594; engine/predictor.py:127-129), timed on the box's host cores
usage: python tools/bench_preprocess.py [--reps 50] [--batch 64] [--src 720x1280]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.preprocess import DevicePreprocessor, letterbox_params

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--src", default="720x1280")
ap.add_argument("--imgsz", type=int, default=640)
args = ap.parse_args()
dev = "cuda"
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
sh, sw = (int(v) for v in args.src.split("x"))
B, S = args.batch, args.imgsz
rs = np.random.RandomState(0)
frames = [rs.randint(0, 256, (sh, sw, 3), dtype=np.uint8) for _ in range(min(B, 8))]
frames = [frames[i % len(frames)] for i in range(B)]

pre = DevicePreprocessor(S, device=dev)
out = torch.empty((B, 3, S, S), dtype=torch.uint8, device=dev)
pb = pre(frames, out=out)
desc = pb.desc
torch.cuda.synchronize()


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


# the device buffer `desc` points into belongs to staging set 0; keep `pre` idle while the kernel-only loop runs
ms_kernel = timed(lambda: ops.letterbox(desc, B, out, S, S, 114, True), args.reps)
nw, nh, *_ = letterbox_params((sh, sw), S)
alg_bytes = B * (min(sh * sw, 4 * nw * nh) * 3 + 3 * S * S)  # source pixels the 2 x 2 taps address (all of them up to a 2x reduction) + the output
shapes = [f.shape[:2] for f in frames]
for _ in range(2):  # fill both pinned staging sets once: the frames count as "decoded into pinned memory" (no host-side copy in the timed loop)
    vs = pre.staging_views(shapes)
    for v, f in zip(vs, frames):
        v[...] = f
    pre(vs, out=out)


def e2e():
    pre(pre.staging_views(shapes), out=out)  # waits for the set's previous use, one H2D of the raw frames + descriptor table, one kernel


ms_e2e = timed(e2e, args.reps)
det = torch.rand(B, 300, 6, device=dev) * 640
cnt = torch.full((B,), 300, dtype=torch.int32, device=dev)
ms_scale = timed(lambda: ops.scale_boxes(det, cnt, desc), args.reps)

cpu = None
try:
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    nw, nh, top, bottom, left, right = letterbox_params((sh, sw), S)

    def host():
        lb = [cv2.copyMakeBorder(cv2.resize(f, (nw, nh), interpolation=cv2.INTER_LINEAR), top, bottom, left, right, cv2.BORDER_CONSTANT,
                                 value=(114, 114, 114)) for f in frames]
        return np.ascontiguousarray(np.stack(lb)[..., ::-1].transpose((0, 3, 1, 2)))
    ref = host()
    same = bool(np.array_equal(ref, out.cpu().numpy()))
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 5.0:
        host()
        n += 1
    cpu = {"value": B * n / (time.perf_counter() - t0), "unit": "img/s", "cores": cv2.getNumThreads(), "kind": "reference",
           "sample": f"{n} batches of {B} frames {sh}x{sw}: cv2 {cv2.__version__} resize + copyMakeBorder + stack/transposes (the calls LetterBox makes)",
           "device_bytes_equal_host_bytes": same}
except ImportError:
    pass

print(json.dumps({
    "metric": "img/s device LetterBox + BGR->RGB + CHW", "value": B / ms_kernel * 1e3, "unit": "img/s", "ms_per_batch": ms_kernel, "dtype": "u8",
    "config": {"workload": f"{B} frames {sh}x{sw}x3 uint8 -> ({B},3,{S},{S}) uint8, resized region {nw}x{nh}"},
    "e2e": {"value": B / ms_e2e * 1e3, "unit": "img/s", "ms_per_batch": ms_e2e, "h2d_bytes_per_step": B * sh * sw * 3 + B * 48, "d2h_bytes_per_step": 0},
    "roofline": {"kernel": "letterbox_kernel", "bound": "hbm", "achieved": alg_bytes / ms_kernel / 1e6, "peak": peak, "unit": "GB/s",
                 "frac": alg_bytes / ms_kernel / 1e6 / peak, "algorithmic_bytes_per_launch": alg_bytes, "traffic": None},
    "scale_boxes": {"ms": ms_scale, "rows": B * 300},
    "cpu_baseline": cpu}))
