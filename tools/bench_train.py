"""Training-step micro-benchmark (GPU box): img/s of TrainEngine.step at a given batch / size, plus a per-entry-point CUDA-event profile.

    python tools/bench_train.py [batch] [size] [steps] [profile json]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from yolo_ad_refine_b200 import ops, synth  # noqa: E402
from yolo_ad_refine_b200.trainer import TrainEngine  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    sd = synth.make_state_dict(seed=1)
    img = torch.from_numpy(synth.make_images(batch, size, size, seed=5)).cuda()
    bi, cl, bb = [torch.from_numpy(a).cuda() for a in synth.make_targets(batch, seed=6, max_per_img=8)]
    eng = TrainEngine(sd, dtype=torch.bfloat16)
    for _ in range(2):
        out4 = eng.step(img, bi, cl, bb)
    torch.cuda.synchronize()
    print("warm loss", out4.cpu().numpy(), "peak mem GB", torch.cuda.max_memory_allocated() / 2**30)
    l0 = ops.LAUNCHES
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s.record()
    for _ in range(steps):
        eng.step(img, bi, cl, bb)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    print(f"batch {batch} size {size}: {ms:.2f} ms/step (wall {1e3 * (time.time() - t0) / steps:.2f}), {batch / ms * 1e3:.1f} img/s, "
          f"{(ops.LAUNCHES - l0) // steps} launches/step")
    if len(sys.argv) > 4:
        ops.PROFILE = {}
        eng.step(img, bi, cl, bb)
        torch.cuda.synchronize()
        prof = {}
        for name, evs in ops.PROFILE.items():
            tot = sum(a.elapsed_time(b) for a, b, _ in evs)
            prof[name] = dict(ms=tot, calls=len(evs))
            shapes = {}
            for a, b, meta in evs:
                if meta:
                    d = shapes.setdefault(meta["shape"], dict(ms=0.0, calls=0, flops=meta["flops"], bytes=meta.get("bytes", 0)))
                    d["ms"] += a.elapsed_time(b)
                    d["calls"] += 1
            if shapes:
                prof[name]["shapes"] = shapes
        ops.PROFILE = None
        tot = sum(v["ms"] for v in prof.values())
        print(f"profiled step: {tot:.2f} ms in kernels (event-bracketed)")
        for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:25]:
            print(f"  {name:28s} {v['ms']:8.2f} ms  {v['calls']:4d} calls")
        json.dump(prof, open(sys.argv[4], "w"), indent=1)


if __name__ == "__main__":
    main()
