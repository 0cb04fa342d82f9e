"""TEST INFRASTRUCTURE (uses the CPU oracle as the checker; lives under tests/ for that reason, run it by hand on the GPU box).
Debug / evidence tool (GPU box): one full training forward + backward through libyad.so against the CPU oracle (oracle.model.train_step_grads,
pinned to the live reference by tests/golden/train_step.npz): head outputs, loss, EVERY parameter gradient, BatchNorm buffer updates.
Prints the relative L2 error per parameter so that a wrong backward kernel shows up at the first parameter it touches.

    python tests/train_debug.py [case] [fp32|bf16] [json out]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, synth  # noqa: E402
from oracle import model as om  # noqa: E402
from yolo_ad_refine_b200.trainer import TrainEngine  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "b2_160"
    dtype = torch.float32 if (len(sys.argv) < 3 or sys.argv[2] == "fp32") else torch.bfloat16
    sd = synth.make_state_dict(seed=1)
    img, bi, cl, bb = cases.train_step_inputs(**cases.TRAIN_STEP_CASES[case])
    t0 = time.time()
    loss, items, grads, bn_upd, feats = om.train_step_grads(sd, torch.from_numpy(img), torch.from_numpy(bi), torch.from_numpy(cl), torch.from_numpy(bb))
    print(f"oracle: loss {loss.item():.6f} items {items.numpy()} ({time.time() - t0:.1f} s)")
    eng = TrainEngine(sd, dtype=dtype, conv_impl=1 if dtype == torch.float32 else 0)
    out4 = eng.forward_backward(torch.from_numpy(img).cuda(), torch.from_numpy(bi), torch.from_numpy(cl), torch.from_numpy(bb), keep=True)
    torch.cuda.synchronize()
    o4 = out4.cpu().numpy()
    print(f"libyad: loss {o4[3]:.6f} items {o4[:3]}")
    report = {"case": case, "dtype": str(dtype), "loss_ref": float(loss), "loss": float(o4[3]), "feats": [], "grads": {}, "bufs": {}}
    for i, (o, f) in enumerate(zip(eng.last["outs"], feats)):
        got = o.nchw().float().cpu()
        err = float((got - f).norm() / f.norm())
        report["feats"].append(err)
        print(f"  head output {i}: rel L2 err {err:.3e}")
    tp = eng.tp
    bad = 0
    for k in tp.keys:
        ref = grads.get(k)
        got = tp.g(k).cpu()
        if ref is None:
            mx = float(got.abs().max())
            if mx != 0.0:
                print(f"  {k}: reference has no gradient, got max {mx:.3e}")
            continue
        rn = float(ref.norm())
        err = float((got - ref).norm()) / (rn + 1e-12)
        report["grads"][k] = [err, rn]
        flag = err > (2e-3 if dtype == torch.float32 else 1e-1) and rn > 1e-6
        bad += flag
        if flag or "-v" in sys.argv:
            print(f"  {'BAD ' if flag else '    '}{k}: rel err {err:.3e} (|ref| {rn:.3e}, |got| {float(got.norm()):.3e})")
    print(f"{bad} of {len(report['grads'])} parameter gradients off")
    for k, v in bn_upd.items():
        got = tp.buf(k).cpu()
        err = float((got - v.reshape(-1)).abs().max() / (v.abs().max() + 1e-12))
        report["bufs"][k] = err
        if err > 1e-3:
            print(f"  BUF {k}: {err:.3e}")
    if len(sys.argv) > 3 and sys.argv[3].endswith(".json"):
        with open(sys.argv[3], "w") as f:
            json.dump(report, f)


if __name__ == "__main__":
    main()
