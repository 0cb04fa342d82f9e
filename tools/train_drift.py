"""Evidence tool (GPU box): bf16 training step against the fp32 training step of the same library, layer by layer -- activations L[i] of the
forward and their gradients -- to tell rounding drift (smooth growth with depth) from a wrong kernel (a jump at one layer).

    python tools/train_drift.py [batch] [size] [json out]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from yolo_ad_refine_b200 import synth  # noqa: E402
from yolo_ad_refine_b200.trainer import TrainEngine  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 320
    sd = synth.make_state_dict(seed=1)
    img = torch.from_numpy(synth.make_images(batch, size, size, seed=5)).cuda()
    bi, cl, bb = [torch.from_numpy(a) for a in synth.make_targets(batch, seed=6, max_per_img=8)]
    res = {}
    for name, dtype, impl in (("fp32", torch.float32, 1), ("bf16", torch.bfloat16, 0)):
        eng = TrainEngine(sd, dtype=dtype, conv_impl=impl)
        out4 = eng.forward_backward(img, bi, cl, bb, keep=True)
        torch.cuda.synchronize()
        g = eng.last["graph"]
        res[name] = dict(out4=out4.cpu(), L={i: a.nchw().float().cpu() for i, a in eng.last["layers"].items()},
                         dL={i: g.grad(a).nchw().float().cpu() for i, a in eng.last["layers"].items()},
                         outs=[o.nchw().float().cpu() for o in eng.last["outs"]],
                         grads={k: eng.tp.g(k).cpu().clone() for k in eng.tp.keys},
                         fg=eng.last["aux"]["fg_mask"].cpu())
    a, b = res["fp32"], res["bf16"]
    print("loss fp32", a["out4"].numpy(), "bf16", b["out4"].numpy())
    print("fg anchors fp32", int(a["fg"].sum()), "bf16", int(b["fg"].sum()), "differing", int((a["fg"] != b["fg"]).sum()))
    rep = {"layers": {}, "dlayers": {}}
    rel = lambda x, y: float((x - y).norm() / (y.norm() + 1e-20))  # noqa: E731
    for i in sorted(a["L"]):
        rep["layers"][i] = rel(b["L"][i], a["L"][i])
        rep["dlayers"][i] = rel(b["dL"][i], a["dL"][i])
        print(f"  L[{i:2d}] act err {rep['layers'][i]:.3e}   grad err {rep['dlayers'][i]:.3e}")
    for i, (x, y) in enumerate(zip(b["outs"], a["outs"])):
        print(f"  head out {i}: {rel(x, y):.3e}")
    errs = sorted(((rel(b["grads"][k], a["grads"][k]), k) for k in a["grads"] if float(a["grads"][k].norm()) > 1e-4), reverse=True)
    print("worst parameter gradients:", [(f"{e:.2e}", k) for e, k in errs[:8]])
    print("median parameter-gradient err", errs[len(errs) // 2][0])
    rep["param_median"] = errs[len(errs) // 2][0]
    if len(sys.argv) > 3:
        json.dump(rep, open(sys.argv[3], "w"))


if __name__ == "__main__":
    main()
