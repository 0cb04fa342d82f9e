"""offset statistics of the head's deformable convolutions in the benchmark model (random-init synthetic weights), batch 8"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yolo_ad_refine_b200 import ops, synth, functional as Fn
from yolo_ad_refine_b200.weights import prepare
sd = synth.make_state_dict(seed=1)
ctx = Fn.Ctx(prepare(sd, torch.bfloat16, torch.device("cuda")), 0)
img = torch.from_numpy(np.random.RandomState(100).randint(0, 256, (8, 3, 640, 640), dtype=np.uint8)).cuda()
seen = []
orig = ops.conv2d
def spy(x, w, y, **kw):
    if kw.get("mode") == ops.CONV_DEFORM:
        om = kw["offmask"]
        t = om.buf.float().reshape(-1, om.ld)[:, :18]
        seen.append((x.h, float(t.abs().mean()), float(t.abs().quantile(0.9)), float((t.abs() > 2).float().mean())))
    return orig(x, w, y, **kw)
ops.conv2d = spy
Fn.ops.conv2d = spy
Fn.forward_model(ctx, img)
torch.cuda.synchronize()
for s in seen:
    print("map %d: |offset| mean %.2f px, p90 %.2f px, share > 2 px %.3f" % s)
