"""Ordered per-launch timeline of one eager inference step (batch 64, 640x640, bf16): every libyad entry point with its start offset and duration
(CUDA events; a queued spin kernel keeps the host ahead of the device).  Used to find what sits on the critical path of the graph-replayed step.
usage: python tools/step_timeline.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from yolo_ad_refine_b200 import ops, synth
from yolo_ad_refine_b200.engine import RefineEngine

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_step_timeline.json"
sd = synth.make_state_dict(seed=1)
eng = RefineEngine(sd, batch=64, imgsz=640, dtype=torch.bfloat16, nms_args=dict(conf_thres=0.25, iou_thres=0.7, max_det=300), input_u8=True, use_graph=False,
                   pipeline_nms=False)
rs = np.random.RandomState(100)
eng.img.copy_(torch.from_numpy(rs.randint(0, 256, (64, 3, 640, 640), dtype=np.uint8)))
for _ in range(3):
    eng._run()
torch.cuda.synchronize()
ops.PROFILE = {}
base = torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(80e6))
base.record()
eng._run()
torch.cuda.synchronize()
rows = []
for name, evs in ops.PROFILE.items():
    for s, e, m in evs:
        rows.append(dict(name=name, start_us=base.elapsed_time(s) * 1e3, us=s.elapsed_time(e) * 1e3, shape=(m or {}).get("shape", "")))
ops.PROFILE = None
rows.sort(key=lambda r: r["start_us"])
t0 = rows[0]["start_us"]
for r in rows:
    r["start_us"] -= t0
json.dump(rows, open(out, "w"), indent=0)
tot = sum(r["us"] for r in rows)
print(f"{len(rows)} calls, sum {tot:.0f} us, span {rows[-1]['start_us'] + rows[-1]['us']:.0f} us")
