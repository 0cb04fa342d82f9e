"""Summarises an `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <command>` launch list into per-kernel
shares (cold-cache, serialised launches: compare shares, not absolutes).  usage: python tools/launch_shares.py launches.csv out.json "<command>" """
import csv
import json
import re
import sys
from collections import defaultdict

src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "").strip()
    a = agg[name]
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
out = {"source": "ncu --metrics gpu__time_duration.sum --clock-control none --csv " + cmd + " (cold-cache, serialised launches: compare shares, not absolutes)",
       "launches_captured": sum(a[0] for a in agg.values()), "total_us": round(total, 1),
       "kernels": [{"kernel": k, "launches": a[0], "us_total": round(a[1], 1), "share": round(a[1] / total, 4)}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open(dst, "w"), indent=1)
for k in out["kernels"][:12]:
    print(f"{k['kernel'][:70]:70s} {k['launches']:5d} {k['share']:.3f}")
