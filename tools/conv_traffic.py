"""profiles/r2_conv_traffic.json from an ncu CSV of one bench step: DRAM bytes (read + written) per launch of the convolution kernels behind yad_conv2d.
usage (GPU box):  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv|dcn' --csv \
                      --log-file gpurun_out/r2_conv_traffic.csv python bench.py --steps 1 --warmup 3 --train-batch 0 --no-cpu-baseline
                  python tools/conv_traffic.py gpurun_out/r2_conv_traffic.csv LAUNCHES_PER_STEP profiles/r2_conv_traffic.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ik, im, iv, iu = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
per = {}
for r in rows[hdr + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6}.get(u, 1)
    per.setdefault(int(r[0]), {"kernel": r[ik]})[r[im]] = v * mult
ids = sorted(per)
calls = int(sys.argv[2])  # yad_conv2d entry-point calls per step (bench.py roofline.launches); a transposed convolution is 4 kernel launches
seq = [per[i]["kernel"] for i in ids]
try:
    n_step = next(p for p in range(calls, len(seq) // 2 + 1) if seq[-p:] == seq[-2 * p:-p])  # period of the kernel sequence = kernel launches per pass
    last = ids[-n_step:]  # the launches of the last (post-warm-up) pass
except StopIteration:
    # two batches in flight: ncu serialises the kernels of both forward streams in an arbitrary interleaving, so the tail has no clean period; the
    # two eager warm-up passes at the start of the run are sequential -- take the second one (ncu flushes the caches before every kernel anyway)
    n_step = next(p for p in range(calls, len(seq) // 2 + 1) if seq[:p] == seq[p:2 * p])
    last = ids[n_step:2 * n_step]
rd = sum(per[i].get("dram__bytes_read.sum", 0) for i in last)
wr = sum(per[i].get("dram__bytes_write.sum", 0) for i in last)
names = {}
for i in last:
    k = per[i]["kernel"].split("(")[0][-40:]
    names[k] = names.get(k, 0) + 1
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on the convolution kernels, last pass of bench.py (batch 64, 640x640, bf16)",
       "launches": calls, "kernel_launches": len(last), "dram_bytes_read": rd, "dram_bytes_written": wr, "dram_bytes_per_launch": (rd + wr) / max(1, calls),
       "kernels": names}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out))
