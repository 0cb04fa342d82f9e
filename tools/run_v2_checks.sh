#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/r2_bench_c3s2.json 2> gpurun_out/r2_bench_c3s2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_c3s2.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['roofline']['share_of_step'])
PY
done
