#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for nf in 2 3 4; do
timeout 600 python bench.py --steps 40 --warmup 4 --train-batch 0 --no-cpu-baseline --in-flight $nf > gpurun_out/r2_bench_ov.json 2> gpurun_out/r2_bench_ov.err; echo "in-flight $nf rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_ov.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
done
