#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for hs in 1 2 4; do
YAD_HEAD_SPLIT=$hs timeout 600 python bench.py --steps 20 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/r2_bench_hs.json 2> gpurun_out/r2_bench_hs.err; echo "head split $hs rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_hs.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['launches_per_step'])
PY
done
