#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train_step.py -q 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --train-batch 128 --no-cpu-baseline > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_e.json').read().strip().split('\n')[-1])
print('inference', d['value'], d['ms_per_step'])
t=d['train']; print({k:v for k,v in t.items() if k not in ('roofline','profile_ms')})
print(t['profile_ms'])
PY
