#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_c3.py tests/test_gpu_conv_v2.py -q -x 2>&1 | tail -3
python tools/step_timeline.py gpurun_out/r2_step_timeline_c3.json > gpurun_out/r2_step_timeline_c3.log 2>&1; tail -2 gpurun_out/r2_step_timeline_c3.log
timeout 600 python bench.py --steps 20 --warmup 3 --train-steps 3 --no-cpu-baseline > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_c3.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['train']['value'], d['train']['ms_per_step'])
PY
