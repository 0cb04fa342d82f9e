#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -x 2>&1 | tail -3
python tools/step_timeline.py gpurun_out/r2_step_timeline_h.json 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --train-batch 0 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
