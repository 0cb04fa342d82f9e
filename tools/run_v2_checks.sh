#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_conv_v2.py tests/test_gpu_tc.py tests/test_gpu_backward_ops.py tests/test_gpu_train_step.py -q -x 2>&1 | tail -3
for shape in "16 32 3 2 320" "64 64 3 2 160" "128 128 3 2 80" "128 128 3 1 40" "128 128 3 1 80"; do
  echo "== $shape"; python tools/conv_probe.py $shape 64 20 2 2>&1 | tail -1
done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_nodiv2.json 2> gpurun_out/r2_bench_nodiv2.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench_nodiv2.json; echo
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_nodiv2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','train') if k in d})
PY
