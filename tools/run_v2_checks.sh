#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_v2.py tests/test_gpu_tc.py -q 2>&1 | tail -3
for shape in "64 64 3 1 80" "128 64 3 1 80" "128 128 1 1 80" "64 64 1 1 80" "192 128 1 1 80"; do
  timeout 120 python tools/conv_probe.py $shape 64 30 4 2>&1 | tail -1
done | tee gpurun_out/v2_probe_stg2.log
for std in 0.2 0.5 1.0; do
DEFORM=1 DEFORM_STD=$std timeout 120 python tools/conv_probe.py 64 64 3 1 80 64 10 2 2>&1 | tail -1 | sed "s/^/std$std /"
done
timeout 600 python bench.py --steps 20 --warmup 5 --train-batch 0 --no-cpu-baseline --profile-json gpurun_out/r2_step_profile_d.json > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench_d.json
echo done
