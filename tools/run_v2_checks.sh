#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_v2.py tests/test_gpu_tc.py -q -x 2>&1 | tail -3
for shape in "64 64 3 1 80" "128 64 3 1 80" "64 64 1 1 80" "128 128 1 1 80" "48 64 1 1 160" "32 32 1 1 160" "64 32 3 1 80" "128 384 1 1 20"; do
  echo "== $shape"; python tools/conv_probe.py $shape 64 20 2 2>&1 | tail -1
done
echo "== 64 64 3 1 80 dbg2"; YAD_CONV2_DBG=2 python tools/conv_probe.py 64 64 3 1 80 64 20 2 2>&1 | tail -1
DEFORM=1 python tools/conv_probe.py 64 64 3 1 80 64 20 2 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/r2_bench_nodiv.json 2> gpurun_out/r2_bench_nodiv.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench_nodiv.json; echo
