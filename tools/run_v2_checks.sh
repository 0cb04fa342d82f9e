#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_c3.py tests/test_gpu_conv_v2.py -q -x 2>&1 | tail -3
for shape in "8 16 3 1 160" "16 32 3 1 80"; do
  echo "== $shape plain"; python tools/conv_probe.py $shape 64 20 6 2>&1 | tail -1
  echo "== $shape add"; ADD=1 python tools/conv_probe.py $shape 64 20 6 2>&1 | tail -1
  echo "== $shape slice"; SLICE=1 python tools/conv_probe.py $shape 64 20 6 2>&1 | tail -1
  echo "== $shape add+slice"; ADD=1 SLICE=1 python tools/conv_probe.py $shape 64 20 6 2>&1 | tail -1
done
