#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_c3.py -q -x 2>&1 | tail -6
echo "== 16 32 3 2 320"; timeout 120 python tools/conv_probe.py 16 32 3 2 320 64 20 2,6 2>&1 | tail -2
