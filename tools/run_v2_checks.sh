#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DEFORM=1 DEFORM_STD=0.4 python tools/conv_probe.py 64 64 3 1 80 64 3 2 > gpurun_out/plain_dcn.log 2>&1 && \
DEFORM=1 DEFORM_STD=0.4 ncu --set full --clock-control none --import-source on -k regex:dcn2_kernel -s 3 -c 1 -o gpurun_out/r2_dcn2 -f python tools/conv_probe.py 64 64 3 1 80 64 3 2 > gpurun_out/ncu_dcn.log 2>&1
tail -2 gpurun_out/ncu_dcn.log
