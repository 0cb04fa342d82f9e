#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for shape in "64 64 3 1 80" "128 64 3 1 80" "64 32 3 1 80" "32 64 3 1 80" "64 64 3 1 40" "64 64 1 1 80" "128 64 1 1 80" "64 80 1 1 80" "64 32 1 1 80"; do
  for ew in 8 16; do
  YAD_CONV2_EW=$ew timeout 120 python tools/conv_probe.py $shape 64 30 4 2>&1 | tail -1 | sed "s/^/ew$ew /"
  done
done | tee gpurun_out/v2_probe_ew.log
echo done
