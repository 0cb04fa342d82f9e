#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/plain_final.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_ncu_launches.csv python bench.py --steps 2 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/launch_shares.py gpurun_out/r2_ncu_launches.csv gpurun_out/r2_launch_shares.json 2>&1 | tail -3
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv2_kernel|conv3_kernel|conv_tma_kernel|conv_tc_kernel|conv_small_kernel|dcn2_kernel" --csv --log-file gpurun_out/r2_conv_traffic.csv python bench.py --steps 1 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
python tools/conv_traffic.py gpurun_out/r2_conv_traffic.csv 146 gpurun_out/r2_conv_traffic.json | cut -c1-700
