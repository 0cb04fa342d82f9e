#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_post.py -q -x 2>&1 | tail -15
timeout 600 python bench.py --workload nms_micro --steps 20 --warmup 3 > gpurun_out/r2_bench_nms_micro3.json 2> gpurun_out/r2_bench_nms_micro3.err; echo "nms_micro rc=$?"; head -c 300 gpurun_out/r2_bench_nms_micro3.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/r2_bench_nms3.json 2> gpurun_out/r2_bench_nms3.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench_nms3.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|decode" -c 14 --csv --log-file gpurun_out/r2_nms_launches2.csv python bench.py --steps 1 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/ncu_nms.log 2>&1
tail -7 gpurun_out/r2_nms_launches2.csv | cut -d, -f5,15- 
