#!/bin/bash
# ncu launch list + convolution DRAM traffic of the current build (run after the plain bench has exited 0); outputs under gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/r2b_bench_short.json 2> gpurun_out/r2b_bench_short.err || exit 1
CALLS=$(python -c "import json; print(json.loads(open('gpurun_out/r2b_bench_short.json').read().strip().splitlines()[-1])['roofline']['launches'])")
echo "yad_conv2d calls per step: $CALLS"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2b_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/r2b_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r2b_ncu_launches.csv gpurun_out/r2b_launch_shares.json "python bench.py --steps 2 --warmup 3 --train-batch 0 --no-cpu-baseline" | head -c 600; echo
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'conv_tma_kernel|conv2_kernel|conv3_kernel|dcn2_kernel|conv_tc_kernel|conv_small_kernel|conv_simt_kernel' --csv \
    --log-file gpurun_out/r2_conv_traffic.csv python bench.py --steps 1 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
python tools/conv_traffic.py gpurun_out/r2_conv_traffic.csv $CALLS gpurun_out/r2_conv_traffic.json
python tools/step_timeline.py gpurun_out/r2b_step_timeline.json
