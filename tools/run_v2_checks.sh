#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_post.py tests/test_gpu_model.py -q -x 2>&1 | tail -5
timeout 600 python bench.py --workload nms_micro --steps 20 --warmup 3 > gpurun_out/r2_bench_nms_micro4.json 2> gpurun_out/r2_bench_nms_micro4.err; echo "nms_micro rc=$?"; head -c 250 gpurun_out/r2_bench_nms_micro4.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench_a.json; echo
python tools/step_timeline.py > gpurun_out/r2_step_timeline2.log 2>&1; tail -3 gpurun_out/r2_step_timeline2.log
