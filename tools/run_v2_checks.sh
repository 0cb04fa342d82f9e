#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_v2.py tests/test_gpu_tc.py -q > gpurun_out/v2_pytest.log 2>&1
echo "v2 pytest rc=$?"; tail -4 gpurun_out/v2_pytest.log
for shape in "128 128 1 1 80" "64 64 3 1 80" "128 64 3 1 80" "64 64 3 2 160" "128 128 3 2 80" "128 256 3 2 40"; do
  timeout 120 python tools/conv_probe.py $shape 64 20 2 2>&1 | tail -1
done | tee gpurun_out/v2_probe.log
DEFORM=1 timeout 120 python tools/conv_probe.py 64 64 3 1 80 64 10 2 2>&1 | tail -1 | tee -a gpurun_out/v2_probe.log
timeout 600 python bench.py --steps 20 --warmup 5 --train-batch 128 --no-cpu-baseline --profile-json gpurun_out/r2_step_profile_b.json > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
echo "bench rc=$?"; head -c 400 gpurun_out/r2_bench_b.json
timeout 200 python tools/conv_probe.py 64 64 3 1 80 64 3 4 > gpurun_out/probe_a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 2 -c 1 -o gpurun_out/r2_v2b_64_64_3 python tools/conv_probe.py 64 64 3 1 80 64 3 4 > gpurun_out/ncu_a.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches_b.csv python bench.py --steps 2 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/ncu_bench_b.log 2>&1
echo done
