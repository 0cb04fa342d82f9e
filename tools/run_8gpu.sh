#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --train-steps 3 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "default rc=$?"; tail -c 600 gpurun_out/r2_bench_${N}gpu.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --imgsz 1280 --batch 4 --steps 20 --warmup 3 --train-batch 0 --no-cpu-baseline > gpurun_out/r2_bench_cfg5_${N}gpu.json 2> gpurun_out/r2_bench_cfg5_${N}gpu.err; echo "cfg5 rc=$?"; head -c 600 gpurun_out/r2_bench_cfg5_${N}gpu.json; echo
