#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_full.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2_bench_full.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; head -c 700 gpurun_out/r2_bench_ref.json; echo
