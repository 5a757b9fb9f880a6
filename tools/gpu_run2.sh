#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -k "not simt" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py --batch 16 --steps 1 --warmup 1 > gpurun_out/bench_b16.log 2>&1; echo "bench16 exit=$?"; tail -n 2 gpurun_out/bench_b16.log
timeout 1500 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_b256.log 2>&1; echo "bench256 exit=$?"; tail -n 2 gpurun_out/bench_b256.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_small.csv python tools/profile_small.py > gpurun_out/profile_small.log 2>&1; echo "ncu small exit=$?"; tail -n 2 gpurun_out/profile_small.log
