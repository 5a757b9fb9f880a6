#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 4 gpurun_out/$name.log | cut -c1-3000; }
t pytest_all 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 300
export CBX_BENCH_WATCHDOG=600
t bench11 700 python bench.py --steps 1 --warmup 1
