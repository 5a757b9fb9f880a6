#!/bin/bash
# Round-end style validation: full GPU test suite (one process), smoke, default bench, ncu evidence.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 3 gpurun_out/$name.log | cut -c1-4000; }
t final_pytest 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 300
t final_smoke 300 python __graft_entry__.py smoke
export CBX_BENCH_WATCHDOG=900
t final_bench 1000 python bench.py
