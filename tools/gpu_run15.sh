#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 3 gpurun_out/$name.log | cut -c1-3000; }
t pytest_gpu 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 -k "not simt" -x
t breakdown 300 python tools/flow_breakdown.py
export CBX_BENCH_WATCHDOG=800
t bench256 900 python bench.py --steps 1 --warmup 1
