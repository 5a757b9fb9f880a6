#!/bin/bash
# Round-end style validation: full GPU test suite (one process), T3 paged-attention bandwidth, smoke, default bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit=$rc"; tail -n 3 gpurun_out/$name.log | cut -c1-4000; return $rc; }
t final_pytest 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 300 || exit 1
TCLS=paged t t3_paged 200 python tools/t3_only.py
t final_smoke 300 python __graft_entry__.py smoke
export CBX_BENCH_WATCHDOG=900
t final_bench 600 python bench.py
