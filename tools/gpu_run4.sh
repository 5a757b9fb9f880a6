#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 8 gpurun_out/$name.log | cut -c1-400; }
t s3gen 300 python -m pytest tests/test_gpu_s3gen.py -q -m gpu -p no:cacheprovider --timeout 120
t e2e 300 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout 200
export CBX_TRACE=1 CBX_BENCH_WATCHDOG=280
t bench4 300 python bench.py --batch 4 --steps 1 --warmup 1 --budget-max 100
t bench16 300 python bench.py --batch 16 --steps 1 --warmup 1 --budget-max 300
