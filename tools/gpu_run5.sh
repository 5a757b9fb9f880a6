#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 4 gpurun_out/$name.log | cut -c1-3000; }
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count())"
t s3gen 200 python -m pytest tests/test_gpu_s3gen.py -q -m gpu -p no:cacheprovider --timeout 120
export CBX_BENCH_WATCHDOG=800
t bench256 900 python bench.py --steps 1 --warmup 1
