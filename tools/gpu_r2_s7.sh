#!/bin/bash
# Round 2, GPU session 7 (re-entry): state check -- full GPU tests, T3/flow/HiFT stage timings, short bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=25 t s7_tests 1200 $PT tests --durations=15
TAILN=3 t s7_bench 900 python bench.py --steps 1 --warmup 1 --no-extra --cpu-sample none
tail -n 1 gpurun_out/s7_bench.log > gpurun_out/s7_bench_line.json
