#!/bin/bash
# Round 2, GPU session 10: register-direct GEMM epilogue -- tests, stage timings (A/B against CBX_EPI_DIRECT=0), short bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=15 t s10_tests 1200 $PT tests -x
TCLS=none TAILN=2 t s10_t3 300 python tools/t3_only.py
CBX_EPI_DIRECT=0 TCLS=none TAILN=2 t s10_t3_old 300 python tools/t3_only.py
TCLS=none TB=1 TSTEPS=150 TAILN=2 t s10_b1 200 python tools/t3_only.py
TAILN=3 t s10_bench 900 python bench.py --steps 1 --warmup 1 --no-extra --cpu-sample none
tail -n 1 gpurun_out/s10_bench.log > gpurun_out/s10_bench_line.json
grep -E "warmup|timed|profile" gpurun_out/s10_bench.log | cut -c1-300
