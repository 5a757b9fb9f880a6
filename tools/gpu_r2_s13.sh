#!/bin/bash
# Round 2, GPU session 13: pipe micro-benchmark; plane-fed CFM convs (tests + timing A/B); batch64 test isolation (stream kernel on/off).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
timeout 120 tools/ubench/pipes.bin 2>&1 | tee gpurun_out/r2_pipes.txt
TAILN=6 t s13_b64_stream 300 $PT tests/test_gpu_long.py -k "batch64 or eos_retirement"
CBX_GEMM_STREAM=0 TAILN=6 t s13_b64_nostream 300 $PT tests/test_gpu_long.py -k "batch64 or eos_retirement"
TAILN=12 t s13_tests 1200 $PT tests --deselect tests/test_gpu_long.py::test_batch64_mixed_lengths_equals_single_runs
echo "== flow timing, plane-fed convs on / off (FB=16 NT=4)"
FB=16 NT=4 FCLS=none,gemm_tc timeout 200 python tools/flow_only.py 2>&1 | tail -2
CBX_CONV_PLANES=0 FB=16 NT=4 FCLS=none,gemm_tc timeout 200 python tools/flow_only.py 2>&1 | tail -2
