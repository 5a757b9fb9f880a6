#!/bin/bash
# Round 2, GPU session 12: attn_otm_kernel (O in TMEM) correctness + time; decode projection micro-benchmark (stream kernel vs tile kernel).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=15 t s12_tests 1200 $PT tests -x
for v in "3 2" "3 1" "2 2"; do set -- $v; echo "== CBX_ATTN_F16=$1 CBX_ATTN_OCC=$2"; CBX_ATTN_F16=$1 CBX_ATTN_OCC=$2 FB=16 NT=2 FCLS=none,flash timeout 200 python tools/flow_only.py 2>&1 | tail -2; done
echo "== decode projections, streaming kernel"; timeout 300 python tools/decode_gemm_bench.py 2>&1 | tail -20
echo "== decode projections, tile kernel (CBX_GEMM_STREAM=0)"; CBX_GEMM_STREAM=0 timeout 300 python tools/decode_gemm_bench.py 2>&1 | tail -20
TCLS=none TAILN=2 t s12_t3 300 python tools/t3_only.py
CBX_GEMM_STREAM=0 TCLS=none TAILN=2 t s12_t3_nostream 300 python tools/t3_only.py
