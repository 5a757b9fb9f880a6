#!/bin/bash
# Round 2, GPU session 15: attn_otm2 (two softmax threads per row) vs attn_otm; TMA-store A/B (warm); determinism probe; T3; bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=12 t s15_tests 1200 $PT tests
echo "== flow timing (FB=16 NT=4), default"; FB=16 NT=4 FCLS=none,gemm_tc,flash timeout 200 python tools/flow_only.py 2>&1 | tail -3
echo "== attn_otm (one thread per row)"; CBX_ATTN_F16=3 FB=16 NT=4 FCLS=none,flash timeout 200 python tools/flow_only.py 2>&1 | tail -1
echo "== attn_otm2, 1 CTA/SM"; CBX_ATTN_OCC=1 FB=16 NT=4 FCLS=none,flash timeout 200 python tools/flow_only.py 2>&1 | tail -1
echo "== gemm_tc without the TMA-store epilogue"; CBX_TMA_STORE=0 FB=16 NT=4 FCLS=none,gemm_tc timeout 200 python tools/flow_only.py 2>&1 | tail -1
echo "== determinism probe fp32 KV"; REPS=8 timeout 300 python tools/t3_determinism.py 2>&1 | tail -10
TCLS=none TAILN=1 t s15_t3 300 python tools/t3_only.py
TCLS=none TB=1 TSTEPS=150 TAILN=1 t s15_b1 200 python tools/t3_only.py
TAILN=3 t s15_bench 900 python bench.py --steps 1 --warmup 1 --no-extra --cpu-sample none
tail -n 1 gpurun_out/s15_bench.log > gpurun_out/s15_bench_line.json
grep -E "warmup|timed|profile" gpurun_out/s15_bench.log | cut -c1-300
