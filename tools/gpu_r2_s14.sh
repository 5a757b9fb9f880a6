#!/bin/bash
# Round 2, GPU session 14: TMA-store epilogue of the weight-resident GEMM (tests + timing), fp32-KV determinism probe, ncu of attn_otm_kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=12 t s14_tests 1200 $PT tests
echo "== flow timing (FB=16 NT=4)"
FB=16 NT=4 FCLS=none,gemm_tc,flash timeout 200 python tools/flow_only.py 2>&1 | tail -3
echo "== same, gemm_tc without the TMA-store epilogue"; CBX_TMA_STORE=0 FB=16 NT=4 FCLS=gemm_tc timeout 200 python tools/flow_only.py 2>&1 | tail -1
echo "== same, attention kernel of session 12 (CBX_ATTN_F16=2)"; CBX_ATTN_F16=2 FB=16 NT=4 FCLS=flash timeout 200 python tools/flow_only.py 2>&1 | tail -1
HB=24 timeout 200 python tools/hift_only.py 2>&1 | tail -2
echo "== determinism probe fp32 KV"; REPS=8 timeout 300 python tools/t3_determinism.py 2>&1 | tail -12
NCU="ncu --set full --clock-control none --import-source on"
FB=8 NT=1 timeout 400 $NCU -k regex:attn_otm_kernel -s 20 -c 1 -o gpurun_out/r2_attn_otm python tools/flow_only.py > gpurun_out/s14_ncu_attn.log 2>&1; echo "ncu attn exit=$?"
FB=8 NT=1 timeout 400 $NCU -k regex:gemm_wres_kernel -s 30 -c 2 -o gpurun_out/r2_gemm_wres2 python tools/flow_only.py > gpurun_out/s14_ncu_wres.log 2>&1; echo "ncu wres exit=$?"
ls -la gpurun_out/*.ncu-rep
