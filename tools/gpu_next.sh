#!/bin/bash
# First GPU run of the next round: validate the opt-in formats written at the end of round 1 and measure them.
#   gpurun --timeout 600 -- 'bash tools/gpu_next.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
CBX_EXPERIMENTAL=1 t exp_tests 300 python -m pytest tests/test_gpu_s3gen.py -q -m gpu -p no:cacheprovider --timeout 120 -k "fp16"
t flow_default 200 python tools/flow_breakdown.py
ATTN_PREC=fp16 t flow_attn16 200 python tools/flow_breakdown.py
ATTN_PREC=fp16 CFM_ACT=fp16 t flow_all16 200 python tools/flow_breakdown.py
