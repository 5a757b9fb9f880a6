#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 5 gpurun_out/$name.log | cut -c1-600; }
t attn_tc 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 100 -k "tcgen05"
t kernels_tc 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 100 -k "not simt"
t s3gen 300 python -m pytest tests/test_gpu_s3gen.py -q -m gpu -p no:cacheprovider --timeout 120
t t3 300 python -m pytest tests/test_gpu_t3.py -q -m gpu -p no:cacheprovider --timeout 200
t breakdown 300 python tools/flow_breakdown.py
