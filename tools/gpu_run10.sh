#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
t kernels_tc 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 100 -k "not simt" -x
t s3gen 300 python -m pytest tests/test_gpu_s3gen.py -q -m gpu -p no:cacheprovider --timeout 120
t t3 300 python -m pytest tests/test_gpu_t3.py -q -m gpu -p no:cacheprovider --timeout 200
timeout 200 python tools/gemm_anatomy.py > gpurun_out/gemm_anatomy.log 2>&1; grep "gemm dbg" gpurun_out/gemm_anatomy.log; grep -A3 "N=256 K=1024" gpurun_out/gemm_anatomy.log | tail -3
t breakdown 300 python tools/flow_breakdown.py
