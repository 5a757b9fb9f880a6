#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/gemm_anatomy.py 2>&1 | grep -E "gemm dbg|epilogue"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 100 -k "not simt" -x 2>&1 | tail -2
timeout 300 python tools/flow_breakdown.py 2>&1 | tail -3
