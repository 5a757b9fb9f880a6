#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mode in 0 1 2; do echo "== EPI_MODE=$mode"; CBX_EPI_MODE=$mode timeout 200 python tools/gemm_anatomy.py 2>&1 | grep -E "gemm dbg|epilogue" | head -4; done
echo "== carveout 70"; CBX_CARVEOUT=70 timeout 200 python tools/gemm_anatomy.py 2>&1 | grep -E "gemm dbg|epilogue" | head -4
