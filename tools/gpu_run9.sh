#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/gemm_anatomy.py > gpurun_out/gemm_anatomy.log 2>&1; echo "exit=$?"; cat gpurun_out/gemm_anatomy.log | tail -70
