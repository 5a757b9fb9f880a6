#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 70 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_t3.py -x -q -s -m gpu -p no:cacheprovider --timeout 60 > gpurun_out/last_t3.log 2>&1; echo "t3+e2e exit=$?"; grep -E "decode .* steps|passed|failed|Error" gpurun_out/last_t3.log | head -8
