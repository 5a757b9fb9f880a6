#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 80 python -m pytest tests/test_gpu_t3.py -x -q -m gpu -p no:cacheprovider --timeout 60 > gpurun_out/last_t3.log 2>&1; echo "t3 exit=$?"; tail -n 3 gpurun_out/last_t3.log | cut -c1-300
timeout 80 python -m pytest tests/test_gpu_turbo.py -q -m gpu -p no:cacheprovider --timeout 60 > gpurun_out/last_turbo.log 2>&1; echo "turbo exit=$?"; tail -n 12 gpurun_out/last_turbo.log | cut -c1-300
