#!/bin/bash
# Round 2, GPU session 17: paged fp32 repeatability with / without the proxy fence; triple-buffered S tiles in attn_otm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== paged fp32 repeatability, proxy fence on (default)"; PD_DT=fp32 PD_REPS=40 timeout 300 python tools/paged_determinism.py 2>&1 | tail -4
echo "== proxy fence off"; CBX_PB_NOFENCE=1 PD_DT=fp32 PD_REPS=40 timeout 300 python tools/paged_determinism.py 2>&1 | tail -4
echo "== flow timing (FB=16 NT=4)"; FB=16 NT=4 FCLS=none,all timeout 200 python tools/flow_only.py 2>&1 | tail -2
python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header tests/test_gpu_s3gen.py tests/test_gpu_long.py tests/test_gpu_kernels.py 2>&1 | tail -4
