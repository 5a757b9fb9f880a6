#!/bin/bash
# Round 2, GPU session 8: paged-attention diagnostics (scratch poisoned with NaN; test order bf16 -> fp32), then the failing pytest subset.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== NaN-poisoned scratch"; PD_REPS=3 timeout 300 python tools/paged_diag.py 2>&1 | tail -14
echo "== 0x42-poisoned scratch"; PD_REPS=3 PD_FILL=66 timeout 300 python tools/paged_diag.py 2>&1 | tail -14
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --no-header tests/test_gpu_kernels.py -k paged 2>&1 | tail -8
