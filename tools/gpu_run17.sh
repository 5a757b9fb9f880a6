#!/bin/bash
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 100 -k "tcgen05 or dual or bn256" 2>&1 | tail -3
timeout 200 python -m pytest tests/test_gpu_s3gen.py tests/test_gpu_variants.py -q -m gpu -p no:cacheprovider --timeout 150 2>&1 | tail -3
echo "== attn v2"; timeout 300 python tools/flow_breakdown.py 2>&1 | tail -3
echo "== attn v1"; CBX_ATTN_TC=1 timeout 300 python tools/flow_breakdown.py 2>&1 | tail -2
