#!/bin/bash
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 100 -k "not simt" -x 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_s3gen.py -q -m gpu -p no:cacheprovider --timeout 100 -x 2>&1 | tail -2
echo "== dual"; timeout 300 python tools/flow_breakdown.py 2>&1 | tail -3
echo "== no dual"; CBX_TILE=1 timeout 300 python tools/flow_breakdown.py 2>&1 | tail -3
