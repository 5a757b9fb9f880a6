#!/bin/bash
# Round 2, GPU session 21: attn_pp_kernel (two query tiles per CTA, anti-phase hand-off) -- correctness under pytest, then timing A/B.
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
CBX_ATTN_F16=5 timeout 600 $PT tests/test_gpu_s3gen.py tests/test_gpu_long.py tests/test_gpu_kernels.py tests/test_gpu_variants.py tests/test_gpu_e2e.py -x 2>&1 | tail -5 | cut -c1-300
for v in 3 5 6; do echo "== CBX_ATTN_F16=$v"; CBX_ATTN_F16=$v FB=16 NT=4 FCLS=none,all timeout 200 python tools/flow_only.py 2>&1 | tail -2 | cut -c1-200; done
timeout 300 $PT tests/test_gpu_kernels.py tests/test_gpu_long.py -k "paged or batch64 or eos" 2>&1 | tail -3
