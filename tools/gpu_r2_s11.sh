#!/bin/bash
# Round 2, GPU session 11: attn_f16_kernel variants (correctness + time), frame-level phase scan, decode tile sweep.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=12 t s11_tests_flow 900 $PT tests/test_gpu_s3gen.py tests/test_gpu_long.py tests/test_gpu_e2e.py tests/test_hift_drift.py tests/test_gpu_variants.py tests/test_gpu_turbo.py -x
for v in "2 2" "1 2" "0 2" "2 1" "1 1"; do set -- $v; echo "== CBX_ATTN_F16=$1 CBX_ATTN_OCC=$2"; CBX_ATTN_F16=$1 CBX_ATTN_OCC=$2 FB=16 NT=2 FCLS=none,flash timeout 200 python tools/flow_only.py 2>&1 | tail -2; done
CBX_ATTN_F16=1 TAILN=6 t s11_tests_f16_1 600 $PT tests/test_gpu_s3gen.py tests/test_gpu_long.py -k "cfm or mel or flow"
HB=24 TAILN=4 t s11_hift 300 python tools/hift_only.py
echo "== decode tile sweep (B=256, 120 steps): qkv_bn,qkv_dual,o_bn,o_split,gu_bn,gu_dual,down_bn,down_split"
for cfg in "64,1,64,2,128,1,64,4" "64,1,64,2,256,0,64,4" "128,1,64,2,128,1,64,4" "128,0,64,2,128,1,64,4" "64,1,64,4,128,1,64,4" "64,1,64,2,128,1,64,8" "64,1,64,2,128,0,64,4" "64,0,64,2,128,1,64,4"; do
  echo "-- $cfg"; CBX_DECODE_TILES=$cfg TCLS=none TB=256 TSTEPS=120 timeout 200 python tools/t3_only.py 2>&1 | tail -1 | cut -c1-120; done
echo "== decode kernel classes (B=256, 120 steps, no graph)"
TCLS=gemm_tc,paged TB=256 TSTEPS=120 timeout 300 python tools/t3_only.py 2>&1 | tail -2 | cut -c1-200
