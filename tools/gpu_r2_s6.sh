#!/bin/bash
# Round 2, GPU session 6: paged producer prefetch, GEMV norm fusion, ncu --set full captures of the top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-500; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=22 t s6_tests 900 $PT -s tests
TCLS=none t s6_t3only 200 python tools/t3_only.py
TCLS=paged t s6_t3only_paged 200 python tools/t3_only.py
TCLS=none TB=1 TSTEPS=150 t s6_b1 200 python tools/t3_only.py
NCU="ncu --set full --clock-control none --import-source on"
FB=8 NT=1 timeout 300 $NCU -k regex:attn_tc_kernel -s 20 -c 2 -o gpurun_out/r2_attn_tc python tools/flow_only.py > gpurun_out/s6_ncu_attn.log 2>&1; echo "ncu attn exit=$?"
FB=8 NT=1 timeout 300 $NCU -k regex:gemm_wres_kernel -s 30 -c 3 -o gpurun_out/r2_gemm_wres python tools/flow_only.py > gpurun_out/s6_ncu_wres.log 2>&1; echo "ncu wres exit=$?"
FB=8 NT=1 timeout 300 $NCU -k regex:gemm_tc_kernel -s 400 -c 4 -o gpurun_out/r2_gemm_tc_flow python tools/flow_only.py > gpurun_out/s6_ncu_gemm_flow.log 2>&1; echo "ncu gemm flow exit=$?"
HB=8 timeout 300 $NCU -k regex:hift_conv_kernel -s 20 -c 4 -o gpurun_out/r2_hift_conv python tools/hift_only.py > gpurun_out/s6_ncu_hift.log 2>&1; echo "ncu hift exit=$?"
CBX_DECODE_GRAPH=0 CBX_DECODE_PDL=0 TCLS=none TB=256 TSTEPS=24 timeout 400 $NCU -k regex:paged_bulk_kernel -s 1300 -c 2 -o gpurun_out/r2_paged_bulk python tools/t3_only.py > gpurun_out/s6_ncu_paged.log 2>&1; echo "ncu paged exit=$?"
CBX_DECODE_GRAPH=0 CBX_DECODE_PDL=0 TCLS=none TB=256 TSTEPS=24 timeout 400 $NCU -k regex:gemm_tc_kernel -s 5200 -c 5 -o gpurun_out/r2_gemm_tc_decode python tools/t3_only.py > gpurun_out/s6_ncu_gemm_decode.log 2>&1; echo "ncu gemm decode exit=$?"
ls -la gpurun_out/*.ncu-rep
TAILN=3 t s6_bench 900 python bench.py --steps 1 --warmup 1 --no-extra --cpu-sample none
tail -n 1 gpurun_out/s6_bench.log | cut -c1-700
