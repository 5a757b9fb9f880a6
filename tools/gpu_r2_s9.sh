#!/bin/bash
# Round 2, GPU session 9: stage-release A/B under pytest, full GPU tests, per-stage ncu launch lists, ncu --set full captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=12 t s9_tests 1200 $PT tests
CBX_PB_EARLY=1 TAILN=6 t s9_tests_early 600 $PT tests/test_gpu_kernels.py tests/test_gpu_long.py
LL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
FB=16 NT=2 timeout 600 $LL --log-file gpurun_out/r2_launches_flow.csv python tools/flow_only.py > gpurun_out/s9_ll_flow.log 2>&1; echo "ll flow exit=$?"
HB=8 timeout 600 $LL --log-file gpurun_out/r2_launches_hift.csv python tools/hift_only.py > gpurun_out/s9_ll_hift.log 2>&1; echo "ll hift exit=$?"
CBX_DECODE_GRAPH=0 CBX_DECODE_PDL=0 TCLS=paged TB=256 TSTEPS=6 timeout 900 $LL --log-file gpurun_out/r2_launches_t3.csv python tools/t3_only.py > gpurun_out/s9_ll_t3.log 2>&1; echo "ll t3 exit=$?"
for f in flow hift t3; do python tools/summarize_launches.py gpurun_out/r2_launches_$f.csv > gpurun_out/r2_launches_$f.txt 2>&1; head -30 gpurun_out/r2_launches_$f.txt; done
NCU="ncu --set full --clock-control none --import-source on"
cap() { local name=$1 kn=$2; shift 2; timeout 400 $NCU "$@" -o gpurun_out/$name > gpurun_out/s9_ncu_$name.log 2>&1; echo "ncu $name exit=$?";
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.txt $kn "" > /dev/null 2>&1; }
FB=8 NT=1 cap r2_attn_tc attn_tc_kernel -k regex:attn_tc_kernel -s 20 -c 2 python tools/flow_only.py
FB=8 NT=1 cap r2_gemm_wres gemm_wres_kernel -k regex:gemm_wres_kernel -s 30 -c 3 python tools/flow_only.py
FB=8 NT=1 cap r2_gemm_tc_flow gemm_tc_kernel -k regex:gemm_tc_kernel -s 400 -c 4 python tools/flow_only.py
FB=8 NT=1 cap r2_flash_enc flash_attn_kernel -k regex:flash_attn_kernel -s 2 -c 2 python tools/flow_only.py
HB=8 cap r2_hift_conv hift_conv_kernel -k regex:hift_conv_kernel -s 20 -c 4 python tools/hift_only.py
export CBX_DECODE_GRAPH=0 CBX_DECODE_PDL=0 TCLS=none TB=256 TSTEPS=24
cap r2_paged_bulk paged_bulk_kernel -k regex:paged_bulk_kernel -s 1300 -c 2 python tools/t3_only.py
cap r2_gemm_tc_decode gemm_tc_kernel -k regex:gemm_tc_kernel -s 5200 -c 5 python tools/t3_only.py
ls -la gpurun_out/*.ncu-rep gpurun_out/*.txt
