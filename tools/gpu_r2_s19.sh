#!/bin/bash
# Round 2, GPU session 19 (1 GPU): compute-sanitizer memcheck of the kernel tests, ncu --set full captures + launch lists on the final build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
PT="python -m pytest -q -m gpu -p no:cacheprovider --no-header"
timeout 200 $PT tests/test_gpu_kernels.py -k "fp16_plane or splitk" 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest -q -m gpu -p no:cacheprovider --no-header tests/test_gpu_kernels.py -x -k "not simt" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck exit=$?"; tail -6 gpurun_out/r2_sanitizer_memcheck.log | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on"
cap() { local name=$1; shift; timeout 400 $NCU -o gpurun_out/$name "$@" > gpurun_out/s19_ncu_$name.log 2>&1; echo "ncu $name exit=$?"; }
FB=8 NT=1 cap r2f_attn_otm -k regex:attn_otm_kernel -s 20 -c 2 python tools/flow_only.py
FB=8 NT=1 cap r2f_gemm_tc_flow -k regex:gemm_tc_kernel -s 120 -c 4 python tools/flow_only.py
HB=8 cap r2f_hift_conv -k regex:hift_conv_kernel -s 20 -c 3 python tools/hift_only.py
export CBX_DECODE_GRAPH=0 CBX_DECODE_PDL=0 TCLS=none TB=256 TSTEPS=24
cap r2f_paged_bulk -k regex:paged_bulk_kernel -s 400 -c 2 python tools/t3_only.py
cap r2f_gemm_stream -k regex:gemm_stream_kernel -s 1600 -c 5 python tools/t3_only.py
unset CBX_DECODE_GRAPH CBX_DECODE_PDL
LL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
FB=16 NT=2 timeout 600 $LL --log-file gpurun_out/r2f_launches_flow.csv python tools/flow_only.py > /dev/null 2>&1; python tools/summarize_launches.py gpurun_out/r2f_launches_flow.csv > gpurun_out/r2f_launches_flow.txt; head -14 gpurun_out/r2f_launches_flow.txt
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
