#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FB=16 NT=1 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none --csv --log-file gpurun_out/launches_flow.csv -k regex:"gemm_tc|attn_tc" python tools/flow_only.py > gpurun_out/flow_only.log 2>&1; echo "list exit=$?"; tail -n 2 gpurun_out/flow_only.log
FB=16 NT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 300 -c 4 -o gpurun_out/prof_gemm2 python tools/flow_only.py > gpurun_out/prof_gemm2.log 2>&1; echo "ncu gemm exit=$?"
FB=16 NT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 20 -c 2 -o gpurun_out/prof_attn_tc python tools/flow_only.py > gpurun_out/prof_attn_tc.log 2>&1; echo "ncu attn exit=$?"
