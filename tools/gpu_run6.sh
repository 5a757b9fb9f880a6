#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/flow_breakdown.py > gpurun_out/flow_breakdown.log 2>&1; echo "breakdown exit=$?"; cat gpurun_out/flow_breakdown.log | tail -5
PB=4 PSTEPS=12 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_small.csv python tools/profile_small.py > gpurun_out/profile_small.log 2>&1; echo "ncu list exit=$?"; tail -n 2 gpurun_out/profile_small.log
PB=4 PSTEPS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 300 -c 2 -o gpurun_out/prof_flash python tools/profile_small.py > gpurun_out/prof_flash.log 2>&1; echo "ncu flash exit=$?"
PB=4 PSTEPS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1500 -c 3 -o gpurun_out/prof_gemm python tools/profile_small.py > gpurun_out/prof_gemm.log 2>&1; echo "ncu gemm exit=$?"
ls -la gpurun_out/*.ncu-rep
