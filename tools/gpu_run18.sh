#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 2 gpurun_out/$name.log | cut -c1-4000; }
t t3 300 python -m pytest tests/test_gpu_t3.py -q -m gpu -p no:cacheprovider --timeout 200
export CBX_BENCH_WATCHDOG=800
t bench_s1 900 python bench.py --steps 1 --warmup 1
PB=4 PSTEPS=12 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_small.csv python tools/profile_small.py > gpurun_out/profile_small.log 2>&1; echo "ncu list exit=$?"
FB=16 NT=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 300 -c 3 -o gpurun_out/r1_prof_gemm_tc python tools/flow_only.py > /dev/null 2>&1; echo "ncu gemm exit=$?"
FB=16 NT=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 20 -c 2 -o gpurun_out/r1_prof_attn_tc python tools/flow_only.py > /dev/null 2>&1; echo "ncu attn exit=$?"
PB=64 PSTEPS=48 timeout 400 ncu --set full --clock-control none --import-source on -k regex:paged_decode -s 600 -c 2 -o gpurun_out/r1_prof_paged python tools/profile_small.py > /dev/null 2>&1; echo "ncu paged exit=$?"
ls -la gpurun_out/r1_*
