#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
t t3_tests 300 python -m pytest tests/test_gpu_t3.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout 200 -x -k "not simt"
TCLS=paged t t3_wide 200 python tools/t3_only.py
CBX_LIB=$PWD/chatterbox_b200/libcbx_narrow.so TCLS=paged t t3_narrow 200 python tools/t3_only.py
TCLS=paged TSTEPS=330 timeout 300 ncu --set full --clock-control none -k regex:paged_decode --launch-skip 9000 --launch-count 2 -f -o gpurun_out/r1_prof_paged_b256 python tools/t3_only.py > gpurun_out/prof_paged_b256.log 2>&1; echo "ncu exit=$?"; tail -n 2 gpurun_out/prof_paged_b256.log
