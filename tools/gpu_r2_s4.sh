#!/bin/bash
# Round 2, GPU session 4: UMMA row-shift probe, PDL in the decode step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-700; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=3 t s4_probe 120 python tools/probe_rowshift.py
TAILN=6 t s4_tests_t3 600 $PT tests/test_gpu_t3.py tests/test_gpu_turbo.py tests/test_gpu_long.py -k "not flow"
TCLS=none t s4_t3only_pdl 200 python tools/t3_only.py
CBX_DECODE_PDL=0 TCLS=none t s4_t3only_nopdl 200 python tools/t3_only.py
TCLS=none TB=1 TSTEPS=150 t s4_b1_pdl 200 python tools/t3_only.py
CBX_DECODE_PDL=0 TCLS=none TB=1 TSTEPS=150 t s4_b1_nopdl 200 python tools/t3_only.py
TCLS=none TB=32 TSTEPS=300 t s4_b32_pdl 200 python tools/t3_only.py
CBX_DECODE_PDL=0 TCLS=none TB=32 TSTEPS=300 t s4_b32_nopdl 200 python tools/t3_only.py
