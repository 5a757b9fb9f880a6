#!/bin/bash
# Round 2, GPU session 1: new decode path (bulk-copy paged attention, split-K, graphs, device-side retirement),
# long-configuration parity fixtures, fp16 operand formats of the CFM, stage timings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
t s1_kernels_new 300 $PT tests/test_gpu_kernels.py -k "paged or splitk"
t s1_t3 400 $PT tests/test_gpu_t3.py tests/test_gpu_turbo.py tests/test_gpu_variants.py
TAILN=30 t s1_long 600 $PT -s tests/test_gpu_long.py
CBX_EXPERIMENTAL=1 t s1_rest 600 $PT tests/test_gpu_s3gen.py tests/test_gpu_e2e.py tests/test_gpu_kernels.py
TCLS=none t s1_t3only_default 200 python tools/t3_only.py
TCLS=none CBX_DECODE_GRAPH=0 t s1_t3only_nograph 200 python tools/t3_only.py
TCLS=none CBX_DECODE_TILES=128,0,64,4,128,1,64,8 t s1_t3only_tilesA 200 python tools/t3_only.py
TCLS=none CBX_DECODE_TILES=64,1,64,1,128,1,64,2 t s1_t3only_tilesB 200 python tools/t3_only.py
TCLS=none CBX_DECODE_TILES=64,1,64,2,256,0,64,4 t s1_t3only_tilesC 200 python tools/t3_only.py
TCLS=paged t s1_t3only_paged 200 python tools/t3_only.py
TCLS=none TB=1 TSTEPS=150 t s1_t3only_b1 200 python tools/t3_only.py
t s1_flow_default 200 python tools/flow_breakdown.py
ATTN_PREC=fp16 t s1_flow_attn16 200 python tools/flow_breakdown.py
ATTN_PREC=fp16 CFM_ACT=fp16 t s1_flow_all16 200 python tools/flow_breakdown.py
t s1_bench 600 python bench.py --steps 1 --warmup 1
tail -n 3 gpurun_out/s1_bench.log | cut -c1-3000
