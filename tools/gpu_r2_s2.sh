#!/bin/bash
# Round 2, GPU session 2: persistent paged kernel, fp16 CFM operand formats (fp16 weight copies), attention occupancy
# variants, per-kernel launch list of the decode step, functional run of every bench leg at a small shape.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=12 t s2_tests 900 $PT -s tests
TCLS=none t s2_t3only_default 200 python tools/t3_only.py
TCLS=paged t s2_t3only_paged 200 python tools/t3_only.py
TCLS=gemm_tc t s2_t3only_gemm 200 python tools/t3_only.py
TCLS=none TB=1 TSTEPS=150 t s2_t3only_b1 200 python tools/t3_only.py
CBX_DECODE_GRAPH=0 TCLS=none TB=256 TSTEPS=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1200 --csv --log-file gpurun_out/s2_decode_launches.csv python tools/t3_only.py > gpurun_out/s2_ncu_decode.log 2>&1
echo "ncu decode exit=$?"; python tools/summarize_launches.py gpurun_out/s2_decode_launches.csv | head -20
for v in "1 1" "1 2" "2 1" "2 2"; do set -- $v
  ATTN_PREC=fp16 CBX_ATTN_TC=$1 CBX_ATTN_OCC=$2 t s2_flow_attn16_tc$1_occ$2 200 python tools/flow_breakdown.py
done
ATTN_PREC=fp16 CFM_ACT=fp16 t s2_flow_all16 200 python tools/flow_breakdown.py
ATTN_PREC=fp16 CFM_ACT=fp16 CBX_ATTN_TC=2 t s2_flow_all16_tc2 200 python tools/flow_breakdown.py
t s2_hift 200 python tools/hift_only.py
TAILN=3 t s2_bench_small 900 python bench.py --steps 1 --warmup 1 --batch 32 --budget-max 300
tail -n 1 gpurun_out/s2_bench_small.log | cut -c1-6000
