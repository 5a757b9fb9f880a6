#!/bin/bash
# Round 2, GPU session 3: weight-resident persistent GEMM, fp16 CFM defaults, fp8 KV, streaming hooks; flow launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=8 t s3_wres 300 $PT tests/test_gpu_kernels.py -k "weight_resident"
TAILN=14 t s3_tests 900 $PT -s tests
t s3_flow_default 200 python tools/flow_breakdown.py
CBX_WRES=0 t s3_flow_nowres 200 python tools/flow_breakdown.py
FB=8 NT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/s3_flow_launches.csv python tools/flow_only.py > gpurun_out/s3_ncu_flow.log 2>&1
echo "ncu flow exit=$?"
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/s3_flow_launches.csv', errors='ignore')) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
agg=collections.defaultdict(list)
for r in rows[1:]:
    name=re.sub(r"\(.*","",r[ki]).replace("void ","").replace("cbx::","")
    agg[(name,r[gi])].append(float(r[vi].replace(",","")))
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1]))[:28]:
    print(f"{k[0][:44]:44s} {k[1]:18s} n={len(v):5d} avg={sum(v)/len(v)/1000:9.2f}us share={100*sum(v)/tot:5.1f}%")
PY
TCLS=none t s3_t3only_default 200 python tools/t3_only.py
TAILN=3 t s3_bench 900 python bench.py --steps 1 --warmup 1 --no-extra --cpu-sample none
tail -n 1 gpurun_out/s3_bench.log | cut -c1-1500
