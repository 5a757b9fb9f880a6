#!/bin/bash
# Round 2, GPU session 16: where does the fp32-KV batch path vary run to run?  (kernel-level probe, then the decode loop with PDL / graph off)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== paged attention, bitwise repeatability"; PD_REPS=30 timeout 300 python tools/paged_determinism.py 2>&1 | tail -8
echo "== decode loop, default"; REPS=6 timeout 200 python tools/t3_determinism.py 2>&1 | tail -7
echo "== decode loop, PDL off"; CBX_DECODE_PDL=0 REPS=6 timeout 200 python tools/t3_determinism.py 2>&1 | tail -7
echo "== decode loop, PDL off, graph off"; CBX_DECODE_PDL=0 CBX_DECODE_GRAPH=0 REPS=6 timeout 200 python tools/t3_determinism.py 2>&1 | tail -7
echo "== decode loop, early stage release"; CBX_PB_EARLY=1 REPS=6 timeout 200 python tools/t3_determinism.py 2>&1 | tail -7
python bench.py --steps 1 --warmup 1 --no-extra --cpu-sample none > gpurun_out/s16_bench.log 2>&1; echo "bench exit=$?"
tail -n 1 gpurun_out/s16_bench.log > gpurun_out/s16_bench_line.json
grep -E "warmup|timed|profile" gpurun_out/s16_bench.log | cut -c1-400
