#!/bin/bash
# Round 2, GPU session 18: producer-staged queries in the paged attention + retuned decode tiles (tests, repeatability, T3 timing),
# then the full default bench (both arms) exactly as the driver runs it, timed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
timeout 1200 $PT tests > gpurun_out/s18_tests.log 2>&1; echo "tests exit=$?"; tail -5 gpurun_out/s18_tests.log | cut -c1-300
echo "== paged repeatability"; PD_REPS=20 timeout 300 python tools/paged_determinism.py 2>&1 | tail -8
echo "== decode loop repeatability"; REPS=6 timeout 200 python tools/t3_determinism.py 2>&1 | tail -7
TCLS=none timeout 300 python tools/t3_only.py 2>&1 | tail -1
TCLS=paged,stream timeout 300 python tools/t3_only.py 2>&1 | tail -2
TCLS=none TB=1 TSTEPS=150 timeout 200 python tools/t3_only.py 2>&1 | tail -1
SECONDS=0; python bench.py > gpurun_out/s18_bench.log 2> gpurun_out/s18_bench.err; echo "bench exit=$? wall=${SECONDS}s"
tail -n 1 gpurun_out/s18_bench.log > gpurun_out/s18_bench_line.json
grep -E "^\[bench" gpurun_out/s18_bench.err | cut -c1-330 | tail -24
SECONDS=0; python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s18_ref.log 2> gpurun_out/s18_ref.err; echo "reference arm exit=$? wall=${SECONDS}s"
tail -n 1 gpurun_out/s18_ref.log | cut -c1-700
