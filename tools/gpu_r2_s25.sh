#!/bin/bash
# Round 2, GPU session 25: last check of the committed state -- repeatability probes, the long tests three times over, a short bench run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
echo "== paged repeatability"; PD_REPS=25 timeout 300 python tools/paged_determinism.py 2>&1 | tail -8
echo "== decode loop repeatability"; REPS=8 timeout 200 python tools/t3_determinism.py 2>&1 | tail -9
for i in 1 2 3; do timeout 300 $PT tests/test_gpu_long.py tests/test_gpu_t3.py 2>&1 | tail -1; done
python bench.py --steps 1 --warmup 1 --no-extra --cpu-sample none > gpurun_out/s25_bench.log 2> gpurun_out/s25_bench.err; echo "bench exit=$?"
grep -E "timed|e2e|profile" gpurun_out/s25_bench.err | cut -c1-300
tail -n 1 gpurun_out/s25_bench.log | cut -c1-200
