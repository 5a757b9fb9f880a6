#!/bin/bash
# Round 2, GPU session 26: weight-resident GEMM with 4 A stages (231.7 KB of shared memory) -- tests, flow timing, short bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
timeout 400 $PT tests/test_gpu_kernels.py tests/test_gpu_s3gen.py -k "fp16_plane or cfm or mel or flow" 2>&1 | tail -2
FB=16 NT=4 FCLS=none,all timeout 200 python tools/flow_only.py 2>&1 | tail -2 | cut -c1-200
python bench.py --steps 1 --warmup 1 --no-extra --cpu-sample none > gpurun_out/s26_bench.log 2> gpurun_out/s26_bench.err; echo "bench exit=$?"
grep -E "warmup|timed|profile" gpurun_out/s26_bench.err | cut -c1-330
