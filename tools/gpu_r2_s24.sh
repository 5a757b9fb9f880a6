#!/bin/bash
# Round 2, GPU session 24: final validation -- the whole GPU test suite, smoke(), the default bench (both arms) as the driver runs them.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
timeout 1200 $PT tests -x > gpurun_out/s24_tests.log 2>&1; echo "tests exit=$?"; tail -4 gpurun_out/s24_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
SECONDS=0; python bench.py > gpurun_out/s24_bench.log 2> gpurun_out/s24_bench.err; echo "bench exit=$? wall=${SECONDS}s"
tail -n 1 gpurun_out/s24_bench.log > gpurun_out/s24_bench_line.json
grep -E "^\[bench" gpurun_out/s24_bench.err | cut -c1-330 | tail -22
SECONDS=0; python bench.py --impl reference > gpurun_out/s24_ref.log 2> gpurun_out/s24_ref.err; echo "reference arm exit=$? wall=${SECONDS}s"; tail -n 1 gpurun_out/s24_ref.log | cut -c1-300
