#!/bin/bash
# Round 2, GPU session 27: the whole GPU test suite + smoke() on the final commit.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header tests > gpurun_out/s27_tests.log 2>&1; echo "tests exit=$?"; tail -4 gpurun_out/s27_tests.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
