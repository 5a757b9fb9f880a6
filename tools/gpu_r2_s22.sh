#!/bin/bash
# Round 2, GPU session 22: decode projections at 128 / 256 / 384 live rows (tile and split choices of the streaming kernel).
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
for m in 128 256 384; do echo "== rows $m"; DM=$m DREPS=3 timeout 200 python tools/decode_gemm_bench.py 2>&1 | grep -v head | tail -16 | cut -c1-120; done
