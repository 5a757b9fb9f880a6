#!/bin/bash
# Round 2, GPU session 20: paged attention 3 vs 4 CTAs per SM; conv-GEMM tile choice (dual 128 vs 256-wide) in the flow stage.
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
echo "== T3 400 steps, paged OCC 3 (default)"; TCLS=none,paged timeout 300 python tools/t3_only.py 2>&1 | tail -2 | cut -c1-160
echo "== paged OCC 4"; CBX_PB_OCC=4 TCLS=none,paged timeout 300 python tools/t3_only.py 2>&1 | tail -2 | cut -c1-160
echo "== flow (FB=16 NT=4) default tiles"; FB=16 NT=4 FCLS=none,all timeout 200 python tools/flow_only.py 2>&1 | tail -2 | cut -c1-200
echo "== flow, CBX_TILE=1 (no dual-CTA tiles: 256-wide tiles for the N=256 GEMMs)"; CBX_TILE=1 FB=16 NT=4 FCLS=none,all timeout 200 python tools/flow_only.py 2>&1 | tail -2 | cut -c1-200
