#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python tools/gemm_anatomy.py 2>&1 | tail -30
FB=2 timeout 200 python tools/flow_only.py 2>&1 | tail -12
