#!/bin/bash
# Retry a gpurun call while the pod answers "transient" (busy slots; nothing is charged).
# usage: tools/gpurun_retry.sh <timeout> <script> [extra gpurun args, e.g. --gpus 2]
t=$1; sc=$2; shift 2
for i in $(seq 1 20); do
  out=$(gpurun --timeout "$t" "$@" -- "bash $sc" 2>&1)
  if echo "$out" | grep -q "status=transient"; then echo "[retry $i] pod busy"; sleep 120; continue; fi
  echo "$out" | tail -70
  exit 0
done
echo "gave up"
