#!/bin/bash
# Retry a gpurun call while the pod answers "transient" (busy slots; nothing is charged).  usage: tools/gpurun_retry.sh <timeout> <script>
for i in $(seq 1 20); do
  out=$(gpurun --timeout "$1" -- "bash $2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then echo "[retry $i] pod busy"; sleep 120; continue; fi
  echo "$out" | tail -70
  exit 0
done
echo "gave up"
