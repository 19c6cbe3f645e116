#!/bin/bash
# usage: gpu_retry.sh <timeout> <script> <outfile>   — retries while the pod answers busy (nothing charged)
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" > "$3" 2>&1
  if ! grep -q "status=transient\|retry in a few minutes" "$3"; then exit 0; fi
  sleep 90
done
