#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free; nothing is charged for those).
#   tools/gpurun_retry.sh <timeout-seconds> '<command>'
t=$1; shift
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gpurun_retry: gave up after 40 transient answers"; exit 3
