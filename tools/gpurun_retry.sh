#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   — retries while the pod answers "busy / transient" (nothing is charged for those)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|answers busy\|no box or slot"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gpurun_retry: gave up after 40 attempts"; exit 3
