#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers "busy" (exit 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 45; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun_retry: still busy after 40 attempts"; exit 3
