#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> [--gpus N] '<command>'   -- retries while the pod answers "busy" (transient)
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift; shift; fi
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T $G -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 45; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun_retry: still busy after 40 attempts"; exit 3
