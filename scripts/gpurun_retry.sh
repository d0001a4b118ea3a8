#!/bin/bash
# usage: gpurun_retry.sh <out-log> <gpus> <timeout> <command...>   — retries while the pod answers busy (rc 3 / transient)
out=$1; gpus=$2; to=$3; shift 3
for attempt in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $out 2>&1; else /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" > $out 2>&1; fi
  if grep -q "status=transient\|retry in a few minutes\|no box or slot" $out; then sleep 60; continue; fi
  break
done
