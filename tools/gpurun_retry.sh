#!/bin/bash
# usage: tools/gpurun_retry.sh <out-file> <timeout> <command...>   — retries while the pod answers busy (exit 3 / transient)
out=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $out 2>&1
  if ! grep -q "status=transient\|retry in a few minutes\|no box" $out; then exit 0; fi
  sleep 90
done
