#!/bin/bash
# usage: scripts/gpurun_retry.sh <outfile> <gpurun args...>   -- retries while the pod answers "transient" (nothing charged)
out=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if ! grep -q "status=transient" "$out"; then break; fi
  sleep 90
done
tail -60 "$out"
