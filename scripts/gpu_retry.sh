#!/bin/bash
# usage: scripts/gpu_retry.sh OUTFILE <gpurun args...>   retries while the pod answers "transient" (no slot / draining)
OUT=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$OUT" 2>&1
  if ! grep -q "status=transient" "$OUT"; then exit 0; fi
  sleep 120
done
