#!/bin/bash
# Per-kernel device-time list of one training step (cold-cache, serialised: compare SHARES).
set -u
mkdir -p gpurun_out
IMPL=${IMPL:-native}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1500} -c ${COUNT:-900} --csv \
  --log-file gpurun_out/launches_${IMPL}.csv python bench.py --impl $IMPL --steps 2 --warmup 3 --no-graph --no-e2e \
  > gpurun_out/launches_${IMPL}.log 2>&1
tail -2 gpurun_out/launches_${IMPL}.log
wc -l gpurun_out/launches_${IMPL}.csv
