#!/bin/bash
# second PDL A/B: combinations around "all conv classes" (FLUXDIST_PDL bit mask, csrc/common.cuh)
set -u
O=gpurun_out/r2pdl2; mkdir -p $O
for m in 0x0f 0x2f 0x4f 0x6f 0x8f 0x07 0x0d 0x0e 0x0f 0x00; do
  echo "=== bench mask $m"; FLUXDIST_PDL=$m timeout 300 python bench.py --steps 40 --warmup 8 --no-e2e 2>$O/err_$m.log | grep -E '^\{' | tail -1 | tee $O/bench_$m.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4))
except Exception as e: print('FAILED', e)"
done
