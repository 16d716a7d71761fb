#!/bin/bash
# NOTE: FLUXDIST_PDL_LATE existed only for this experiment (results: profiles/r2/pdl_ab.txt, run 3); the switch was removed afterwards.
# third PDL A/B: conv kernels release their dependents late (FLUXDIST_PDL_LATE=1) so BN kernels can be dependents
set -u
O=gpurun_out/r2pdl3; mkdir -p $O
for cfg in "0:0x07" "1:0x07" "1:0x17" "1:0x27" "1:0x47" "1:0x77" "1:0xf7" "1:0x0f" "0:0x07"; do
  late=${cfg%%:*}; m=${cfg#*:}
  echo "=== bench late=$late mask $m"; FLUXDIST_PDL_LATE=$late FLUXDIST_PDL=$m timeout 300 python bench.py --steps 40 --warmup 8 --no-e2e 2>$O/err_${late}_$m.log | grep -E '^\{' | tail -1 | tee -a $O/bench_${late}_$m.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4))
except Exception as e: print('FAILED', e)"
done
