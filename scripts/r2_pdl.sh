#!/bin/bash
# A/B of programmatic dependent launch per kernel class (FLUXDIST_PDL bit mask, csrc/common.cuh) on the final build.
set -u
O=gpurun_out/r2pdl; mkdir -p $O
echo "=== pytest ring/loader"; timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "staged_loader or ring or loader" 2>&1 | tail -3 | tee $O/pytest_ring.log
for cfg in "off:FLUXDIST_PDL=0" "all:FLUXDIST_PDL=0xff" "wgrad_only:FLUXDIST_PDL=0x0a" "no_wgrad:FLUXDIST_PDL=0xf5" "bn_only:FLUXDIST_PDL=0xf0" "conv_only:FLUXDIST_PDL=0x05" "bnfwd_only:FLUXDIST_PDL=0x10" "bnbwd_only:FLUXDIST_PDL=0x60" \
           "noside_off:FLUXDIST_PDL=0 FLUXDIST_SIDE_WGRAD=0" "noside_all:FLUXDIST_PDL=0xff FLUXDIST_SIDE_WGRAD=0" "off2:FLUXDIST_PDL=0"; do
  name=${cfg%%:*}; e=${cfg#*:}
  echo "=== bench $name [$e]"; env $e timeout 300 python bench.py --steps 40 --warmup 8 --no-e2e 2>$O/err_$name.log | grep -E '^\{' | tail -1 | tee $O/bench_$name.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4))
except Exception as e: print('FAILED', e)"
done
