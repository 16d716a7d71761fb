#!/bin/bash
# re-check of the per-shape policies on the PDL build (they were tuned before programmatic dependent launch)
set -u
O=gpurun_out/r2ab3; mkdir -p $O
for cfg in "default:FLUXDIST_X=1" "wgrad_igemm:FLUXDIST_WGRAD=igemm" "wgrad_halo:FLUXDIST_WGRAD=halo" "dz_rows_100k:FLUXDIST_DZ_MAX_ROWS=100000" "dz_rows_2m:FLUXDIST_DZ_MAX_ROWS=2000000" "halo_dz:FLUXDIST_HALO_DZ=1" "halo_dz_2m:FLUXDIST_HALO_DZ=1 FLUXDIST_DZ_MAX_ROWS=2000000" "narrow:FLUXDIST_WIDE=0" "epi4:FLUXDIST_EPI8=0" "pool_fused:FLUXDIST_POOL_BWD_FUSED=1"; do
  name=${cfg%%:*}; e=${cfg#*:}
  echo "=== bench $name [$e]"; env $e timeout 300 python bench.py --steps 40 --warmup 8 --no-e2e 2>$O/err_$name.log | grep -E '^\{' | tail -1 | tee $O/bench_$name.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4))
except Exception as e: print('FAILED', e)"
done
