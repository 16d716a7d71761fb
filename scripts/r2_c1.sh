#!/bin/bash
set -u
O=gpurun_out/r2c1; mkdir -p $O
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee $O/pytest.log
bash scripts/r2_ab.sh "default:" "nopoolfuse:FLUXDIST_POOL_BWD_FUSED=0" "nohalodz:FLUXDIST_HALO_DZ=0" "neither:FLUXDIST_POOL_BWD_FUSED=0 FLUXDIST_HALO_DZ=0" 2>&1 | grep -v "^===" | tail -5
echo "=== conv bench"; timeout 600 python scripts/conv_bench_r2.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    keys=['hw','C','K','R','stride','cudnn_fprop','fprop_auto','fprop_stats_auto','halo_fprop_v1','halo_fprop_stats_v1','cudnn_dgrad','dgrad_auto','halo_dgrad_v1','cudnn_wgrad','wgrad_im2col_s0','halo_wgrad']
    print({k:d.get(k) for k in keys if d.get(k) is not None})
"
