#!/bin/bash
set -u
O=gpurun_out/r2c2; mkdir -p $O
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee $O/pytest.log
RUN_TESTS=0 bash scripts/r2_ab.sh "default:" "epi4:FLUXDIST_EPI8=0" 2>&1 | grep -v "^===" | tail -3
echo "=== conv bench"; timeout 600 python scripts/conv_bench_r2.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    keys=['hw','C','K','R','stride','cudnn_fprop','fprop_auto','fprop_stats_auto','halo_fprop_v1','halo_fprop_stats_v1','cudnn_dgrad','dgrad_auto','halo_dgrad_v1','cudnn_wgrad','wgrad_im2col_s0','halo_wgrad']
    print({k:d.get(k) for k in keys if d.get(k) is not None})
"
echo "=== dz bench"; timeout 300 python scripts/dz_bench.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print({k:v for k,v in d.items() if 'auto' in k or k in ('hw','C','bn_bwd_fused_us','bn_bwd_apply_us')})
"
