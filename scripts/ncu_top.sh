#!/bin/bash
# ncu --set full of the hot kernels (one GPU).  Report -> gpurun_out/ncu_top.ncu-rep; summarise with scripts/ncu_summary.py
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/ncu_top \
  -k regex:'igemm_kernel|conv_halo_kernel|wgrad_halo_kernel|bn_fwd_fused|bn_bwd_fused' -s 11 -c 11 \
  python scripts/ncu_top.py > gpurun_out/ncu_top.log 2>&1
tail -3 gpurun_out/ncu_top.log
ls -la gpurun_out/ncu_top.ncu-rep
