#!/bin/bash
# ncu --set full of the hot kernels (one GPU).  Report -> gpurun_out/ncu_top.ncu-rep; summarise with scripts/ncu_summary.py
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/ncu_top \
  -k regex:'igemm_kernel|conv_halo_kernel|wgrad_halo_kernel|bn_apply_kernel|bn_bwd_apply|bn_bwd_fused|bn_relu_pool|maxpool_bwd_idx' -s 13 -c 13 \
  python scripts/ncu_top.py > gpurun_out/ncu_top.log 2>&1
tail -3 gpurun_out/ncu_top.log
ls -la gpurun_out/ncu_top.ncu-rep
