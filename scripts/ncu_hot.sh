#!/bin/bash
# ncu --set full of the hot kernels (one GPU).  Report -> gpurun_out/ncu_hot.ncu-rep; summarise here with
# scripts/ncu_export.sh (ncu -i ... --page details / raw / source).
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/ncu_hot \
  -k regex:'igemm_kernel|conv_halo_kernel|wgrad_halo_kernel|bn_apply_kernel|bn_bwd_apply_kernel|bn_bwd_fused_kernel|bn_relu_pool3s2|maxpool_bwd_idx|fused_allreduce_update' \
  -s 17 -c 17 python scripts/ncu_hot.py > gpurun_out/ncu_hot.log 2>&1
tail -20 gpurun_out/ncu_hot.log
ls -la gpurun_out/ncu_hot.ncu-rep
