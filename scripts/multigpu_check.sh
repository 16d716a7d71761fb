#!/bin/bash
# Multi-GPU check: process-mode fused all-reduce+update over symmetric memory (P2P and NVLS), thread-mode
# (task-based) world, and the bench at N GPUs.
set -u
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "=== pytest multigpu"
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -30 | tee gpurun_out/pytest_multigpu.log
for impl in ddp native; do
  echo "=== bench $impl x$N"
  NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --impl $impl 2>&1 | tail -4 | tee gpurun_out/bench_${impl}_x$N.log
done
