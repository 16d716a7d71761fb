#!/bin/bash
# Multi-GPU check: fused all-reduce+update correctness (thread + process mode, P2P + NVLS), bench at N GPUs for the
# native path and the DDP/NCCL baseline, and the all-reduce+update bandwidth sweep.
set -u
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "=== pytest multigpu"
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -8 | tee gpurun_out/pytest_multigpu.log
fi
for impl in ddp native; do
  echo "=== bench $impl x$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps ${STEPS:-30} --warmup 5 --impl $impl 2>&1 | grep -E '^\{' | tail -1 | tee gpurun_out/bench_${impl}_x$N.json
done
echo "=== allreduce+update sweep x$N"
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench_allreduce.py > gpurun_out/allreduce_x$N.log 2>&1
grep -E '^\{' gpurun_out/allreduce_x$N.log > gpurun_out/allreduce_x$N.jsonl
grep -m3 -i "nvls" gpurun_out/allreduce_x$N.log | cut -c1-200
tail -1 gpurun_out/allreduce_x$N.jsonl
python - <<PY
import json
rows=[json.loads(l) for l in open("gpurun_out/allreduce_x$N.jsonl")]
rows=[r for r in rows if "variant" in r and "bytes" in r]
sizes=sorted({r["bytes"] for r in rows})
print("bytes        " + "  ".join(f"{v:>16s}" for v in ["fused/p2p","fused/nvls","fused/oneshot","nccl+opt"]))
for s in sizes:
    d={r["variant"]:r for r in rows if r["bytes"]==s}
    print(f"{s:12d} " + "  ".join((f"{d[v]['us']:8.1f}us {d[v]['frac_of_roofline']*100:4.0f}%" if v in d else " "*16) for v in ["fused/p2p","fused/nvls","fused/oneshot","nccl+opt"]))
PY
