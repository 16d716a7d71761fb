#!/bin/bash
# Multi-GPU validation: N = $N GPUs.  tests, benches (native with process-API e2e, ddp, ddp_tuned, staged), sweep.
set -u
N=${N:-2}
O=gpurun_out/r2multi_x$N; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
run() { timeout ${TMO:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "=== pytest multigpu"; timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 300 2>&1 | tail -12 | tee $O/pytest_multigpu.log
fi
p=29600
for cfg in ${CFGS:-native ddp ddp_tuned}; do
  p=$((p+1))
  echo "=== bench $cfg x$N"
  run $p bench.py --gpus $N --steps ${STEPS:-30} --warmup 5 --impl $cfg 2>$O/bench_${cfg}_err.log | grep -E '^\{' | tail -1 | tee $O/bench_${cfg}.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], (d.get('e2e') or {}), d['config'].get('comm'), d.get('replicas_synced'))
except Exception as e: print('FAILED', e)"
  tail -2 $O/bench_${cfg}_err.log | cut -c1-300
done
for envs in ${ENVS:-}; do
  p=$((p+1))
  echo "=== bench native [$envs]"
  env $(echo $envs | tr ';' ' ') timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $p bench.py --gpus $N --steps ${STEPS:-30} --warmup 5 --no-e2e 2>/dev/null | grep -E '^\{' | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])
except Exception as e: print('FAILED', e)"
done
if [ "${STAGED:-1}" = "1" ]; then
echo "=== bench staged (B1) x$N"; p=$((p+1))
TMO=900 run $p bench.py --gpus $N --steps 10 --warmup 3 --impl staged 2>$O/bench_staged_err.log | grep -E '^\{' | tail -1 | tee $O/bench_staged.json | cut -c1-300
tail -2 $O/bench_staged_err.log | cut -c1-300
fi
if [ "${SWEEP:-1}" = "1" ]; then
echo "=== allreduce+update sweep x$N"; p=$((p+1))
NCCL_DEBUG=INFO run $p bench_allreduce.py > $O/allreduce.log 2>&1
grep -E '^\{' $O/allreduce.log > $O/allreduce.jsonl
grep -m2 -i "nvls" $O/allreduce.log | cut -c1-160
python - <<PY
import json
rows=[json.loads(l) for l in open("$O/allreduce.jsonl")]
rows=[r for r in rows if "variant" in r and "bytes" in r]
sizes=sorted({r["bytes"] for r in rows})
V=["fused/p2p","fused/p2p/u4","fused/nvls","fused/nvls/u4","fused/oneshot","nccl+opt"]
print("bytes        " + "  ".join(f"{v:>20s}" for v in V))
for s in sizes:
    d={r["variant"]:r for r in rows if r["bytes"]==s}
    print(f"{s:12d} " + "  ".join((f"{d[v]['us']:8.1f}us {d[v]['frac_of_roofline']*100:3.0f}%/{(d[v].get('frac_of_900') or 0)*100:3.0f}%" if v in d else " "*20) for v in V))
PY
fi
