#!/bin/bash
# final 8-GPU numbers: native (100 steps), Adam, ResNet-50 (process mode), B1 staged.  No tests, no sweep (done earlier).
set -u
N=8; O=gpurun_out/r2final8; mkdir -p $O
run() { timeout ${TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); e=d.get('e2e') or {}; print(d['value'], d['ms_per_step'], e.get('value'), e.get('api'), d.get('clocks'), d.get('replicas_synced'))
except Exception as ex: print('FAILED', ex)"; }
echo "=== native x8 100 steps"; run 29701 bench.py --gpus 8 --steps 100 --warmup 10 2>$O/err_native.log | grep -E '^\{' | tail -1 | tee $O/bench_native_x8.json | show
echo "=== adam x8"; run 29702 bench.py --gpus 8 --steps 50 --warmup 5 --opt adam 2>$O/err_adam.log | grep -E '^\{' | tail -1 | tee $O/bench_adam_x8.json | show
echo "=== resnet50 x8"; run 29703 bench.py --gpus 8 --steps 50 --warmup 5 --model resnet50 2>$O/err_r50.log | grep -E '^\{' | tail -1 | tee $O/bench_resnet50_x8.json | show
echo "=== staged B1 x8"; TMO=240 run 29704 bench.py --gpus 8 --steps 8 --warmup 3 --impl staged 2>$O/err_staged.log | grep -E '^\{' | tail -1 | tee $O/bench_staged_x8.json | cut -c1-200
tail -2 $O/err_staged.log | cut -c1-200
