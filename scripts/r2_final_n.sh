#!/bin/bash
# N-GPU validation of the final (PDL) build: multi-GPU tests (TESTS=1) and the native bench line.  N=2|4|8.
set -u
N=${N:-2}; O=gpurun_out/r2final_x$N; mkdir -p $O
run() { timeout ${TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); e=d.get('e2e') or {}; print(d['value'], d['ms_per_step'], e.get('value'), e.get('api'), d.get('clocks'), d.get('replicas_synced'))
except Exception as ex: print('FAILED', ex)"; }
if [ "${TESTS:-1}" = 1 ]; then echo "=== pytest multigpu"; timeout 600 python -m pytest tests/test_multigpu.py -q --timeout 300 2>&1 | tail -4 | tee $O/pytest_multigpu.log; fi
echo "=== native x$N"; run 29711 bench.py --gpus $N --steps ${STEPS:-100} --warmup 10 2>$O/err_native.log | grep -E '^\{' | tail -1 | tee $O/bench_native_x$N.json | show
if [ "${MORE:-0}" = 1 ]; then
echo "=== resnet50 x$N"; run 29713 bench.py --gpus $N --steps 50 --warmup 5 --model resnet50 2>$O/err_r50.log | grep -E '^\{' | tail -1 | tee $O/bench_resnet50_x$N.json | show
fi
