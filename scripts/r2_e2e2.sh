#!/bin/bash
set -u
O=gpurun_out/r2e2e2; mkdir -p $O
echo "=== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 2>$O/bench_err.log | tail -1 | tee $O/bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
tail -3 $O/bench_err.log
echo "=== bench api trainer"; timeout 600 python bench.py --steps 30 --warmup 5 --api trainer 2>>$O/bench_err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
echo "=== dz bench"; timeout 300 python scripts/dz_bench.py 2>&1 | tee $O/dz.txt | tail -5
echo "=== numerics"; timeout 600 python scripts/numerics_report.py 2>&1 | tee $O/numerics.txt | tail -40; cp gpurun_out/numerics_r2.json $O/ 2>/dev/null
echo "=== resnet50 bench"; timeout 600 python bench.py --model resnet50 --steps 20 --warmup 5 --no-e2e 2>>$O/bench_err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
echo "=== pytest resnet"; timeout 600 python -m pytest tests/test_native_gpu.py -m gpu -q --timeout 300 -k "resnet or block or chain" 2>&1 | tail -5
