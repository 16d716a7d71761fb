#!/bin/bash
set -u
O=gpurun_out/r2e2e; mkdir -p $O
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee $O/pytest.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench default (20 steps like the driver)"; timeout 600 python bench.py --steps 20 --warmup 5 2>$O/bench_err.log | tail -1 | tee $O/bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks'])"
tail -5 $O/bench_err.log
echo "=== bench 100 steps"; timeout 600 python bench.py 2>>$O/bench_err.log | tail -1 | tee $O/bench100.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
echo "=== bench api trainer"; timeout 600 python bench.py --steps 30 --warmup 5 --api trainer 2>>$O/bench_err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 420 --csv --log-file $O/launches.csv python bench.py --steps 6 --warmup 3 --no-graph --no-e2e > $O/launch_run.log 2>&1; tail -2 $O/launch_run.log | cut -c1-200
