#!/bin/bash
set -u
O=gpurun_out/r2final1; mkdir -p $O
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee $O/pytest.log
RUN_TESTS=0 bash scripts/r2_ab.sh "default:" "halodz:FLUXDIST_HALO_DZ=1" "poolperpixel:FLUXDIST_POOL_BWD_BLOCK=0" 2>&1 | grep -v "^===" | tail -4
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 420 --csv --log-file $O/launches.csv python bench.py --steps 6 --warmup 3 --no-graph --no-e2e > $O/launch_run.log 2>&1; tail -1 $O/launch_run.log | cut -c1-120
for cfg in "staged:--impl staged --steps 10 --warmup 3" "ddp:--impl ddp --steps 30 --warmup 5" "r50:--model resnet50 --steps 20 --warmup 5" "r18:--model resnet18 --steps 30 --warmup 5 --no-e2e" "adam:--opt adam --steps 30 --warmup 5 --no-e2e"; do
  name=${cfg%%:*}; a=${cfg#*:}
  echo "=== bench $name"; timeout 400 python bench.py $a 2>$O/err_$name.log | grep -E '^\{' | tail -1 | tee $O/bench_$name.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'))
except Exception as e: print('FAILED', e)"
done
echo "=== conv bench"; timeout 600 python scripts/conv_bench_r2.py > $O/conv_bench.log 2>&1; cp gpurun_out/conv_bench_r2.json $O/; tail -1 $O/conv_bench.log | cut -c1-200
