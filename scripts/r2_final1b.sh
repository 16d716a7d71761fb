#!/bin/bash
# final 1-GPU pass on the PDL build: full GPU tests, the default bench line (100 steps), side-stream A/B, other configs, launch list
set -u
O=gpurun_out/r2final1b; mkdir -p $O
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee $O/pytest.log
echo "=== bench default"; timeout 400 python bench.py --steps 100 --warmup 10 2>$O/err_default.log | grep -E '^\{' | tail -1 | tee $O/bench_native_x1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d.get('gpu_launches'), d.get('clocks'))"
for cfg in "noside07:FLUXDIST_PDL=0x07 FLUXDIST_SIDE_WGRAD=0" "noside0f:FLUXDIST_PDL=0x0f FLUXDIST_SIDE_WGRAD=0" "nograph:FLUXDIST_X=1"; do
  name=${cfg%%:*}; e=${cfg#*:}; extra=""; [ $name = nograph ] && extra="--no-graph"
  echo "=== bench $name [$e]"; env $e timeout 300 python bench.py --steps 40 --warmup 8 --no-e2e $extra 2>$O/err_$name.log | grep -E '^\{' | tail -1 | tee $O/bench_$name.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4))
except Exception as e: print('FAILED', e)"
done
for cfg in "r50:--model resnet50 --steps 30 --warmup 5" "r18:--model resnet18 --steps 30 --warmup 5 --no-e2e" "adam:--opt adam --steps 30 --warmup 5 --no-e2e"; do
  name=${cfg%%:*}; a=${cfg#*:}
  echo "=== bench $name"; timeout 400 python bench.py $a 2>$O/err_$name.log | grep -E '^\{' | tail -1 | tee $O/bench_$name.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'))
except Exception as e: print('FAILED', e)"
done
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 420 --csv --log-file $O/launches.csv python bench.py --steps 6 --warmup 3 --no-graph --no-e2e > $O/launch_run.log 2>&1; tail -1 $O/launch_run.log | cut -c1-120
