#!/bin/bash
# final 1-GPU pass (PDL + halo dz epilogue defaults): full GPU tests, smoke, default bench line (100 steps), other configs
set -u
O=gpurun_out/r2final1c; mkdir -p $O
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee $O/pytest.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench default"; timeout 400 python bench.py --steps 100 --warmup 10 2>$O/err_default.log | grep -E '^\{' | tail -1 | tee $O/bench_native_x1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d.get('gpu_launches'), d.get('clocks'))"
for cfg in "r50:--model resnet50 --steps 30 --warmup 5" "r18:--model resnet18 --steps 30 --warmup 5 --no-e2e" "adam:--opt adam --steps 30 --warmup 5 --no-e2e" "halodz0:--steps 40 --warmup 8 --no-e2e"; do
  name=${cfg%%:*}; a=${cfg#*:}; e="FLUXDIST_X=1"; [ $name = halodz0 ] && e="FLUXDIST_HALO_DZ=0"
  echo "=== bench $name"; env $e timeout 400 python bench.py $a 2>$O/err_$name.log | grep -E '^\{' | tail -1 | tee $O/bench_$name.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'))
except Exception as e: print('FAILED', e)"
done
