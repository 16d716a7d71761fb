#!/bin/bash
set -u
O=gpurun_out/r2ev; mkdir -p $O
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee $O/pytest.log
echo "=== bench"; timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | tail -1 | tee $O/bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], (d['e2e'] or {}).get('value'))"
echo "=== ncu hot"; bash scripts/ncu_hot.sh 2>&1 | tail -6
echo "=== sanitizers"; SAN_TIMEOUT=200 bash scripts/sanitize.sh memcheck racecheck synccheck 2>&1 | tail -20
