#!/bin/bash
# stream-priority x PDL A/B (side stream = weight gradients, main = captured step) + the 1-GPU debug-mode tests
set -u
O=gpurun_out/r2prio; mkdir -p $O
echo "=== pytest verify_exchange (1 GPU)"; timeout 600 python -m pytest tests/test_native_gpu.py -q --timeout 300 -k "verify_exchange or staged_loader" 2>&1 | tail -4 | tee $O/pytest_verify.log
for cfg in "0x07:0:-1" "0x07:-1:-1" "0x07:-1:0" "0x07:-2:-1" "0x77:-1:-1" "0x77:-1:0" "0x77:-2:-1" "0x67:-2:-1" "0x17:-2:-1" "0x00:-1:0"; do
  IFS=: read m sp mp <<< "$cfg"
  echo "=== bench mask $m side_prio $sp main_prio $mp"; FLUXDIST_PDL=$m FLUXDIST_SIDE_PRIO=$sp FLUXDIST_MAIN_PRIO=$mp timeout 300 python bench.py --steps 40 --warmup 8 --no-e2e 2>$O/err_${m}_${sp}_${mp}.log | grep -E '^\{' | tail -1 | tee $O/bench_${m}_${sp}_${mp}.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4))
except Exception as e: print('FAILED', e)"
done
