#!/bin/bash
set -u
O=gpurun_out/r2ab2; mkdir -p $O
echo "=== pytest gpu (1 GPU visible tests)"; timeout 900 python -m pytest tests/test_native_gpu.py -m gpu -q --timeout 300 2>&1 | tail -6 | tee $O/pytest.log
echo "=== bench 1 GPU"; timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
N=2 SKIP_TESTS=1 STAGED=0 SWEEP=0 CFGS="" ENVS="X=1 FLUXDIST_BUCKETS=0 FLUXDIST_OVERLAP=0 FLUXDIST_EARLY_GRID=16 FLUXDIST_EARLY_GRID=64 FLUXDIST_BUCKETS=0.6 FLUXDIST_NVLS=1" bash scripts/r2_multi.sh
