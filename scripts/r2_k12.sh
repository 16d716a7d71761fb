#!/bin/bash
set -u
O=gpurun_out/r2k12; mkdir -p $O
echo "=== pytest native gpu"; timeout 900 python -m pytest tests/test_native_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -12 | tee $O/pytest.log
echo "=== conv bench r2"; timeout 600 python scripts/conv_bench_r2.py 2>&1 | tail -12 | tee $O/conv_bench.log; cp gpurun_out/conv_bench_r2.json $O/
echo "=== pytest all gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee $O/pytest_all.log
echo "=== bench"; timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e 2>&1 | grep -E '^\{' | tail -1 | tee $O/bench.json | cut -c1-300
echo "=== bench narrow"; FLUXDIST_WIDE=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e 2>&1 | grep -E '^\{' | tail -1 | tee $O/bench_narrow.json | cut -c1-200
echo "=== bench wgrad halo"; FLUXDIST_WGRAD=halo timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e 2>&1 | grep -E '^\{' | tail -1 | tee $O/bench_wgradhalo.json | cut -c1-200
