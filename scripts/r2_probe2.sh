#!/bin/bash
set -u
O=gpurun_out/r2p2; mkdir -p $O
echo "=== probe mma rate"; timeout 150 python scripts/probe_mma_rate.py 2>&1 | tee $O/mma_rate.txt | tail -60
echo "=== probe tma bw"; timeout 120 python scripts/probe_tma_bw.py 2>&1 | tee $O/tma_bw.txt | tail -12
echo "=== halo ablate"; timeout 200 python scripts/halo_ablate.py 2>&1 | tee $O/halo_ablate.txt | tail -20
echo "=== pytest gpu (new defaults)"; timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -5 | tee $O/pytest.log
echo "=== bench (new defaults)"; timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | grep -E '^\{' | tail -1 | tee $O/bench.json | cut -c1-400
