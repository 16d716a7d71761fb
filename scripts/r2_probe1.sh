#!/bin/bash
# Round-2 first GPU call: hardware probes + state of the experimental switches left by round 1.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2p1
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "=== pytest gpu (baseline)"; timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -5 | tee $O/pytest_base.log
echo "=== probe mma rate"; timeout 120 python scripts/probe_mma_rate.py 2>&1 | tee $O/mma_rate.txt | tail -40
echo "=== probe tma bw"; timeout 120 python scripts/probe_tma_bw.py 2>&1 | tee $O/tma_bw.txt | tail -12
echo "=== halo ablate"; timeout 200 python scripts/halo_ablate.py 2>&1 | tee $O/halo_ablate.txt | tail -20
echo "=== pytest EPI8"; FLUXDIST_EPI8=1 timeout 600 python -m pytest tests/test_native_gpu.py -m gpu -q --timeout 300 2>&1 | tail -8 | tee $O/pytest_epi8.log
echo "=== pytest DZ_CHAIN"; FLUXDIST_DZ_CHAIN=1 timeout 600 python -m pytest tests/test_native_gpu.py -m gpu -q --timeout 300 -k "chain or block or resnet" 2>&1 | tail -8 | tee $O/pytest_chain.log
for cfg in "base:" "epi8:FLUXDIST_EPI8=1" "chain:FLUXDIST_DZ_CHAIN=1" "both:FLUXDIST_EPI8=1 FLUXDIST_DZ_CHAIN=1" "epi8_rows:FLUXDIST_EPI8=1 FLUXDIST_DZ_MAX_ROWS=400000" "both_rows:FLUXDIST_EPI8=1 FLUXDIST_DZ_CHAIN=1 FLUXDIST_DZ_MAX_ROWS=400000"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "=== bench $name ($envs)"
  env $envs timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e 2>&1 | grep -E '^\{' | tail -1 | tee $O/bench_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['final_loss'])"
done
echo "=== conv bench"; timeout 400 python scripts/conv_bench.py 2>&1 | tail -6; cp gpurun_out/conv_bench.json $O/conv_bench_base.json
echo "=== conv bench EPI8 (im2col rows only matter)"; FLUXDIST_EPI8=1 timeout 400 python scripts/conv_bench.py 2>&1 | tail -6; cp gpurun_out/conv_bench.json $O/conv_bench_epi8.json
echo "=== stats bench"; timeout 200 python scripts/stats_bench.py 2>&1 | tee $O/stats_base.txt | tail -8
FLUXDIST_EPI8=1 timeout 200 python scripts/stats_bench.py 2>&1 | tee $O/stats_epi8.txt | tail -8
echo "=== dz bench"; timeout 200 python scripts/dz_bench.py 2>&1 | tee $O/dz_base.txt | tail -8
FLUXDIST_EPI8=1 timeout 200 python scripts/dz_bench.py 2>&1 | tee $O/dz_epi8.txt | tail -8
