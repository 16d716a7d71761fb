#!/bin/bash
# GPU check: native numerics tests (grouped, each under its own timeout so one hung kernel does not
# hide the rest) + smoke + short benches; all output kept under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
: > gpurun_out/pytest_gpu.log
for k in "gemm" "conv_fprop" "conv_dgrad" "conv_wgrad" "bn_ or maxpool or softmax or normalize or staging" "convbn or resnet18" "fused_update"; do
  echo "##### -k $k" | tee -a gpurun_out/pytest_gpu.log
  timeout 420 python -m pytest tests/test_native_gpu.py tests/test_workflow.py -m gpu -q --timeout 300 -k "$k" 2>&1 | grep -vE "^\s*$" | tail -${TAILN:-40} | tee -a gpurun_out/pytest_gpu.log
done
if [ "${SKIP_BENCH:-0}" != "1" ]; then
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "=== bench torch"
timeout 300 python bench.py --impl torch --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_torch.log
echo "=== bench native"
timeout 300 python bench.py --impl native --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_native.log
fi
