#!/bin/bash
# compute-sanitizer targets: memcheck / racecheck / synccheck on the small-shape kernel tests.
set -u
TOOL=${1:-memcheck}
compute-sanitizer --tool "$TOOL" --error-exitcode 1 python -m pytest tests/test_native_gpu.py -m gpu -q -x \
  -k "gemm_nt or conv_fprop or softmax or bn_single or halo_wgrad or fused_update_single" --timeout 1200
