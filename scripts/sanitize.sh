#!/bin/bash
# compute-sanitizer targets: memcheck / racecheck / synccheck / initcheck over the small-shape kernel tests.
# usage: scripts/sanitize.sh [tool ...]   (default: memcheck racecheck synccheck); logs -> gpurun_out/sanitizer_<tool>.txt
#        scripts/sanitize.sh host         ThreadSanitizer over the C++ host code of the pinned staging ring (no GPU needed):
#                                         csrc/staging.cpp built against the mock CUDA runtime (csrc/mock) with
#                                         -fsanitize=thread, driven by tests/test_staging_mock.py; log -> gpurun_out/tsan_staging.*
set -u
mkdir -p gpurun_out
if [ "${1:-}" = host ]; then
  python - <<'PY' || exit 1
import importlib.util
spec = importlib.util.spec_from_file_location("b", "fluxdistributed.jl_b200/build.py")
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
print("built", b.build_staging_mock(tsan=True))
PY
  rm -f gpurun_out/tsan_staging.*
  LD_PRELOAD=$(gcc -print-file-name=libtsan.so) \
  TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0 log_path=gpurun_out/tsan_staging" \
  FLUXDIST_STAGING_MOCK_MODULE=fluxdist._Cstaging_mock_tsan \
    timeout 900 python -m pytest tests/test_staging_mock.py -q -p no:cacheprovider | tail -2
  n=$(cat gpurun_out/tsan_staging.* 2>/dev/null | grep -c "WARNING: ThreadSanitizer")
  echo "ThreadSanitizer warnings: $n"
  exit $([ "$n" = 0 ] && echo 0 || echo 1)
fi
TOOLS=${@:-memcheck racecheck synccheck}
# small shapes only: the sanitizers serialise and instrument every access (50-500x slower)
SEL='(conv_fprop and 8-8-64) or (conv_dgrad and 8-8-64) or (conv_wgrad_wide and 8-8-64 and splits0) or (tile_shapes and 8-8-64) or softmax or normalize or (bn_single and 20) or (halo_fprop and 2-16-8-64) or (halo_dgrad and 2-16-8-64 and not dz) or (halo_wgrad and 16-8-64) or (test_conv_wgrad and 2-8-8-64) or (fused_update_single and dtype0 and opt0) or direct_conv or (gemm_nt and 128-128-64) or (halo_dgrad_dz and 20-13) or staging_ring'
for TOOL in $TOOLS; do
  echo "=== compute-sanitizer --tool $TOOL"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool "$TOOL" --error-exitcode 1 \
    python -m pytest tests/test_native_gpu.py -m gpu -q -x -k "$SEL" --timeout 1200 -p no:cacheprovider \
    > gpurun_out/sanitizer_$TOOL.txt 2>&1
  echo "rc=$?"; tail -4 gpurun_out/sanitizer_$TOOL.txt
done
