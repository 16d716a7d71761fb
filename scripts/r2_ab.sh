#!/bin/bash
# A/B of the step under env switches: usage scripts/r2_ab.sh "name:ENV=.. ENV=.." ...
set -u
O=gpurun_out/r2ab; mkdir -p $O
if [ "${RUN_TESTS:-1}" = "1" ]; then
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee $O/pytest.log
fi
for cfg in "$@"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps ${STEPS:-30} --warmup 5 ${BENCH_ARGS:---no-e2e} 2>&1 | grep -E '^\{' | tail -1 > $O/bench_$name.json
  python - "$name" "$envs" $O/bench_$name.json <<'PY'
import sys, json
name, envs, f = sys.argv[1:4]
try:
    d = json.loads(open(f).read())
    print(f"{name:24s} {d['value']:9.1f} img/s {d['ms_per_step']:.3f} ms loss {d.get('final_loss')} launches/step {d.get('gpu_launches',0)//max(d['steps'],1)} [{envs}]")
except Exception as e:
    print(name, "FAILED", e)
PY
done
