#!/bin/bash
# tools/gpu_quick.sh -- gpurun payload: BA parity tests + one bench line per environment variant given as arguments.
mkdir -p gpurun_out
echo "== pytest BA" ; timeout 900 python -m pytest tests/test_gpu_ba.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_ba.log 2>&1 ; echo "pytest rc=$?" ; tail -5 gpurun_out/pytest_ba.log
for variant in default "$@"; do
  name=$(echo $variant | tr '=' '_')
  echo "== bench $variant"
  if [ "$variant" = default ]; then timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  else env $variant timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; fi
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "dense_solve_ms", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["kernels_ms"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/ab_$name.err").read()[-2000:])
PY
done
