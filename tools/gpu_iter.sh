#!/bin/bash
# tools/gpu_iter.sh -- gpurun payload for one tuning iteration: BA parity tests, A/B bench lines, fp64 peaks, launch list.
mkdir -p gpurun_out
echo "== pytest BA" ; timeout 900 python -m pytest tests/test_gpu_ba.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_ba.log 2>&1 ; echo "pytest rc=$?" ; tail -8 gpurun_out/pytest_ba.log
for variant in default "SFMB200_BA_CHOL=steps" "SFMB200_BA_GRID=legacy" $EXTRA_VARIANTS; do
  name=$(echo $variant | tr '=' '_')
  echo "== bench $variant"
  if [ "$variant" = default ]; then timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  else env $variant timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; fi
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["kernels_ms"], d["clocks"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/ab_$name.err").read()[-2000:])
PY
done
if [ -x tools/_build/fp64_peak ]; then echo "== fp64 peak"; timeout 120 tools/_build/fp64_peak | tee gpurun_out/fp64_peak.log; fi
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1 ; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -40
