#!/bin/bash
# tools/gpu_chol.sh -- gpurun payload: K4 microbench (timeline), BA parity tests, one bench line.
mkdir -p gpurun_out
for nc in 100 180 20; do timeout 120 tools/_build/chol_microbench $nc 20 > gpurun_out/chol_microbench_$nc.log 2>&1; echo "microbench $nc rc=$?"; done
cat gpurun_out/chol_microbench_100.log; grep variant gpurun_out/chol_microbench_180.log gpurun_out/chol_microbench_20.log
echo "== pytest BA" ; timeout 900 python -m pytest tests/test_gpu_ba.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_ba.log 2>&1 ; echo "pytest rc=$?" ; tail -5 gpurun_out/pytest_ba.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/ab_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["kernels_ms"])
PY
