#!/bin/bash
# tools/gpu_prof2.sh -- gpurun payload: one bench line, then ONE ncu --set full pass over the second LM iteration's six heavy kernels.
mkdir -p gpurun_out
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/ab_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "dense_solve_ms", "gpu_launches")}, d["e2e"], d["roofline"]["kernels_ms"])
PY
echo "== ncu full (6 kernels of iteration 2)"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:ba_pair_kernel|ba_point_kernel|ba_camera_kernel|ba_backsub_eval|chol_stream|chol_backsolve' -s 6 -c 6 -o gpurun_out/prof_step -f python bench.py --steps 3 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_step.log 2>&1 ; echo "rc=$?"; tail -3 gpurun_out/ncu_full_step.log
ls -la gpurun_out/prof_step.ncu-rep
