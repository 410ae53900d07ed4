#!/bin/bash
# tools/gpu_prof3.sh -- launch list of a short bench run + ncu --set full of selected kernels (regex in $1) of iteration 2
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -14
if [ -n "$1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$1" -s 4 -c ${2:-2} -o gpurun_out/prof_sel -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_sel.log 2>&1; echo "rc=$?"
fi
