#!/bin/bash
# tools/gpu_prof.sh -- gpurun payload: GPU tests, bench, ncu launch list, optional full captures (args: kernel regexes)
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -12 gpurun_out/pytest_gpu.log
echo "== bench ours" ; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err ; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== stage benches"; timeout 900 python tools/bench_stages.py > gpurun_out/bench_stages.json 2> gpurun_out/bench_stages.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_stages.json; tail -3 gpurun_out/bench_stages.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1 ; echo "rc=$?"
for k in "$@"; do
  echo "== ncu full: $k"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_$k -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1 ; echo "rc=$?"
done
ls -la gpurun_out | head -30
