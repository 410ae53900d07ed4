#!/bin/bash
# tools/gpu_bench.sh -- gpurun payload: full GPU test suite, bench lines, ncu launch list + full capture of the top kernel.
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
echo "== bench ours" ; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err ; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== bench cfg2" ; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --workload cfg2 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err ; echo "rc=$?"; cat gpurun_out/bench_cfg2.json
echo "== bench reference" ; timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; echo "rc=$?"; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
if [ "$1" != "noncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1 ; echo "rc=$?"
echo "== ncu full: ba_point_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ba_point_kernel -s 1 -c 2 -o gpurun_out/prof_ba_point -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1 ; echo "rc=$?"
ls -la gpurun_out
fi
