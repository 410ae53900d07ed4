#!/bin/bash
# tools/gpu_final.sh -- gpurun payload at a measurement point: every GPU parity test, smoke, both bench arms, stage benches, the ncu launch
# list of the default bench command and ONE ncu --set full pass over the kernels of an LM iteration (-> profiles/traffic.json).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
echo "== bench ours"; timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
echo "== stage benches"; timeout 600 python tools/bench_stages.py > gpurun_out/bench_stages.json 2> gpurun_out/bench_stages.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_stages.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-stages > gpurun_out/ncu_launch_bench.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -30
echo "== ncu full (kernels of LM iteration 2)"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:ba_pair_kernel|ba_point_kernel|ba_camera_kernel|ba_combine_kernel|ba_backsub_z|chol_stream|chol_backsolve' -s 8 -c 7 -o gpurun_out/prof_step -f python tools/prof_ba.py --iters 4 > gpurun_out/ncu_full_step.log 2>&1 ; echo "rc=$?"; tail -3 gpurun_out/ncu_full_step.log
ls -la gpurun_out/prof_step.ncu-rep
echo "== ORB stage"; timeout 300 python tools/bench_orb.py --images 7 > gpurun_out/bench_orb7.json 2> gpurun_out/bench_orb7.err; echo "rc=$?"; cut -c1-900 gpurun_out/bench_orb7.json
timeout 300 python tools/bench_orb.py --images 50 --no-cpu > gpurun_out/bench_orb50.json 2> gpurun_out/bench_orb50.err; echo "rc=$?"; cut -c1-500 gpurun_out/bench_orb50.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/orb_launches.csv python tools/bench_orb.py --images 7 --reps 2 --no-cpu > gpurun_out/ncu_orb.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/orb_launches.csv 2>/dev/null | head -12
