#!/bin/bash
# tools/gpu_orb.sh -- gpurun payload for the ORB stage: stage-by-stage debug table, parity tests, memcheck, stage bench, launch list
mkdir -p gpurun_out
echo "== debug"; timeout 600 python tools/debug_orb.py > gpurun_out/orb_debug.log 2>&1; echo "rc=$?"; cut -c1-400 gpurun_out/orb_debug.log | tail -16
echo "== pytest orb"; timeout 900 python -m pytest tests/test_gpu_orb.py -q -m gpu --timeout 300 > gpurun_out/pytest_orb.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_orb.log
echo "== bench"; timeout 600 python tools/bench_orb.py --images 7 > gpurun_out/bench_orb7.json 2> gpurun_out/bench_orb7.err; echo "rc=$?"; cat gpurun_out/bench_orb7.json; tail -3 gpurun_out/bench_orb7.err
timeout 600 python tools/bench_orb.py --images 50 --no-cpu > gpurun_out/bench_orb50.json 2> gpurun_out/bench_orb50.err; echo "rc=$?"; cat gpurun_out/bench_orb50.json
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/orb_launches.csv python tools/bench_orb.py --images 7 --reps 2 --no-cpu > gpurun_out/ncu_orb.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/orb_launches.csv 2>/dev/null | head -16
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/bench_orb.py --images 2 --reps 1 --no-cpu > gpurun_out/orb_memcheck.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/orb_memcheck.log
