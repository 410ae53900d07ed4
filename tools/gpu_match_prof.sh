#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_hamming_tc -s 1 -c 1 -o gpurun_out/prof_match_tc -f python tools/bench_stages.py --images 12 --points 1000 --reps 1 > gpurun_out/ncu_match_tc.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_match_tc.log
timeout 300 python -m pytest tests/test_gpu_triangulate.py -q -m gpu --timeout 120 > gpurun_out/pytest_tri.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_tri.log
timeout 300 python tools/bench_stages.py --images 4 > gpurun_out/bench_stages_small.json 2>&1; tail -1 gpurun_out/bench_stages_small.json | cut -c1-300
