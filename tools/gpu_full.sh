#!/bin/bash
# tools/gpu_full.sh -- full GPU validation: all parity tests, compute-sanitizer on the smoke test, stage benches, matcher profile
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== compute-sanitizer memcheck (smoke)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py --smoke > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
echo "== stage benches"; timeout 600 python tools/bench_stages.py > gpurun_out/bench_stages.json 2> gpurun_out/bench_stages.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_stages.json
echo "== ncu tc matcher"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_hamming_tc -s 1 -c 1 -o gpurun_out/prof_match_tc -f python tools/bench_stages.py --images 12 --points 1000 --reps 1 > gpurun_out/ncu_match_tc.log 2>&1; echo "rc=$?"
echo "== ncu triangulate"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:triangulate_kernel -s 1 -c 1 -o gpurun_out/prof_triangulate -f python tools/bench_stages.py --images 2 --reps 1 > gpurun_out/ncu_tri.log 2>&1; echo "rc=$?"
