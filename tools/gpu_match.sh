#!/bin/bash
# tools/gpu_match.sh -- gpurun payload for the matcher: parity tests (short timeouts: a tcgen05 mistake must not hang the box) + stage bench
mkdir -p gpurun_out
echo "== pytest match"; timeout 300 python -m pytest tests/test_gpu_match.py -q -m gpu --timeout 120 -x > gpurun_out/pytest_match.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_match.log
echo "== stage benches"; timeout 600 python tools/bench_stages.py > gpurun_out/bench_stages.json 2> gpurun_out/bench_stages.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_stages.json; tail -3 gpurun_out/bench_stages.err
