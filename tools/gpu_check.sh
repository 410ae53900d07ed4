#!/bin/bash
# tools/gpu_check.sh -- what a gpurun call executes: GPU parity tests, smoke, microbench; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x "$@" > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -40 gpurun_out/pytest_gpu.log
if [ -x tools/_build/red_microbench ]; then echo "== red microbench"; timeout 120 tools/_build/red_microbench > gpurun_out/red_microbench.log 2>&1; cat gpurun_out/red_microbench.log; fi
