#!/bin/bash
# tools/gpu_multi_quick.sh N [notests] -- multi-GPU equivalence tests (2 ranks), then bench at N with both exchange paths
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" != "notests" ]; then
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 600 > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_multi.log
fi
for mode in peer nccl; do
echo "== bench N=$N $mode"
SFMB200_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
echo "rc=$?"; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n${N}_$mode.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["rep_ms"], d["clocks"]["samples"])
PY
done
