#!/bin/bash
# tools/gpu_multi_ab.sh N -- A/B of the bench clock sampler's effect on the multi-GPU step (peer exchange)
N=${1:-2}
mkdir -p gpurun_out
for v in "SFMB200_BENCH_CLOCKS=off" "SFMB200_BENCH_CLOCK_PERIOD_MS=4" "SFMB200_BENCH_CLOCK_PERIOD_MS=1"; do
  name=$(echo $v | tr '=' '_')
  echo "== bench N=$N peer $v"
  env $v SFMB200_EXCHANGE=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/ab_n${N}_$name.json 2> gpurun_out/ab_n${N}_$name.err
  echo "rc=$?"; python - <<PY
import json
d = json.loads(open("gpurun_out/ab_n${N}_$name.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["rep_ms"], d["clocks"])
PY
done
echo "== bench N=1 default"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/ab_n1.json 2> gpurun_out/ab_n1.err; python - <<PY
import json
d = json.loads(open("gpurun_out/ab_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["rep_ms"], d["clocks"])
PY
