#!/bin/bash
# tools/gpu_multi.sh N [nccl] -- gpurun --gpus N payload: multi-GPU equivalence tests + weak-scaling bench lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt 2>&1
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 600 > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_multi.log
echo "(N=1 bench skipped here)"
for mode in peer nccl; do
  echo "== bench N=$N exchange=$mode"
  SFMB200_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  echo "rc=$?"; cut -c1-200 gpurun_out/bench_n${N}_$mode.json; tail -3 gpurun_out/bench_n${N}_$mode.err
done
