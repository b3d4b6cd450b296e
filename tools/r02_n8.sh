#!/bin/bash
# 8-GPU bench line with rank 0's per-family breakdown (sharded step: NCCL + host family included) and in-run parity.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-eager --breakdown gpurun_out/r02_step_breakdown_n${N}.md \
  > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
echo "rc=$?"; tail -c 1200 gpurun_out/r02_bench_n${N}.err; cat gpurun_out/r02_bench_n${N}.json; head -60 gpurun_out/r02_step_breakdown_n${N}.md
