#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VISTA_B200_TEST_ATTN_IMPLS=5 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention_spatial" > gpurun_out/r02_attn5b_tests.log 2>&1
tail -n 5 gpurun_out/r02_attn5b_tests.log
for mode in 0 2; do
  echo "v5b exp=$mode" >> gpurun_out/r02_attn5b_bench.txt
  VB_ATTN5_EXP=$mode BENCH_ATTN_IMPLS=5 timeout 200 python tools/bench_kernels.py attention >> gpurun_out/r02_attn5b_bench.txt 2>&1
done
cat gpurun_out/r02_attn5b_bench.txt
VISTA_B200_ATTN=5 timeout 600 python bench.py --config small --steps 5 --warmup 3 --no-cpu --no-eager > gpurun_out/r02_bench_small.json 2> gpurun_out/r02_bench_small.err
tail -c 1500 gpurun_out/r02_bench_small.err; head -c 1500 gpurun_out/r02_bench_small.json
VISTA_B200_ATTN=5 timeout 900 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r02_step_breakdown_attn5.md > gpurun_out/r02_bench_attn5.json 2> gpurun_out/r02_bench_attn5.err
tail -c 1500 gpurun_out/r02_bench_attn5.err; cat gpurun_out/r02_bench_attn5.json
