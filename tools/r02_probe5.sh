#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r02_attn5d_bench.txt
VISTA_B200_TEST_ATTN_IMPLS=5 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention_spatial" > gpurun_out/r02_attn5d_tests.log 2>&1
tail -n 3 gpurun_out/r02_attn5d_tests.log
for ch in 1 0; do
for mode in 0 2; do
  echo "v5d chunked=$ch exp=$mode" >> gpurun_out/r02_attn5d_bench.txt
  VB_ATTN5_CHUNKED=$ch VB_ATTN5_EXP=$mode BENCH_ATTN_IMPLS=5 timeout 200 python tools/bench_kernels.py attention >> gpurun_out/r02_attn5d_bench.txt 2>&1
done
done
cat gpurun_out/r02_attn5d_bench.txt
VISTA_B200_TEST_ATTN_IMPLS=3 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gnfuse.log 2>&1
tail -n 15 gpurun_out/r02_pytest_gnfuse.log
python tools/bench_kernels.py groupnorm > gpurun_out/r02_gn_bench.txt 2>&1; cat gpurun_out/r02_gn_bench.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu --no-eager --breakdown gpurun_out/r02_step_breakdown_gnfuse.md > gpurun_out/r02_bench_gnfuse.json 2> gpurun_out/r02_bench_gnfuse.err
tail -c 600 gpurun_out/r02_bench_gnfuse.err; head -c 900 gpurun_out/r02_bench_gnfuse.json; echo; head -20 gpurun_out/r02_step_breakdown_gnfuse.md
