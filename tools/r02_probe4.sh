#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r02_attn5c_bench.txt
VISTA_B200_TEST_ATTN_IMPLS=5 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention_spatial" > gpurun_out/r02_attn5c_tests.log 2>&1
tail -n 3 gpurun_out/r02_attn5c_tests.log
for ch in 1 0; do
for mode in 0 2; do
  echo "v5c chunked=$ch exp=$mode" >> gpurun_out/r02_attn5c_bench.txt
  VB_ATTN5_CHUNKED=$ch VB_ATTN5_EXP=$mode BENCH_ATTN_IMPLS=5 timeout 200 python tools/bench_kernels.py attention >> gpurun_out/r02_attn5c_bench.txt 2>&1
done
done
cat gpurun_out/r02_attn5c_bench.txt
