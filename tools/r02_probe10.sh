#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VISTA_B200_TEST_ATTN_IMPLS=7 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_spatial" > gpurun_out/r02_p10_tests.log 2>&1
echo "tests rc=$?"; tail -n 8 gpurun_out/r02_p10_tests.log
for pp in 1 0; do for ch in 1 0; do for mode in 0 2; do
  echo "v7 pingpong=$pp chunked=$ch exp=$mode"
  VB_ATTN7_PINGPONG=$pp VB_ATTN7_CHUNKED=$ch VB_ATTN7_EXP=$mode BENCH_ATTN_IMPLS=7 timeout 120 python tools/bench_kernels.py attention 2>&1 | grep attention
done; done; done 2>&1 | tee gpurun_out/r02_p10_attn7_matrix.txt
