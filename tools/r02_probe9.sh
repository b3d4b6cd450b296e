#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VISTA_B200_TEST_ATTN_IMPLS=6 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_spatial" > gpurun_out/r02_p9_tests.log 2>&1
echo "tests rc=$?"; tail -n 8 gpurun_out/r02_p9_tests.log
for pp in 1 0; do for ch in 1 0; do for mode in 0 2; do
  echo "v6 pingpong=$pp chunked=$ch exp=$mode"
  VB_ATTN6_PINGPONG=$pp VB_ATTN6_CHUNKED=$ch VB_ATTN6_EXP=$mode BENCH_ATTN_IMPLS=6 timeout 200 python tools/bench_kernels.py attention 2>&1 | grep attention
done; done; done 2>&1 | tee gpurun_out/r02_p9_attn6_matrix.txt
timeout 300 python -m pytest tests/test_decoder_gpu.py -q -x -s -k "cond_frames" 2>&1 | tail -4
