#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_glue_gpu.py tests/test_kernels_gpu.py -q -x -s -k "glue or time_mix or rollout or ensemble or layernorm" > gpurun_out/r02_p8_tests.log 2>&1
echo "tests rc=$?"; tail -n 12 gpurun_out/r02_p8_tests.log
for v in 1 0; do echo "VB_LN40=$v"; VB_LN40=$v timeout 200 python tools/bench_kernels.py layernorm; done 2>&1 | tee gpurun_out/r02_p8_ln.txt
for pp in 1 0; do for ch in 1 0; do for mode in 0 1 2 3; do
  echo "v5 pingpong=$pp chunked=$ch exp=$mode"
  VB_ATTN5_PINGPONG=$pp VB_ATTN5_CHUNKED=$ch VB_ATTN5_EXP=$mode BENCH_ATTN_IMPLS=5 timeout 200 python tools/bench_kernels.py attention 2>&1 | grep attention
done; done; done 2>&1 | tee gpurun_out/r02_p8_attn5_matrix.txt
