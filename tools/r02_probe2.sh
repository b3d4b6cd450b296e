#!/bin/bash
# Round-2 second GPU call: new parity goldens on hardware + attention v5 (parity by exp mode, speed vs v3 / v4).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mode in 0 1 2; do
  VB_ATTN5_EXP=$mode VISTA_B200_TEST_ATTN_IMPLS=5 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention_spatial" > gpurun_out/r02_attn5_tests_exp$mode.log 2>&1
  echo "attn5 exp=$mode"; tail -n 5 gpurun_out/r02_attn5_tests_exp$mode.log
done
BENCH_ATTN_IMPLS=3,4 timeout 400 python tools/bench_kernels.py attention > gpurun_out/r02_attn_bench.txt 2>&1
for mode in 0 1 2 3 4; do
  echo "v5 exp=$mode" >> gpurun_out/r02_attn_bench.txt
  VB_ATTN5_EXP=$mode BENCH_ATTN_IMPLS=5 timeout 200 python tools/bench_kernels.py attention >> gpurun_out/r02_attn_bench.txt 2>&1
done
cat gpurun_out/r02_attn_bench.txt
VISTA_B200_TEST_ATTN_IMPLS=3 timeout 900 python -m pytest tests -m gpu -q -s -k "vista_arch or 50_step or encoder" > gpurun_out/r02_pytest_new.log 2>&1
grep -E "rel-L2|passed|failed|rror" gpurun_out/r02_pytest_new.log | tail -n 40
