#!/bin/bash
# Round-2 first GPU call: baseline test run + the open hardware questions (MMA cost by operand source, softmax-side
# throughput, attention v4 / v5 parity and speed, encoder executor on hardware).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt
timeout 120 ./tools/sm_probe.bin > gpurun_out/r02_sm_probe.md 2>&1
cat gpurun_out/r02_sm_probe.md
timeout 300 python tools/mma_probe.py > gpurun_out/r02_mma_probe.md 2>&1
tail -n 60 gpurun_out/r02_mma_probe.md
VISTA_B200_TEST_ATTN_IMPLS=1,2,3 timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_base.log 2>&1
grep -E "rel-L2|passed|failed|error" gpurun_out/r02_pytest_base.log | tail -n 40
VISTA_B200_TEST_ATTN_IMPLS=4 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention_spatial" > gpurun_out/r02_attn4_tests.log 2>&1
tail -n 5 gpurun_out/r02_attn4_tests.log
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
