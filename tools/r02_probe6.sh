#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r02_attn5e_bench.txt
VISTA_B200_TEST_ATTN_IMPLS=5 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention_spatial" > gpurun_out/r02_attn5e_tests.log 2>&1
tail -n 3 gpurun_out/r02_attn5e_tests.log
for pp in 1 0; do
for mode in 0 2; do
  echo "v5e pingpong=$pp exp=$mode" >> gpurun_out/r02_attn5e_bench.txt
  VB_ATTN5_PINGPONG=$pp VB_ATTN5_EXP=$mode BENCH_ATTN_IMPLS=5 timeout 200 python tools/bench_kernels.py attention >> gpurun_out/r02_attn5e_bench.txt 2>&1
done
done
cat gpurun_out/r02_attn5e_bench.txt
VISTA_B200_TEST_ATTN_IMPLS=3 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gnfuse.log 2>&1
tail -n 15 gpurun_out/r02_pytest_gnfuse.log
python tools/bench_kernels.py norm > gpurun_out/r02_gn_bench.txt 2>&1; cat gpurun_out/r02_gn_bench.txt
VISTA_B200_ATTN=5 timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu --no-eager --breakdown gpurun_out/r02_step_breakdown_v2.md > gpurun_out/r02_bench_v2.json 2> gpurun_out/r02_bench_v2.err
tail -c 600 gpurun_out/r02_bench_v2.err; head -c 700 gpurun_out/r02_bench_v2.json; echo; head -60 gpurun_out/r02_step_breakdown_v2.md
