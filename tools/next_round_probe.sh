#!/bin/bash
# First GPU call of the next round: settles the two open questions of DESIGN.md section 8 in one go.
#   1. what one tcgen05.mma costs by shape / operand source (M = 64 vs 128, A in shared vs tensor memory)
#   2. does the experimental attention v4 (P in tensor memory) pass parity, and is it faster than v3
#   3. does the experimental VAE-encoder executor match the reference fixtures
# Usage (on the GPU box): bash tools/next_round_probe.sh      -> gpurun_out/mma_probe.md, attn4_tests.log, attn4_bench.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/mma_probe.py > gpurun_out/mma_probe.md 2>&1
tail -n 40 gpurun_out/mma_probe.md
VISTA_B200_TEST_ATTN4=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention_spatial" > gpurun_out/attn4_tests.log 2>&1
tail -n 5 gpurun_out/attn4_tests.log
BENCH_ATTN_IMPLS=3,4 timeout 400 python tools/bench_kernels.py attention > gpurun_out/attn4_bench.txt 2>&1
cat gpurun_out/attn4_bench.txt
VISTA_B200_TEST_ENCODER=1 timeout 300 python -m pytest tests/test_decoder_gpu.py -q -s -k "encoder" > gpurun_out/encoder_tests.log 2>&1
tail -n 6 gpurun_out/encoder_tests.log
