#!/bin/bash
# Re-entry state check: full GPU test-suite, then the bench line with breakdown (round 2).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r02_step_breakdown_v3.md > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_v3.err; cat gpurun_out/r02_bench_v3.json; head -70 gpurun_out/r02_step_breakdown_v3.md
