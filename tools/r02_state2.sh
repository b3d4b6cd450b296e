#!/bin/bash
# full GPU suite, bench line + breakdown, decode breakdown
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=${1:-v4}
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/r02_pytest_gpu_$V.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r02_pytest_gpu_$V.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --breakdown gpurun_out/r02_step_breakdown_$V.md > gpurun_out/r02_bench_$V.json 2> gpurun_out/r02_bench_$V.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_$V.err; head -c 900 gpurun_out/r02_bench_$V.json; echo; head -34 gpurun_out/r02_step_breakdown_$V.md
timeout 600 python tools/profile_decode.py gpurun_out/r02_decode_breakdown_$V.md | head -40
