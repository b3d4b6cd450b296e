#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "from vista_b200 import lib; lib.load(); print('library ok')" || exit 9
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/r02_pytest_gpu_v5.log 2>&1
echo "pytest rc=$?"; tail -n 8 gpurun_out/r02_pytest_gpu_v5.log
for L in 2048 512; do echo "ATTN_LONG=$L"; VISTA_B200_ATTN_LONG=$L BENCH_ATTN_IMPLS=3,7 timeout 120 python tools/bench_kernels.py "attention" 2>&1 | grep "seq=576"; done
