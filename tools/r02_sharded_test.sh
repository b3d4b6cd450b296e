#!/bin/bash
# sharded parity tests on N GPUs (peer-memory collectives included), then the bench line with in-run parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
K=${2:-four}
python -c "from vista_b200 import lib; lib.load(); print('library ok')" || exit 9
timeout 300 python -m pytest tests/test_sharded_gpu.py -q -x -s -k "$K" > gpurun_out/r02_sharded_tests_n$N.log 2>&1
rc=$?
echo "tests rc=$rc"; grep -E "vs-|passed|failed|Error|error|timeout|assert|peer" gpurun_out/r02_sharded_tests_n$N.log | tail -30
if [ "$rc" = "0" ]; then bash tools/r02_n8.sh $N; fi
