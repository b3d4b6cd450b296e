#!/bin/bash
# (the committed captures of round 2 were taken with this script when the long-sequence kernel was attn5; it now names attn7)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn7 -c 1 -o gpurun_out/r02_prof_attn7 -f python tools/ncu_kernels.py attn > gpurun_out/r02_ncu_attn5.log 2>&1
tail -n 3 gpurun_out/r02_ncu_attn5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm -c 4 -o gpurun_out/r02_prof_gemm -f python tools/ncu_kernels.py geglu conv > gpurun_out/r02_ncu_gemm.log 2>&1
tail -n 3 gpurun_out/r02_ncu_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:layernorm -c 1 -o gpurun_out/r02_prof_ln -f python tools/ncu_kernels.py ln > gpurun_out/r02_ncu_ln.log 2>&1
tail -n 3 gpurun_out/r02_ncu_ln.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches.csv python tools/one_step.py > gpurun_out/r02_one_step.log 2>&1
tail -n 2 gpurun_out/r02_one_step.log; wc -l gpurun_out/r02_ncu_launches.csv
