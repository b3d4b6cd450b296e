#!/bin/bash
# Runs the GPU kernel parity tests group by group (one process per group, bounded by `timeout`),
# so that a trap in one kernel does not hide the state of the others.  Logs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
: > gpurun_out/summary.txt
run() {
  name="$1"; shift
  timeout "${TMO:-420}" python -m pytest -q --no-header -p no:cacheprovider "$@" > "gpurun_out/t_${name}.log" 2>&1
  echo "${name} rc=$?  $(tail -n 1 gpurun_out/t_${name}.log)" >> gpurun_out/summary.txt
}
F=tests/test_kernels_gpu.py
if [ -z "$SKIP_KERNELS" ]; then
run gemm_plain   $F -k "gemm_linear_plain"
run gemm_epi     $F -k "gemm_epilogue or gemm_silu or gemm_geglu"
run gemm_conv    $F -k "gemm_conv3x3"
run gemm_tconv   $F -k "gemm_temporal or downsample"
run attn_sp      $F -k "attention_spatial"
run attn_t       $F -k "attention_temporal"
run norms        $F -k "groupnorm or layernorm"
run small        $F -k "small or upsample or timestep or layout or sampler"
fi
for extra in "$@"; do run "extra_$(basename $extra .py)" "$extra" -s; done
cat gpurun_out/summary.txt
