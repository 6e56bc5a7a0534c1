#!/bin/bash
# ncu --set full captures of single kernels (tools/ncu_ops.py): raw + source pages as CSV into gpurun_out/
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name, kernel regex, target
  name=$1; rx=$2; shift 2
  timeout 300 $NCU -k "regex:$rx" -s 1 -c 1 -f -o gpurun_out/ncu_$name python tools/ncu_ops.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>> gpurun_out/ncu_$name.log
  ncu -i gpurun_out/ncu_$name.ncu-rep --page source --csv > gpurun_out/ncu_${name}_source.csv 2>> gpurun_out/ncu_$name.log
  rm -f gpurun_out/ncu_$name.ncu-rep
  tail -1 gpurun_out/ncu_$name.log
}
for t in "$@"; do
  case $t in
    mix_cait|mix_deepvit) cap $t attn_mix_kernel $t ;;
    attention) cap $t attn_fwd_kernel $t ;;
    *) cap $t gemm_bf16_kernel $t ;;
  esac
done
