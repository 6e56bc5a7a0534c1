#!/bin/bash
# Round-2 evidence in one gpurun call: launch list + DRAM traffic of the bench command, full captures of the dominant kernels.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 80 --csv --log-file gpurun_out/launches_vit_b16.csv python bench.py --steps 1 --warmup 5 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:gemm_bf16|attn_" -s 200 -c 100 --csv --log-file gpurun_out/traffic_vit_b16.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/traffic_bench.log 2>&1
for c in deepvit cait_s36; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:attn_" -s 20 -c 12 --csv --log-file gpurun_out/launches_attn_$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$c.log 2>&1
done
bash tools/ncu_r02.sh attention ln_qkv ln_fc1_gelu out_proj fc2 mix_cait mix_deepvit
ls -la gpurun_out | tail -30
