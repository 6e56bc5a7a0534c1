#!/bin/bash
# One gpurun call: ncu --set full captures of this build's kernels (-> gpurun_out/ncu_*), then the round-end sequence
# (GPU tests, smoke, bench, reference arm) exactly as the driver runs it.   bash tools/ncu_round.sh
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name, kernel regex, skip, count, target args...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  timeout 240 $NCU -k "regex:$rx" -s $skip -c $cnt -f -o gpurun_out/ncu_$name python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>> gpurun_out/ncu_$name.log
  ls -la gpurun_out/ncu_$name.ncu-rep
}
cap attn 'attn_fwd_kernel' 13 1 vit_b16 256 2
ncu -i gpurun_out/ncu_attn.ncu-rep --page source --csv > gpurun_out/ncu_attn_source.csv 2>/dev/null
cap gemm 'gemm_bf16_kernel' 50 4 vit_b16 256 2
cap membound 'im2col_kernel|row_stats_finalize_kernel' 25 2 vit_b16 256 2
cap t2t 'unfold_same_kernel|layernorm_generic_kernel|attn_scores_kernel|attn_softmax_kernel|attn_pv_kernel|gemm_simt_kernel' 0 10 t2t 8 1
rm -f gpurun_out/ncu_membound.ncu-rep gpurun_out/ncu_t2t.ncu-rep
du -sh gpurun_out
# launch list of the bench command (device time per launch, cold and serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_vit_b16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
# ---- the round-end sequence
(timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final_tests.log); tail -3 gpurun_out/final_tests.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/final_smoke.log); tail -2 gpurun_out/final_smoke.log
(timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "rc=$?")
(timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?")
tail -c 600 gpurun_out/final_bench.json
