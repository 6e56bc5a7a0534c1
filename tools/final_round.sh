#!/bin/bash
# Last gpurun call of the round: the round-end sequence on the default library first, then the split-Q-producer attention
# variants (ab/libvitb200_sq{3,4}.so): phase-offset sweep, parity tests and a bench line of the best point.
set -x
mkdir -p gpurun_out
(timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/final2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final2_tests.log); tail -4 gpurun_out/final2_tests.log
(timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/final2_smoke.log); tail -2 gpurun_out/final2_smoke.log
(timeout 150 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; echo "bench rc=$?")
tail -c 400 gpurun_out/final2_bench.json
timeout 120 python tools/sweep_attn.py $PWD/vit_tensorflow_b200/libvitb200.so:0 $PWD/ab/libvitb200_sq3.so:0,1300,2000,3000,4500,6000 $PWD/ab/libvitb200_sq4.so:0,1300,3000,6000 > gpurun_out/sweep2_attn.log 2>&1
cat gpurun_out/sweep2_attn.log
python - <<'PY' > gpurun_out/best2.env
import json
r = [x for x in json.load(open("gpurun_out/sweep_attn.json")) if "vit_b16" in x and x["lib"] != "libvitb200.so"]
b = min(r, key=lambda x: x["vit_b16"]["ms"])
print(f"export VB_LIB_PATH=$PWD/ab/{b['lib']} VB_ATTN_STAGGER={b['stagger']}")
PY
cat gpurun_out/best2.env
(. gpurun_out/best2.env; timeout 100 python -m pytest tests -x -q -m gpu -k "attention or bf16_vs_oracle or determinism or ragged" > gpurun_out/final2_variant_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final2_variant_tests.log); tail -3 gpurun_out/final2_variant_tests.log
(. gpurun_out/best2.env; timeout 120 python bench.py --no-cpu-baseline > gpurun_out/final2_bench_variant.json 2> gpurun_out/final2_bench_variant.err; echo "bench variant rc=$?")
python - <<'PY'
import json
for f in ("final2_bench", "final2_bench_variant"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], d["clocks"], d["roofline"]["other_kernels"]["attention"])
    except Exception as e:
        print(f, "failed", e)
PY
