set -x
mkdir -p gpurun_out
# 1. correctness of the new attention kernel, default and staggered (n = 577 included), bounded
(timeout 150 python -m pytest tests -m gpu -x -q -k "attention or bf16_vs_oracle or determinism" > gpurun_out/attn_tests_default.log 2>&1; echo "rc=$?" >> gpurun_out/attn_tests_default.log)
tail -3 gpurun_out/attn_tests_default.log
(VB_ATTN_STAGGER=1400 timeout 150 python -m pytest tests -m gpu -x -q -k "attention or bf16_vs_oracle or determinism" > gpurun_out/attn_tests_stagger.log 2>&1; echo "rc=$?" >> gpurun_out/attn_tests_stagger.log)
tail -3 gpurun_out/attn_tests_stagger.log
(VB_LIB_PATH=$PWD/ab/libvitb200_kv4.so VB_ATTN_STAGGER=3000 timeout 150 python -m pytest tests -m gpu -x -q -k "attention or bf16_vs_oracle" > gpurun_out/attn_tests_kv4.log 2>&1; echo "rc=$?" >> gpurun_out/attn_tests_kv4.log)
tail -3 gpurun_out/attn_tests_kv4.log
# 2. sweep
timeout 400 python tools/sweep_attn.py $PWD/vit_tensorflow_b200/libvitb200.so $PWD/ab/libvitb200_kv4.so > gpurun_out/sweep_attn.log 2>&1
cat gpurun_out/sweep_attn.log
# 3. bench A/B on this box: baseline (stagger 0, default lib) vs the best point of the sweep
python - <<'PY' > gpurun_out/best.env
import json
r = [x for x in json.load(open("gpurun_out/sweep_attn.json")) if "vit_b16" in x]
b = min(r, key=lambda x: x["vit_b16"]["ms"])
lib = "vit_tensorflow_b200/libvitb200.so" if b["lib"] == "libvitb200.so" else "ab/" + b["lib"]
print(f"export VB_LIB_PATH=$PWD/{lib} VB_ATTN_STAGGER={b['stagger']}")
PY
cat gpurun_out/best.env
(timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err)
(. gpurun_out/best.env; timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_best.json 2> gpurun_out/bench_best.err)
(timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_base2.json 2> gpurun_out/bench_base2.err)
python - <<'PY'
import json
for f in ("bench_base", "bench_best", "bench_base2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], d["clocks"], d["roofline"]["other_kernels"]["attention"])
    except Exception as e:
        print(f, "failed", e)
PY
