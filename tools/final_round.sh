#!/bin/bash
# Last gpurun call of the round: the round-end sequence on the default library first, then the split-Q-producer attention
# variants (ab/libvitb200_sq{3,4}.so): phase-offset sweep, parity tests and a bench line of the best point.
set -x
mkdir -p gpurun_out
(timeout 200 python -m pytest tests -q -m gpu > gpurun_out/final2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final2_tests.log); tail -15 gpurun_out/final2_tests.log
(timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/final2_smoke.log); tail -2 gpurun_out/final2_smoke.log
(timeout 150 python bench.py --no-cpu-baseline > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; echo "bench rc=$?")
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
(. gpurun_out/best2.env; timeout 120 python -m pytest tests -q -m gpu > gpurun_out/final2_variant_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final2_variant_tests.log); tail -8 gpurun_out/final2_variant_tests.log
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
# T2TViT (the reference's usage example, t2t.py:118-131) through the host-buffer call: images/s for DESIGN.md
timeout 60 python - <<'PY' > gpurun_out/t2t_probe.log 2>&1
import time, numpy as np
from vit_tensorflow_b200 import T2TViT
m = T2TViT(dim=512, image_size=224, depth=5, heads=8, mlp_dim=512, num_classes=1000, seed=0)
img = np.random.default_rng(0).standard_normal((16, 224, 224, 3), dtype=np.float32)
m(img); t0 = time.perf_counter(); m(img); m(img); dt = (time.perf_counter() - t0) / 2
print("t2t_vit bf16 B=16:", dt * 1e3, "ms per forward,", 16 / dt, "images/s, launches", m.last_launch_count)
PY
cat gpurun_out/t2t_probe.log
