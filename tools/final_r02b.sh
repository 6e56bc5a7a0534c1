#!/bin/bash
# Validation call of round 2's last builds: GPU tests, smoke, default bench line, then extra bench lines while time remains.
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
tail -2 gpurun_out/smoke.txt
timeout 200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_default.json
for c in ${EXTRA_CONFIGS:-vit_l16_384 t2t_readme vit_tiny_gate}; do
  timeout 120 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "$c rc=$?"
done
