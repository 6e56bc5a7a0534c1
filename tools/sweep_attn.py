"""Sweep the tile phase offset (VB_ATTN_STAGGER, cycles) of the tcgen05 attention kernel at the ViT-B/16 shape for the
split-Q-producer variant builds (`python -m vit_tensorflow_b200.build --variant sq3 VB_ATTN_SPLIT_Q=1`), next to the default
library.  One subprocess per point (the offset is read once per process).

    python tools/sweep_attn.py lib.so[:offset,offset,...] ...         -> gpurun_out/sweep_attn.json
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, os, json, zlib
sys.path.insert(0, %r)
import numpy as np
from vit_tensorflow_b200 import _lib
out = {}
for name, (B, n, h) in {"vit_b16": (256, 197, 12)}.items():
    rng = np.random.default_rng(1)
    q, k, v = (rng.standard_normal((B, n, h * 64), dtype=np.float32) for _ in range(3))
    o, ms = _lib.op_attention(q, k, v, h, 0, precision="bf16", iters=25)
    out[name] = dict(ms=ms, tflops=4.0 * B * h * n * n * 64 / ms / 1e9, crc=zlib.crc32(o.tobytes()), finite=bool(np.isfinite(o).all()))
print(json.dumps(out))
""" % ROOT

specs = sys.argv[1:] or [os.path.join(ROOT, "vit_tensorflow_b200", "libvitb200.so") + ":0"]
res = []
for spec in specs:
    lib, _, offs = spec.partition(":")
    for st in [int(x) for x in (offs or "0").split(",")]:
        env = dict(os.environ, VB_LIB_PATH=lib, VB_ATTN_STAGGER=str(st))
        try:
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=30)
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            d = {"error": repr(e)[:200], "stderr": (r.stderr[-300:] if "r" in dir() else "")}
        rec = dict(lib=os.path.basename(lib), stagger=st, **d)
        res.append(rec)
        print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep_attn.json"), "w"), indent=1)
