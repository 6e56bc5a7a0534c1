"""Sweep the tile phase offset (VB_ATTN_STAGGER, cycles) of the tcgen05 attention kernel at the ViT-B/16 and ViT-L/16-384
shapes, for one or more library builds (VB_LIB_PATH).  One subprocess per point (the offset is read once per process).

    python tools/sweep_attn.py [lib.so ...]         -> gpurun_out/sweep_attn.json
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
for name, (B, n, h) in {"vit_b16": (256, 197, 12), "vit_l16_384": (48, 577, 16)}.items():
    rng = np.random.default_rng(1)
    q, k, v = (rng.standard_normal((B, n, h * 64), dtype=np.float32) for _ in range(3))
    o, ms = _lib.op_attention(q, k, v, h, 0, precision="bf16", iters=30)
    out[name] = dict(ms=ms, tflops=4.0 * B * h * n * n * 64 / ms / 1e9, crc=zlib.crc32(o.tobytes()), finite=bool(np.isfinite(o).all()))
print(json.dumps(out))
""" % ROOT

libs = sys.argv[1:] or [os.path.join(ROOT, "vit_tensorflow_b200", "libvitb200.so")]
staggers = [0, 600, 1000, 1400, 1800, 2400, 3000, 3900, 5000]
res = []
for lib in libs:
    for st in staggers:
        env = dict(os.environ, VB_LIB_PATH=lib, VB_ATTN_STAGGER=str(st))
        try:
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            d = {"error": repr(e)[:200], "stderr": (r.stderr[-300:] if "r" in dir() else "")}
        rec = dict(lib=os.path.basename(lib), stagger=st, **d)
        res.append(rec)
        print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep_attn.json"), "w"), indent=1)
