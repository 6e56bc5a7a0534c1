"""GPU probe: per-kernel timings at the BASELINE ViT-B/16 (B=256) shapes, through the C-ABI op entry points."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vit_tensorflow_b200 import _lib

rng = np.random.default_rng(0)
out = {}
M = 50432
for name, (N, K, kw) in {"qkv": (2304, 768, {}), "out_proj": (768, 768, dict(bias=1, res=1)), "fc1_gelu": (3072, 768, dict(bias=1, gelu=1)),
                         "fc1_nogelu": (3072, 768, dict(bias=1)), "fc2": (768, 3072, dict(bias=1, res=1)), "patch": (768, 768, dict(bias=1, res=1))}.items():
    a = rng.standard_normal((M, K), dtype=np.float32)
    w = (rng.standard_normal((K, N), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) if kw.get("bias") else None
    r = rng.standard_normal((M, N), dtype=np.float32) if kw.get("res") else None
    _, ms = _lib.op_linear(a, w, b, None, r, kw.get("gelu", 0), "bf16", 20)
    out[name] = dict(ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
    print(name, out[name], flush=True)
B, n, h, dh = 256, 197, 12, 64
q = rng.standard_normal((B, n, h * dh), dtype=np.float32)
k = rng.standard_normal((B, n, h * dh), dtype=np.float32)
v = rng.standard_normal((B, n, h * dh), dtype=np.float32)
_, ms = _lib.op_attention(q, k, v, h, 0, precision="bf16", iters=20)
out["attention_vitb"] = dict(ms=ms, tflops=4.0 * B * h * n * n * dh / ms / 1e9)
print("attention", out["attention_vitb"], flush=True)
x = rng.standard_normal((M, 768), dtype=np.float32)
_, ms = _lib.op_layernorm(x, np.ones(768, np.float32), np.zeros(768, np.float32), "bf16", 20)
out["layernorm"] = dict(ms=ms, gbps=2 * 2 * M * 768 / ms / 1e6)
print("layernorm", out["layernorm"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_ops.json", "w"), indent=1)
