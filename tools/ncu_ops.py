"""ncu target: single kernels at the BASELINE shapes through the C-ABI op entries (one launch of each, after one warm-up).

    python tools/ncu_ops.py mix_cait | mix_deepvit | attention | ln_qkv | ln_fc1_gelu | out_proj | fc2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from vit_tensorflow_b200 import _lib  # noqa: E402

rng = np.random.default_rng(0)
which = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
M = 50432
for _ in range(reps):
    if which in ("mix_cait", "mix_deepvit"):
        B, n, h, dh, var = (128, 196, 8, 48, 2) if which == "mix_cait" else (128, 197, 16, 64, 1)
        q, k, v = (rng.standard_normal((B, n, h * dh), dtype=np.float32) for _ in range(3))
        a = rng.standard_normal((h, h)).astype(np.float32)
        b = rng.standard_normal((h, h)).astype(np.float32) if var == 2 else None
        g = rng.uniform(0.5, 1.5, h).astype(np.float32) if var == 1 else None
        be = rng.standard_normal(h).astype(np.float32) if var == 1 else None
        o, _ = _lib.op_attention(q, k, v, h, var, a, b, g, be, "bf16")
    elif which == "attention":
        B, n, h = 256, 197, 12
        q, k, v = (rng.standard_normal((B, n, h * 64), dtype=np.float32) for _ in range(3))
        o, _ = _lib.op_attention(q, k, v, h, 0, precision="bf16")
    elif which.startswith("ln_"):
        N, K, gelu = {"ln_qkv": (2304, 768, 0), "ln_fc1_gelu": (3072, 768, 1)}[which]
        x = rng.standard_normal((M, K), dtype=np.float32)
        w = (rng.standard_normal((K, N), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
        o, _ = _lib.op_ln_linear(x, np.ones(K, np.float32), np.zeros(K, np.float32), w, np.zeros(N, np.float32), gelu)
    else:
        N, K = {"out_proj": (768, 768), "fc2": (768, 3072)}[which]
        x = rng.standard_normal((M, K), dtype=np.float32)
        w = (rng.standard_normal((K, N), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
        r = rng.standard_normal((M, N), dtype=np.float32)
        o, _ = _lib.op_linear(x, w, np.zeros(N, np.float32), None, r, 0, "bf16")
print(which, o.shape, float(np.abs(o).mean()))
