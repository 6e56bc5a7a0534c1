"""Run one op shape a few times (for ncu captures).  usage: probe_one.py gemm M N K [gelu] [res] | attn B n heads"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vit_tensorflow_b200 import _lib
rng = np.random.default_rng(0)
if sys.argv[1] == "gemm":
    M, N, K = map(int, sys.argv[2:5])
    gelu = "gelu" in sys.argv
    res = "res" in sys.argv
    a = rng.standard_normal((M, K), dtype=np.float32)
    w = (rng.standard_normal((K, N), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) if (gelu or res) else None
    r = rng.standard_normal((M, N), dtype=np.float32) if res else None
    _, ms = _lib.op_linear(a, w, b, None, r, gelu, "bf16", 5)
    print("gemm", M, N, K, "gelu" if gelu else "", "res" if res else "", ms, 2.0 * M * N * K / ms / 1e9, "TF/s")
else:
    B, n, h = map(int, sys.argv[2:5])
    q = rng.standard_normal((B, n, h * 64), dtype=np.float32)
    _, ms = _lib.op_attention(q, q, q, h, 0, precision="bf16", iters=5)
    print("attn", B, n, h, ms, 4.0 * B * h * n * n * 64 / ms / 1e9, "TF/s")
