"""GPU probe: tcgen05 GEMM vs numpy on bf16-rounded operands, with diagnostics (run under gpurun)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vit_tensorflow_b200 import _lib


def bf16_round(x):
    x = np.ascontiguousarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def gelu(x):
    from scipy.special import erf
    return 0.5 * x * (1 + erf(x / np.sqrt(2)))


def run(M, N, K, bias=False, scale=False, res=False, g=False, iters=0, seed=0):
    rng = np.random.default_rng(seed)
    a = bf16_round(rng.standard_normal((M, K), dtype=np.float32))
    w = bf16_round(rng.standard_normal((K, N), dtype=np.float32) / np.sqrt(K))
    b = rng.standard_normal(N).astype(np.float32) if bias else None
    s = rng.uniform(0.5, 1.5, N).astype(np.float32) if scale else None
    r = bf16_round(rng.standard_normal((M, N), dtype=np.float32)) if res else None
    out, ms = _lib.op_linear(a, w, b, s, r, g, "bf16", iters)
    ref = a.astype(np.float64) @ w.astype(np.float64)
    if bias: ref += b
    if g: ref = gelu(ref)
    if scale: ref *= s
    if res: ref += r
    err = np.abs(out - ref)
    tol = 0.02 + 0.01 * np.abs(ref)
    bad = err > tol
    rec = dict(M=M, N=N, K=K, bias=bias, scale=scale, res=res, gelu=g, max_err=float(err.max()), n_bad=int(bad.sum()), ms=ms)
    if ms:
        rec["tflops"] = 2.0 * M * N * K / ms / 1e9
    if bad.any():
        idx = np.argwhere(bad)
        rec["bad_rows"] = [int(idx[:, 0].min()), int(idx[:, 0].max())]
        rec["bad_cols"] = [int(idx[:, 1].min()), int(idx[:, 1].max())]
        rec["first_bad"] = [[int(i), int(j), float(out[i, j]), float(ref[i, j])] for i, j in idx[:8]]
        rec["bad_per_coltile64"] = np.bincount(idx[:, 1] // 64, minlength=N // 64).tolist()[:16]
        rec["bad_per_rowtile32"] = np.bincount(idx[:, 0] // 32, minlength=(M + 31) // 32).tolist()[:16]
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    recs = []
    try:
        recs.append(run(128, 256, 64))
        recs.append(run(128, 256, 256))
        recs.append(run(256, 512, 768))
        recs.append(run(128, 128, 64))
        recs.append(run(394, 768, 768, bias=True))
        recs.append(run(1000, 384, 384, bias=True, scale=True, res=True))
        recs.append(run(777, 3072, 768, bias=True, g=True))
        recs.append(run(50432, 768, 768, bias=True, res=True, iters=10))
        recs.append(run(50432, 2304, 768, iters=10))
        recs.append(run(50432, 3072, 768, bias=True, g=True, iters=10))
        recs.append(run(50432, 768, 3072, bias=True, res=True, iters=10))
    finally:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(recs, open("gpurun_out/probe_gemm.json", "w"), indent=1)
