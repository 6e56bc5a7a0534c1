"""clock64 timeline of CTA 0 of one GEMM launch (VB_GEMM_TRACE, csrc/gemm_tcgen05.cu `trace`): per-tile budget of the MMA warp
(wait for a free accumulator / k-loop) and of the two epilogue warps of TMEM lane quarter 0 (wait for the accumulator / TMEM
load + store-slab wait / residual wait / math / store issue).

    python tools/gemm_trace.py out_proj fc2 qkv fc1_gelu        # -> gpurun_out/gemm_trace_<name>.txt + a summary on stdout
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from vit_tensorflow_b200 import _lib  # noqa: E402

SHAPES = {"qkv": (2304, 768, 0, 0, 0), "out_proj": (768, 768, 1, 1, 0), "fc1_gelu": (3072, 768, 1, 0, 1), "fc2": (768, 3072, 1, 1, 0)}
M = 50432


def summarize(path):
    ev = [tuple(int(x) for x in l.split()) for l in open(path)]
    out = []
    mma = [(t, c) for r, t, c in ev if r == 0]
    waits, loops = [], []
    for i in range(0, len(mma) - 2, 3):
        if [t for t, _ in mma[i:i + 3]] == [1, 2, 3]:
            waits.append(mma[i + 1][1] - mma[i][1])
            loops.append(mma[i + 2][1] - mma[i + 1][1])
    per_tile = [(mma[i + 3][1] - mma[i][1]) for i in range(0, len(mma) - 5, 3)]
    if per_tile:
        out.append(f"  MMA warp, {len(per_tile)} tiles: period median {np.median(per_tile):.0f} cyc; wait-for-accumulator median {np.median(waits):.0f}, "
                   f"k-loop issue median {np.median(loops):.0f}")
    for role in (1, 2):
        e = [(t, c) for r, t, c in ev if r == role]
        seg = {"wait tfull (10->11)": [], "tmem ld + slab wait (->30)": [], "residual wait / 2nd half ld (30->40)": [],
               "math (40->50)": [], "store issue (50->20)": []}
        prev = None
        for t, c in e:
            if prev is not None:
                pt, pc = prev
                if pt == 10 and t == 11: seg["wait tfull (10->11)"].append(c - pc)
                elif 30 <= t < 40 and (pt == 11 or 20 <= pt < 30): seg["tmem ld + slab wait (->30)"].append(c - pc)
                elif 40 <= t < 50 and 30 <= pt < 40: seg["residual wait / 2nd half ld (30->40)"].append(c - pc)
                elif 50 <= t < 60 and 40 <= pt < 50: seg["math (40->50)"].append(c - pc)
                elif 20 <= t < 30 and 50 <= pt < 60: seg["store issue (50->20)"].append(c - pc)
            prev = (t, c)
        if e:
            out.append(f"  epilogue warp column {role - 1}: " + "; ".join(f"{k} n={len(v)} med {np.median(v):.0f} sum {np.sum(v):.0f}" for k, v in seg.items() if v)
                       + f"; span {e[-1][1] - e[0][1]}")
    return "\n".join(out)


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    rng = np.random.default_rng(0)
    for name in sys.argv[1:]:
        N, K, bias, res, gelu = SHAPES[name]
        a = rng.standard_normal((M, K), dtype=np.float32)
        w = (rng.standard_normal((K, N), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
        b = rng.standard_normal(N).astype(np.float32) if bias else None
        r = rng.standard_normal((M, N), dtype=np.float32) if res else None
        path = f"gpurun_out/gemm_trace_{name}.txt"
        os.environ["VB_GEMM_TRACE"] = path
        _, ms = _lib.op_linear(a, w, b, None, r, gelu, "bf16", 10)
        print(f"{name}: {ms * 1e3:.1f} us ({2.0 * M * N * K / ms / 1e9:.0f} TF/s)")
        print(summarize(path), flush=True)
