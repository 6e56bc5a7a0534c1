"""A/B of library builds at the ViT-B/16 (B = 256) kernel shapes, one gpurun call for all variants.

    python -m vit_tensorflow_b200.build --variant kv4 VB_ATTN_KV_ST=4          # ab/libvitb200_kv4.so (git-ignored, travels)
    python tools/ab_ops.py vit_tensorflow_b200/libvitb200.so ab/libvitb200_kv4.so [--ops qkv,fc1_gelu,attention] [--iters 30]
                                                                                 [--env VB_FOO=1,2,3]

One subprocess per (library, env value) so that per-process statics (function attributes, env-read knobs) start fresh; every
point also records a CRC of the op's output so that a variant that changes results is visible at once.  Box-to-box and
run-to-run variance on this pool is 2-8 %: only compare numbers from ONE call, and repeat the baseline at the end (--repeat).
Writes gpurun_out/ab_ops.json and prints a table."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {  # name: (N, K, bias, res, gelu) at M = 50432
    "qkv": (2304, 768, 0, 0, 0), "out_proj": (768, 768, 1, 1, 0), "fc1_gelu": (3072, 768, 1, 0, 1), "fc2": (768, 3072, 1, 1, 0),
    "out_proj_cait": (384, 384, 1, 1, 0), "fc2_cait": (384, 1536, 1, 1, 0),          # these two at M = 25088 (CaiT-S36, B = 128)
}   # plus ln_qkv / ln_fc1_gelu (LayerNorm-folded, as the model runs them), ln_*_cait (M = 25088, K = 384), attention, attention_l
CHILD = r"""
import sys, json, zlib
sys.path.insert(0, %(root)r)
import numpy as np
from vit_tensorflow_b200 import _lib
ops, iters, shapes = %(ops)r, %(iters)d, %(shapes)r
rng = np.random.default_rng(0)
out = {}
M = 50432
for name in ops:
    if name.startswith("ln_"):
        N, K, gelu = {"ln_qkv": (2304, 768, 0), "ln_fc1_gelu": (3072, 768, 1), "ln_qkv_cait": (1152, 384, 0), "ln_fc1_cait": (1536, 384, 1)}[name]
        Mx = 25088 if name.endswith("cait") else M
        a = rng.standard_normal((Mx, K), dtype=np.float32)
        w = (rng.standard_normal((K, N), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
        g = rng.uniform(0.5, 1.5, K).astype(np.float32); bt = rng.standard_normal(K).astype(np.float32) * 0.1
        b = rng.standard_normal(N).astype(np.float32)
        o, ms = _lib.op_ln_linear(a, g, bt, w, b, gelu, iters)      # times row_stats_bf16 + the folded GEMM
        out[name] = dict(ms=ms, tflops=2.0 * Mx * N * K / ms / 1e9, crc=zlib.crc32(o.tobytes()))
    elif name in ("mix_cait", "mix_deepvit"):
        B, n, h, dh, var = (128, 196, 8, 48, 2) if name == "mix_cait" else (128, 197, 16, 64, 1)
        q, k, v = (rng.standard_normal((B, n, h * dh), dtype=np.float32) for _ in range(3))
        a = rng.standard_normal((h, h)).astype(np.float32)
        b = rng.standard_normal((h, h)).astype(np.float32) if var == 2 else None
        g = rng.uniform(0.5, 1.5, h).astype(np.float32) if var == 1 else None
        be = rng.standard_normal(h).astype(np.float32) if var == 1 else None
        o, ms = _lib.op_attention(q, k, v, h, var, a, b, g, be, "bf16", iters=iters)
        out[name] = dict(ms=ms, tflops=4.0 * B * h * n * n * dh / ms / 1e9, crc=zlib.crc32(o.tobytes()))
    elif name == "attention":
        B, n, h = 256, 197, 12
        q, k, v = (rng.standard_normal((B, n, h * 64), dtype=np.float32) for _ in range(3))
        o, ms = _lib.op_attention(q, k, v, h, 0, precision="bf16", iters=iters)
        out[name] = dict(ms=ms, tflops=4.0 * B * h * n * n * 64 / ms / 1e9, crc=zlib.crc32(o.tobytes()))
    elif name == "attention_l":
        B, n, h = 32, 577, 16
        q, k, v = (rng.standard_normal((B, n, h * 64), dtype=np.float32) for _ in range(3))
        o, ms = _lib.op_attention(q, k, v, h, 0, precision="bf16", iters=iters)
        out[name] = dict(ms=ms, tflops=4.0 * B * h * n * n * 64 / ms / 1e9, crc=zlib.crc32(o.tobytes()))
    else:
        N, K, bias, res, gelu = shapes[name]
        M = 25088 if name.endswith("cait") else 50432
        a = rng.standard_normal((M, K), dtype=np.float32)
        w = (rng.standard_normal((K, N), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
        b = rng.standard_normal(N).astype(np.float32) if bias else None
        r = rng.standard_normal((M, N), dtype=np.float32) if res else None
        o, ms = _lib.op_linear(a, w, b, None, r, gelu, "bf16", iters)
        out[name] = dict(ms=ms, tflops=2.0 * M * N * K / ms / 1e9, crc=zlib.crc32(o.tobytes()))
print(json.dumps(out))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--ops", default="ln_qkv,out_proj,ln_fc1_gelu,fc2,attention")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--env", default=None, help="NAME=v1,v2,...: sweep one environment variable per library")
    ap.add_argument("--repeat", action="store_true", help="run the first library once more at the end (drift check)")
    args = ap.parse_args()
    ops = [o for o in args.ops.split(",") if o]
    env_name, env_vals = (args.env.split("=")[0], args.env.split("=")[1].split(",")) if args.env else (None, [None])
    libs = list(args.libs) + ([args.libs[0]] if args.repeat else [])
    res = []
    for lib in libs:
        for val in env_vals:
            env = dict(os.environ, VB_LIB_PATH=os.path.abspath(lib))
            if env_name:
                env[env_name] = val
            rec = dict(lib=os.path.basename(lib), **({env_name: val} if env_name else {}))
            try:
                r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, ops=ops, iters=args.iters, shapes=SHAPES)], env=env,
                                   capture_output=True, text=True, timeout=120 + 20 * len(ops))
                if r.returncode != 0 or not r.stdout.strip():
                    raise RuntimeError((r.stderr or "no output").strip().splitlines()[-1])
                rec["ops"] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as e:  # noqa: BLE001
                rec["error"] = repr(e)[:300]
            res.append(rec)
            cells = " ".join(f"{o}={rec['ops'][o]['tflops']:.0f}TF/{rec['ops'][o]['ms'] * 1e3:.1f}us" for o in ops) if "ops" in rec else rec["error"]
            print(f"{rec['lib']:28s} {('' if not env_name else env_name + '=' + str(val)):18s} {cells}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_ops.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
