"""Measuring stick, not product: cuBLAS (torch.matmul, bf16) at the ViT-B/16 GEMM shapes, and under ncu the tile / cluster
configuration and the L2 -> SM traffic of the kernel it picks.  Used to decide whether the L2 delivery wall of DESIGN.md
section 4.1 is a property of the 256 x 256 pair tile or of the chip.

    python tools/cublas_probe.py            # timing table -> gpurun_out/cublas_probe.json
    ncu --set full -k regex:gemm -s 3 -c 1 ... python tools/cublas_probe.py --one 50432,3072,768
"""
import argparse
import json
import os

import torch

SHAPES = [(50432, 2304, 768), (50432, 768, 768), (50432, 3072, 768), (50432, 768, 3072), (8192, 8192, 8192), (25088, 1536, 384)]


def run(M, N, K, iters):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)      # K-major weight, as the engine stores it
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, w.t(), out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return dict(M=M, N=N, K=K, ms=ms, tflops=2.0 * M * N * K / ms / 1e9)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", default=None)
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    if args.one:
        M, N, K = (int(v) for v in args.one.split(","))
        print(run(M, N, K, 3))
    else:
        res = [run(*s, args.iters) for s in SHAPES]
        for r in res:
            print(f"cuBLAS bf16 {r['M']}x{r['N']}x{r['K']}: {r['ms'] * 1e3:.1f} us  {r['tflops']:.0f} TF/s")
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open("gpurun_out/cublas_probe.json", "w"), indent=1)
