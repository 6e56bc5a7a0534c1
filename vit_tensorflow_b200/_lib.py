"""ctypes binding of libvitb200.so (include/vitb200.h).  No torch, no numpy-side compute: this module only
marshals pointers and sizes across the C-ABI.  There is NO CPU fallback: if the library is missing it is an
ImportError-style failure, and every entry point fails loudly when no B200 is visible."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VB_LIB_PATH") or os.path.join(_HERE, "libvitb200.so")   # VB_LIB_PATH: developer A/B builds

KIND = {"vit": 0, "deepvit": 1, "cait": 2, "crossvit": 3, "parallel_vit": 4, "patch_merger_vit": 5, "t2t_vit": 6}
PRECISION = {"fp32": 0, "float32": 0, "bf16": 1, "bfloat16": 1}
MEM_HOST, MEM_DEVICE = 0, 1


class VbConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "struct_size", "kind", "precision", "image_h", "image_w", "patch_h", "patch_w", "channels", "num_classes",
        "dim", "depth", "heads", "dim_head", "mlp_dim", "pool", "cls_depth", "max_batch",
        "sm_dim", "lg_dim",
        "sm_patch_size", "sm_enc_depth", "sm_enc_heads", "sm_enc_mlp_dim", "sm_enc_dim_head",
        "lg_patch_size", "lg_enc_depth", "lg_enc_heads", "lg_enc_mlp_dim", "lg_enc_dim_head",
        "cross_attn_depth", "cross_attn_heads", "cross_attn_dim_head", "cross_depth", "parallel_branches",
        "patch_merge_layer_index", "patch_merge_num_tokens",
        "t2t_num_layers", "t2t_k0", "t2t_s0", "t2t_k1", "t2t_s1", "t2t_k2", "t2t_s2", "t2t_k3", "t2t_s3")]


class VbError(RuntimeError):
    pass


_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list EVERY symbol include/vitb200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "vb_abi_version": (C.c_int, []),
    "vb_create": (C.c_int, [C.POINTER(VbConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "vb_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, _i64p, C.c_int32]),
    "vb_num_weights": (C.c_int, [C.c_void_p]),
    "vb_weight_info": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), _i64p, C.POINTER(C.c_int32)]),
    "vb_finalize": (C.c_int, [C.c_void_p]),
    "vb_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vb_forward_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vb_forward_distill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int32, C.c_void_p]),
    "vb_embed_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "vb_forward_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vb_forward_head": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vb_to_patch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vb_patch_to_emb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vb_dp_unique_id": (C.c_int, [C.c_void_p]),
    "vb_dp_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "vb_forward_allgather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vb_last_launch_count": (C.c_int64, [C.c_void_p]),
    "vb_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "vb_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), _i64p, C.c_int32]),
    "vb_last_error": (C.c_char_p, [C.c_void_p]),
    "vb_destroy": (None, [C.c_void_p]),
    "vb_op_linear": (C.c_int, [C.c_int32] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p] + [C.c_int32] * 4 + [_f32p]),
    "vb_op_attention": (C.c_int, [C.c_int32, C.c_int32] + [C.c_void_p] * 8 + [C.c_int32] * 6 + [_f32p]),
    "vb_op_layernorm": (C.c_int, [C.c_int32] + [C.c_void_p] * 4 + [C.c_int32] * 3 + [_f32p]),
    "vb_op_patch_merger": (C.c_int, [C.c_int32] + [C.c_void_p] * 5 + [C.c_int32] * 5 + [_f32p]),
    "vb_op_ln_linear": (C.c_int, [C.c_void_p] * 5 + [C.c_int32, C.c_void_p] + [C.c_int32] * 4 + [_f32p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libvitb200.so (building is `python -m vit_tensorflow_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VbError(f"{LIB_PATH} not found: build it with `python -m vit_tensorflow_b200.build` "
                          "(there is no CPU / PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.vb_abi_version() != 4:
            raise VbError("libvitb200 ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().vb_last_error(handle)
        raise VbError(f"libvitb200 error {rc}: {msg.decode() if msg else '?'}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------------------- single-operator helpers
def op_linear(a, w, bias=None, scale=None, res=None, gelu=False, precision="bf16", iters=0):
    """out = epi(a @ w); returns (out float32 [M,N], ms per launch or None)."""
    a, w, bias, scale, res = map(_f32, (a, w, bias, scale, res))
    M, K = a.shape
    K2, N = w.shape
    assert K == K2
    out = np.empty((M, N), np.float32)
    ms = C.c_float(0)
    check(load().vb_op_linear(PRECISION[precision], _ptr(a), _ptr(w), _ptr(bias), _ptr(scale), _ptr(res), int(bool(gelu)),
                              _ptr(out), M, N, K, iters, C.byref(ms)))
    return out, (ms.value if iters > 0 else None)


def op_attention(q, k, v, heads, variant=0, mix_a=None, mix_b=None, ln_gamma=None, ln_beta=None, precision="bf16", iters=0):
    q, k, v, mix_a, mix_b, ln_gamma, ln_beta = map(_f32, (q, k, v, mix_a, mix_b, ln_gamma, ln_beta))
    B, nq, inner = q.shape
    nk = k.shape[1]
    out = np.empty_like(q)
    ms = C.c_float(0)
    check(load().vb_op_attention(PRECISION[precision], variant, _ptr(q), _ptr(k), _ptr(v), _ptr(mix_a), _ptr(mix_b),
                                 _ptr(ln_gamma), _ptr(ln_beta), _ptr(out), B, nq, nk, heads, inner // heads, iters, C.byref(ms)))
    return out, (ms.value if iters > 0 else None)


def op_layernorm(x, gamma, beta, precision="bf16", iters=0):
    x, gamma, beta = map(_f32, (x, gamma, beta))
    M, D = x.shape
    out = np.empty_like(x)
    ms = C.c_float(0)
    check(load().vb_op_layernorm(PRECISION[precision], _ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), M, D, iters, C.byref(ms)))
    return out, (ms.value if iters > 0 else None)


def op_ln_linear(x, gamma, beta, w, bias=None, gelu=False, iters=0):
    """out = act(LayerNorm(x) @ w + bias) through the LayerNorm-folded tcgen05 GEMM (bf16 engine); returns (out, ms)."""
    x, gamma, beta, w, bias = map(_f32, (x, gamma, beta, w, bias))
    M, K = x.shape
    N = w.shape[1]
    out = np.empty((M, N), np.float32)
    ms = C.c_float(0)
    check(load().vb_op_ln_linear(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(w), _ptr(bias), int(bool(gelu)), _ptr(out), M, N, K, iters,
                                 C.byref(ms)))
    return out, (ms.value if iters > 0 else None)


def op_patch_merger(x, gamma, beta, queries, precision="bf16", iters=0):
    """PatchMerger.call (vit_with_patch_merger.py:49-55): x [B, n, D], queries [nt, D] -> [B, nt, D]."""
    x, gamma, beta, queries = map(_f32, (x, gamma, beta, queries))
    B, n, D = x.shape
    nt = queries.shape[0]
    out = np.empty((B, nt, D), np.float32)
    ms = C.c_float(0)
    check(load().vb_op_patch_merger(PRECISION[precision], _ptr(x), _ptr(gamma), _ptr(beta), _ptr(queries), _ptr(out), B, n, D, nt,
                                    iters, C.byref(ms)))
    return out, (ms.value if iters > 0 else None)
