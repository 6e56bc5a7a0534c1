"""Host-side mirror of the reference's model classes over the libvitb200 C-ABI.

    ViT       vit_tensorflow/vit.py:106-177          DeepViT   vit_tensorflow/deepvit.py:112-157
    CaiT      vit_tensorflow/cait.py:155-194         CrossViT  vit_tensorflow/cross_vit.py:232-303

Same constructor kwargs, defaults and assertion messages; `model(img, training=True, **kwargs) -> logits`
with `img` NHWC float32 `[b, H, W, 3]` and logits float32 `[b, num_classes]`.  Everything below the call is
hand-written sm_100a CUDA behind `include/vitb200.h`; this file only validates arguments, owns the weight dict
(Keras layouts) and marshals pointers.  Two extra keyword-only constructor arguments that the
reference does not have: `precision` ("bf16" tcgen05 path, default; "fp32" exact gate path) and `device`.

Semantics notes: inference only -- dropout is the identity, so a non-zero
dropout / emb_dropout / layer_dropout with `training=True` (the reference's default!) cannot be reproduced and
raises unless `training=False` is passed (CaiT's layer_dropout is active even then in the reference,
cait.py:147, so it must be 0).
"""
from __future__ import annotations

import collections
import ctypes as C

import numpy as np

from . import _lib


def pair(t):  # vit.py:11
    return t if isinstance(t, tuple) else (t, t)


def _layerscale_eps(depth):  # cait.py:36-41
    if depth <= 18:
        return 0.1
    if depth <= 24:
        return 1e-5
    return 1e-6


class _Tensorish:
    """Tiny attribute carrier so wrappers can read `model.pos_embedding.shape` (mae.py:33)."""

    def __init__(self, model, name):
        self._m, self._n = model, name

    @property
    def shape(self):
        return self._m.get_weight(self._n).shape

    def numpy(self):
        return self._m.get_weight(self._n)

    def __array__(self, dtype=None, copy=None):
        a = self._m.get_weight(self._n)
        return a.astype(dtype) if dtype is not None else a


class _Transformer:
    """`model.transformer(tokens)` (vit.py:99-104): the entry the reference's wrappers call with any n."""

    def __init__(self, model):
        self._m = model

    def __call__(self, x, training=True):
        return self._m.forward_tokens(x)


class _EngineModel:
    """Common machinery: engine handle, weight dict, forward call."""

    _kind = None

    def _create(self, precision, device, **cfgkw):
        self.precision = precision
        self.device = int(device)
        cfg = _lib.VbConfig()
        cfg.struct_size = C.sizeof(_lib.VbConfig)
        cfg.kind = _lib.KIND[self._kind]
        if precision not in _lib.PRECISION:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISION)}")
        cfg.precision = _lib.PRECISION[precision]
        cfg.channels = 3
        for k, v in cfgkw.items():
            setattr(cfg, k, int(v))
        self._cfg = cfg
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.vb_create(C.byref(cfg), self.device, C.byref(h)))
        self._h = h
        self._finalized = False
        self._specs = collections.OrderedDict()
        name, shape, ndim = C.c_char_p(), (C.c_int64 * 4)(), C.c_int32()
        for i in range(self._lib.vb_num_weights(h)):
            _lib.check(self._lib.vb_weight_info(h, i, C.byref(name), shape, C.byref(ndim)), h)
            self._specs[name.value.decode()] = tuple(int(shape[j]) for j in range(ndim.value))
        self._weights = collections.OrderedDict()

    # ---- weights -------------------------------------------------------------------------------------
    def weight_specs(self):
        return collections.OrderedDict(self._specs)

    def init_weights(self, seed=None):
        """The reference's initial distributions: Dense glorot-uniform / zero bias (Keras defaults),
        LayerNormalization ones/zeros, tf.random.normal Variables N(0,1) (vit.py:146-147, deepvit.py:57,
        cait.py:97-98), LayerScale fill (cait.py:36-44)."""
        rng = np.random.default_rng(seed)
        w = collections.OrderedDict()
        for name, shape in self._specs.items():
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "kernel":
                lim = np.sqrt(6.0 / (shape[0] + shape[1]))
                a = rng.uniform(-lim, lim, size=shape)
            elif leaf in ("bias", "beta"):
                a = np.zeros(shape)
            elif leaf == "gamma":
                a = np.ones(shape)
            elif leaf in ("attn_scale", "ff_scale"):
                layer = int(name.split(".layers.")[1].split(".")[0])
                a = np.full(shape, _layerscale_eps(layer + 1))
            else:  # pos_embedding, cls_token, reattn_weights, mix_pre, mix_post
                a = rng.standard_normal(shape)
            w[name] = a.astype(np.float32)
        self.set_weights_dict(w)

    def set_weights_dict(self, weights):
        """weights: mapping name -> array in the Keras layout.  Missing names keep their value."""
        for name, arr in weights.items():
            if name not in self._specs:
                raise KeyError(f"{type(self).__name__} has no weight named {name!r}")
            a = np.ascontiguousarray(arr, dtype=np.float32)
            if tuple(a.shape) != self._specs[name]:
                raise ValueError(f"weight {name!r}: expected shape {self._specs[name]}, got {tuple(a.shape)}")
            shape = (C.c_int64 * a.ndim)(*a.shape)
            _lib.check(self._lib.vb_set_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim), self._h)
            self._weights[name] = a
        self._finalized = False

    def get_weights_dict(self):
        return collections.OrderedDict((k, v.copy()) for k, v in self._weights.items())

    def get_weight(self, name):
        return self._weights[name]

    def load_weights(self, path):
        with np.load(path) as z:
            self.set_weights_dict({k: z[k] for k in z.files})

    def save_weights(self, path):
        np.savez(path, **self._weights)

    def _finalize(self):
        if not self._finalized:
            _lib.check(self._lib.vb_finalize(self._h), self._h)
            self._finalized = True

    def build(self, input_shape=None):  # Keras API used by mae.py:32; weights exist from construction here
        self._finalize()

    # ---- forward ------------------------------------------------------------------------------------
    def _check_training(self, training):
        if training and any(r != 0 for r in self._dropout_rates):
            raise NotImplementedError(
                "libvitb200 implements inference semantics: stochastic dropout (rate > 0 with training=True, the "
                "reference's default) is not reproducible; pass training=False or construct with dropout = 0")

    def __call__(self, img, training=True, **kwargs):
        """Reference call surface (vit.py:159): numpy NHWC float image batch -> numpy float32 logits."""
        self._check_training(training)
        x = np.ascontiguousarray(img, dtype=np.float32)
        if x.ndim != 4 or x.shape[3] != 3:
            raise ValueError("img must be NHWC float [b, H, W, 3]")
        b, h, w, _ = x.shape
        out = np.empty((b, self.num_classes), np.float32)
        self.forward_raw(x.ctypes.data, _lib.MEM_HOST, b, h, w, out.ctypes.data, _lib.MEM_HOST, None)
        return out

    call = __call__

    def forward_raw(self, img_ptr, img_mem, batch, h, w, logits_ptr, logits_mem, stream=None):
        """Pointer-level forward: host or device (e.g. torch tensor .data_ptr()) buffers, optional cudaStream_t."""
        self._finalize()
        _lib.check(self._lib.vb_forward(self._h, C.c_void_p(img_ptr), img_mem, batch, h, w, C.c_void_p(logits_ptr), logits_mem,
                                        C.c_void_p(stream) if stream else None), self._h)

    def forward_tokens(self, tokens):
        self._finalize()
        x = np.ascontiguousarray(tokens, dtype=np.float32)
        b, n, _ = x.shape
        out = np.empty_like(x)
        _lib.check(self._lib.vb_forward_tokens(self._h, x.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, b, n,
                                               out.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, None), self._h)
        return out

    PROFILE_CLASSES = ("gemm_tcgen05", "attention", "layernorm", "im2col", "other", "gemm_tcgen05_gelu", "gemm_tcgen05_residual")

    def profile(self, on=True):
        """Record CUDA events around every kernel class on the launch stream (for the roofline report)."""
        _lib.check(self._lib.vb_profile_enable(self._h, int(bool(on))), self._h)

    def profile_read(self, reset=True):
        n = len(self.PROFILE_CLASSES)
        ms, fl, by, calls = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
        _lib.check(self._lib.vb_profile_read(self._h, ms, fl, by, calls, int(bool(reset))), self._h)
        return {c: dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=int(calls[i])) for i, c in enumerate(self.PROFILE_CLASSES)}

    @property
    def last_launch_count(self):
        return int(self._lib.vb_last_launch_count(self._h))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.vb_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ViT(_EngineModel):
    """vit.py:106-177."""
    _kind = "vit"

    def __init__(self, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim,
                 pool='cls', dim_head=64, dropout=0.0, emb_dropout=0.0, *, precision="bf16", device=0, seed=None):
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, 'Image dimensions must be divisible by the patch size.'
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'
        self.num_classes, self.pool, self.dim = num_classes, pool, dim
        self._dropout_rates = (dropout, emb_dropout)
        self._create(precision, device, image_h=image_height, image_w=image_width, patch_h=patch_height, patch_w=patch_width,
                     num_classes=num_classes, dim=dim, depth=depth, heads=heads, dim_head=dim_head, mlp_dim=mlp_dim,
                     pool=0 if pool == 'cls' else 1)
        self.init_weights(seed)
        # attribute surface used by the reference's wrappers
        self.pos_embedding = _Tensorish(self, "pos_embedding")
        self.cls_token = _Tensorish(self, "cls_token")
        self.transformer = _Transformer(self)


class ParallelViT(_EngineModel):
    """parallel_vit.py:120-185 (`parallel_vit.ViT`): every layer sums `num_parallel_branches` attention blocks and then as
    many feed-forward blocks, each behind its own LayerNorm (Parallel parallel_vit.py:36-42, Transformer :99-117)."""
    _kind = "parallel_vit"

    def __init__(self, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, pool='cls', num_parallel_branches=2,
                 dim_head=64, dropout=0.0, emb_dropout=0.0, *, precision="bf16", device=0, seed=None):
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, 'Image dimensions must be divisible by the patch size.'
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'
        self.num_classes, self.pool, self.dim = num_classes, pool, dim
        self._dropout_rates = (dropout, emb_dropout)
        self._create(precision, device, image_h=image_height, image_w=image_width, patch_h=patch_height, patch_w=patch_width,
                     num_classes=num_classes, dim=dim, depth=depth, heads=heads, dim_head=dim_head, mlp_dim=mlp_dim,
                     pool=0 if pool == 'cls' else 1, parallel_branches=num_parallel_branches)
        self.init_weights(seed)
        self.pos_embedding = _Tensorish(self, "pos_embedding")
        self.cls_token = _Tensorish(self, "cls_token")


class DistillableViT(ViT):
    """distill.py:47-58 (DistillMixin.call distill.py:16-45): a ViT whose call takes an optional distillation token
    `[1, 1, dim]`; with it the call returns `(logits, distill_tokens [b, dim])`, without it plain ViT logits.
    Forward only -- DistillWrapper's losses (distill.py:100-170) stay with the caller."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.args, self.kwargs = args, kwargs
        self.dim, self.num_classes = kwargs["dim"], kwargs["num_classes"]

    def __call__(self, img, distill_token=None, training=True):
        if distill_token is None:
            return super().__call__(img, training=training)
        self._check_training(training)
        self._finalize()
        x = np.ascontiguousarray(img, dtype=np.float32)
        if x.ndim != 4 or x.shape[3] != 3:
            raise ValueError("img must be NHWC float [b, H, W, 3]")
        tok = np.ascontiguousarray(distill_token, dtype=np.float32).reshape(-1)
        if tok.size != self.dim:
            raise ValueError(f"distill_token must hold dim = {self.dim} values, got shape {np.shape(distill_token)}")
        b, h, w, _ = x.shape
        logits = np.empty((b, self.num_classes), np.float32)
        dist = np.empty((b, self.dim), np.float32)
        _lib.check(self._lib.vb_forward_distill(self._h, x.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, b, h, w,
                                                tok.ctypes.data_as(C.c_void_p), logits.ctypes.data_as(C.c_void_p),
                                                dist.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, None), self._h)
        return logits, dist

    call = __call__


class DeepViT(ViT):
    """deepvit.py:112-157 (integer image/patch sizes only, :117-118)."""
    _kind = "deepvit"

    def __init__(self, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim,
                 pool='cls', dim_head=64, dropout=0.0, emb_dropout=0.0, *, precision="bf16", device=0, seed=None):
        assert image_size % patch_size == 0, 'Image dimensions must be divisible by the patch size.'
        super().__init__(image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, pool, dim_head, dropout,
                         emb_dropout, precision=precision, device=device, seed=seed)


class CaiT(_EngineModel):
    """cait.py:155-194."""
    _kind = "cait"

    def __init__(self, image_size, patch_size, num_classes, dim, depth, cls_depth, heads, mlp_dim,
                 dim_head=64, dropout=0.0, emb_dropout=0.0, layer_dropout=0.0, *, precision="bf16", device=0, seed=None):
        assert image_size % patch_size == 0, 'Image dimensions must be divisible by the patch size.'
        if layer_dropout != 0:
            raise NotImplementedError("layer_dropout drops layers even at inference in the reference (cait.py:147); "
                                      "parity is only defined for layer_dropout = 0")
        self.num_classes, self.dim = num_classes, dim
        self._dropout_rates = (dropout, emb_dropout)
        self._create(precision, device, image_h=image_size, image_w=image_size, patch_h=patch_size, patch_w=patch_size,
                     num_classes=num_classes, dim=dim, depth=depth, cls_depth=cls_depth, heads=heads, dim_head=dim_head,
                     mlp_dim=mlp_dim)
        self.init_weights(seed)
        self.pos_embedding = _Tensorish(self, "pos_embedding")
        self.cls_token = _Tensorish(self, "cls_token")


class CrossViT(_EngineModel):
    """cross_vit.py:232-303."""
    _kind = "crossvit"

    def __init__(self, image_size, num_classes, sm_dim, lg_dim, sm_patch_size=12, sm_enc_depth=1, sm_enc_heads=8,
                 sm_enc_mlp_dim=2048, sm_enc_dim_head=64, lg_patch_size=16, lg_enc_depth=4, lg_enc_heads=8,
                 lg_enc_mlp_dim=2048, lg_enc_dim_head=64, cross_attn_depth=2, cross_attn_heads=8, cross_attn_dim_head=64,
                 depth=3, dropout=0.1, emb_dropout=0.1, *, precision="bf16", device=0, seed=None):
        assert image_size % sm_patch_size == 0, 'Image dimensions must be divisible by the patch size.'
        assert image_size % lg_patch_size == 0, 'Image dimensions must be divisible by the patch size.'
        self.num_classes = num_classes
        self._dropout_rates = (dropout, emb_dropout)
        self._create(precision, device, image_h=image_size, image_w=image_size, num_classes=num_classes, sm_dim=sm_dim,
                     lg_dim=lg_dim, sm_patch_size=sm_patch_size, sm_enc_depth=sm_enc_depth, sm_enc_heads=sm_enc_heads,
                     sm_enc_mlp_dim=sm_enc_mlp_dim, sm_enc_dim_head=sm_enc_dim_head, lg_patch_size=lg_patch_size,
                     lg_enc_depth=lg_enc_depth, lg_enc_heads=lg_enc_heads, lg_enc_mlp_dim=lg_enc_mlp_dim,
                     lg_enc_dim_head=lg_enc_dim_head, cross_attn_depth=cross_attn_depth, cross_attn_heads=cross_attn_heads,
                     cross_attn_dim_head=cross_attn_dim_head, cross_depth=depth)
        self.init_weights(seed)


def from_config(cfg: dict, precision="bf16", device=0, seed=None):
    """Build a model from an oracle-style config dict (kind + reference kwargs)."""
    kw = {k: v for k, v in cfg.items() if k not in ("kind", "channels", "image_h", "image_w", "patch_h", "patch_w", "num_patches")}
    cls = {"vit": ViT, "deepvit": DeepViT, "cait": CaiT, "crossvit": CrossViT, "parallel_vit": ParallelViT}[cfg["kind"]]
    if cfg["kind"] == "crossvit":
        kw.setdefault("dropout", 0.0)
        kw.setdefault("emb_dropout", 0.0)
    return cls(**kw, precision=precision, device=device, seed=seed)
