"""Host-side mirror of the reference's model classes over the libvitb200 C-ABI.

    ViT       vit_tensorflow/vit.py:106-177          DeepViT   vit_tensorflow/deepvit.py:112-157
    CaiT      vit_tensorflow/cait.py:155-194         CrossViT  vit_tensorflow/cross_vit.py:232-303
    parallel_vit.ViT  parallel_vit.py:120-185        DistillableViT  distill.py:47-58 (forward)
    T2TViT    vit_tensorflow/t2t.py:50-116           vit_with_patch_merger.ViT / PatchMerger  vit_with_patch_merger.py:42-55,134-185
    efficient.ViT  vit_tensorflow/efficient.py:12-55 (injected transformer between the engine's embed and head stages)

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


class _Array(np.ndarray):
    """What the host classes hand back: a float32 numpy array that also answers `.numpy()`.  The reference's outputs are eager
    TensorFlow tensors and its wrappers call `.numpy()` on them and on anything derived from them -- `tokens.numpy()[batch_range,
    unmasked_indices]` (mae.py:63), `encoded.numpy()[batch_range, masked_indices]`, `patches.numpy()` (simmim.py:119,125) -- so a
    plain ndarray would break those wrappers; arithmetic on an `_Array` yields an `_Array`, so derived values keep the method."""

    def numpy(self):
        return self.view(np.ndarray)

    def __array_wrap__(self, obj, context=None, return_scalar=False):
        out = super().__array_wrap__(obj, context, return_scalar)
        return out[()] if isinstance(out, np.ndarray) and out.ndim == 0 else out    # reductions give numpy scalars, as on ndarray


def _as_tensor(a):
    return a.view(_Array)


class _Transformer:
    """`model.transformer(tokens)` (vit.py:99-104): the entry the reference's wrappers call with any n."""

    def __init__(self, model):
        self._m = model

    def __call__(self, x, training=True):
        return self._m.forward_tokens(x)


class _Layer:
    """A callable standing where the reference has a Keras layer (`patch_embedding.layers[i]`, `mlp_head`, `dropout`)."""

    def __init__(self, fn):
        self._fn = fn

    def __call__(self, x, training=True):
        return self._fn(x)


class _DenseLayer(_Layer):
    """Stands where the reference has an `nn.Dense`: callable, plus the Keras variable list the wrappers read the patch
    size from -- `patch_to_emb.weights[0].shape[0]` (mae.py:38, simmim.py:80).  `weights` = [kernel [in, units], bias [units]]
    as read-only views of the arrays the model holds (assign through `set_weights_dict`)."""

    def __init__(self, fn, model, name):
        super().__init__(fn)
        self._m, self._name = model, name

    @property
    def kernel(self):
        return self._m._weight_view(self._name + ".kernel")

    @property
    def bias(self):
        return self._m._weight_view(self._name + ".bias")

    @property
    def weights(self):
        return [self.kernel, self.bias]

    def get_weights(self):
        return [w.copy() for w in self.weights]


class _PatchEmbedding(_Layer):
    """`model.patch_embedding` (vit.py:141-144): Sequential([Rearrange, Dense]); the wrappers take `.layers[:2]` apart
    (mae.py:37, simmim.py:79) or call `.layers[-1]` (mpp.py:200)."""

    def __init__(self, model):
        super().__init__(model.forward_patch_embedding)
        self.layers = [_Layer(model.to_patch), _DenseLayer(model.patch_to_emb, model, "patch")]


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

    # `model.pos_embedding` / `model.cls_token` as the reference's wrappers use them: `.shape` (mae.py:33), slicing
    # `encoder.pos_embedding[:, 1:(n + 1)]` (mae.py:54, simmim.py:95), `pos_embedding[:, :(n + 1)]` (mpp.py:208), einops
    # `repeat(transformer.cls_token, ...)` (mpp.py:204) and arithmetic with token arrays: the float32 numpy array the model
    # currently holds for that weight (read-only view; assign through set_weights_dict).
    def _weight_view(self, name):
        if name not in self._specs:
            raise AttributeError(f"{type(self).__name__} has no weight '{name}'")
        a = self._weights[name].view()
        a.flags.writeable = False
        return a.view(_Array)

    @property
    def pos_embedding(self):
        return self._weight_view("pos_embedding")

    @property
    def cls_token(self):
        return self._weight_view("cls_token")

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
        return _as_tensor(out)

    call = __call__

    def forward_raw(self, img_ptr, img_mem, batch, h, w, logits_ptr, logits_mem, stream=None):
        """Pointer-level forward: host or device (e.g. torch tensor .data_ptr()) buffers, optional cudaStream_t."""
        self._finalize()
        _lib.check(self._lib.vb_forward(self._h, C.c_void_p(img_ptr), img_mem, batch, h, w, C.c_void_p(logits_ptr), logits_mem,
                                        C.c_void_p(stream) if stream else None), self._h)

    def forward_tokens(self, tokens):
        self._finalize()
        x = np.ascontiguousarray(tokens, dtype=np.float32)
        if x.ndim != 3 or x.shape[2] != self._cfg.dim:
            raise ValueError(f"transformer(tokens): expected [batch, n, {self._cfg.dim}] tokens, got {x.shape}")
        b, n, _ = x.shape
        out = np.empty_like(x)
        _lib.check(self._lib.vb_forward_tokens(self._h, x.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, b, n,
                                               out.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, None), self._h)
        return _as_tensor(out)

    # ---- the stages of `call` on their own ------------------------------------------
    def _img(self, img):
        x = np.ascontiguousarray(img, dtype=np.float32)
        if x.ndim != 4 or x.shape[3] != 3:
            raise ValueError("img must be NHWC float [b, H, W, 3]")
        return x

    def forward_embed(self, img):
        """`call` up to the transformer (vit.py:160-166): patch embedding, cls token, positions -> [b, rows, dim]."""
        self._finalize()
        x = self._img(img)
        b, h, w, _ = x.shape
        rows = self._lib.vb_embed_rows(self._h, h, w)
        if rows < 0:
            _lib.check(-rows, self._h)
        out = np.empty((b, rows, self._cfg.dim), np.float32)
        _lib.check(self._lib.vb_forward_embed(self._h, x.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, b, h, w,
                                              out.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, None), self._h)
        return _as_tensor(out)

    def forward_head(self, tokens):
        """`call` after the transformer (vit.py:170-175): pooling + mlp_head; [b, n, dim] (or [b, dim]) -> logits."""
        self._finalize()
        x = np.ascontiguousarray(tokens, dtype=np.float32)
        if x.ndim == 2:
            x = x[:, None, :]
        b, n, d = x.shape
        if d != self._cfg.dim:
            raise ValueError(f"tokens must have last dimension {self._cfg.dim}")
        out = np.empty((b, self.num_classes), np.float32)
        _lib.check(self._lib.vb_forward_head(self._h, x.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, b, n,
                                             out.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, None), self._h)
        return _as_tensor(out)

    def to_patch(self, img):
        """`patch_embedding.layers[0]`: Rearrange('b (h p1) (w p2) c -> b (h w) (p1 p2 c)') (vit.py:142)."""
        x = self._img(img)
        b, h, w, c = x.shape
        ph, pw = self._cfg.patch_h, self._cfg.patch_w
        if ph <= 0 or pw <= 0 or h % ph or w % pw:
            raise ValueError("Image dimensions must be divisible by the patch size.")
        out = np.empty((b, (h // ph) * (w // pw), ph * pw * c), np.float32)
        _lib.check(self._lib.vb_to_patch(self._h, x.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, b, h, w,
                                         out.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, None), self._h)
        return _as_tensor(out)

    def patch_to_emb(self, patches):
        """`patch_embedding.layers[1]`: the Dense(dim) on patch vectors [..., patch_dim] (vit.py:143)."""
        self._finalize()
        x = np.ascontiguousarray(patches, dtype=np.float32)
        pd = self._specs["patch.kernel"][0]
        if x.shape[-1] != pd:
            raise ValueError(f"patches must have last dimension {pd}")
        rows = int(np.prod(x.shape[:-1]))
        out = np.empty(x.shape[:-1] + (self._cfg.dim,), np.float32)
        _lib.check(self._lib.vb_patch_to_emb(self._h, x.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, rows,
                                             out.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, None), self._h)
        return _as_tensor(out)

    def forward_patch_embedding(self, img):
        """`model.patch_embedding(img)` (vit.py:160)."""
        if self._kind == "t2t_vit":
            raise NotImplementedError("T2TViT: the tokens-to-token module runs inside forward_embed(img) (with cls token and "
                                      "positions); only patch_embedding.layers[-1] (the Dense) is exposed separately")
        return self.patch_to_emb(self.to_patch(img))

    def _attach_stage_attributes(self):
        """The attribute surface the reference's wrappers use: patch_embedding(.layers), pos_embedding,
        cls_token, dropout, mlp_head."""
        self.patch_embedding = _PatchEmbedding(self)
        self.dropout = _Layer(lambda x: x)               # inference semantics: identity (vit.py:148,166)
        self.mlp_head = _Layer(self.forward_head)

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
        self._attach_stage_attributes()
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
        self._attach_stage_attributes()


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
        return _as_tensor(logits), _as_tensor(dist)

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
        self._attach_stage_attributes()


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


class T2TViT(_EngineModel):
    """t2t.py:50-116.  `transformer=` injection (t2t.py:82-86): pass any callable over [b, n, dim] numpy tokens (e.g. another
    engine model's `.transformer`); then depth / heads / mlp_dim may be omitted and the engine runs the tokens-to-token
    module and the head around it."""
    _kind = "t2t_vit"

    def __init__(self, image_size, num_classes, dim, depth=None, heads=None, mlp_dim=None, pool='cls', channels=3, dim_head=64,
                 dropout=0.0, emb_dropout=0.0, transformer=None, t2t_layers=((7, 4), (3, 2), (3, 2)), *, precision="bf16",
                 device=0, seed=None):
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'
        if transformer is None:
            assert all(v is not None for v in (depth, heads, mlp_dim)), 'depth, heads, and mlp_dim must be supplied'
        else:
            depth, heads, mlp_dim = 0, 1, 1          # the engine holds no main-transformer layers
        if channels != 3:
            raise NotImplementedError("libvitb200 takes NHWC images with 3 channels")
        if isinstance(image_size, tuple):
            raise ValueError("T2TViT takes an integer image_size (t2t.py:66)")
        t2t_layers = tuple((int(k), int(s)) for k, s in t2t_layers)
        if not 1 <= len(t2t_layers) <= 4:
            raise NotImplementedError("libvitb200 supports 1 to 4 t2t_layers")
        self.num_classes, self.pool, self.dim = num_classes, pool, dim
        self._dropout_rates = (dropout, emb_dropout)
        kw = {}
        for i, (k, s) in enumerate(t2t_layers):
            kw[f"t2t_k{i}"], kw[f"t2t_s{i}"] = k, s
        self._create(precision, device, image_h=image_size, image_w=image_size, num_classes=num_classes, dim=dim, depth=depth,
                     heads=heads, dim_head=dim_head, mlp_dim=mlp_dim, pool=0 if pool == 'cls' else 1,
                     t2t_num_layers=len(t2t_layers), **kw)
        self.init_weights(seed)
        self._attach_stage_attributes()
        # t2t.py:58-74: patch_embedding is Sequential([RearrangeUnfoldTransformer..., Dense]); only its last layer (the Dense,
        # what mpp.py:200 calls) is exposed on its own -- the soft splits run inside vb_forward_embed
        self.patch_embedding.layers = [_Layer(self.patch_to_emb)]
        self._injected = transformer
        self.transformer = transformer if transformer is not None else _Transformer(self)

    def __call__(self, img, training=True, **kwargs):
        if self._injected is None:
            return super().__call__(img, training=training)
        self._check_training(training)
        x = self.forward_embed(img)                                   # t2t.py:97-103
        x = np.asarray(self._injected(x, training=training), dtype=np.float32)   # :105
        return self.forward_head(x)                                   # :107-112

    call = __call__


class PatchMergerViT(_EngineModel):
    """vit_with_patch_merger.py:134-185 (`vit_with_patch_merger.ViT`): no cls token, a PatchMerger after layer
    `patch_merge_layer` (default depth // 2) shrinking the stream to `patch_merge_num_tokens` rows, mean pooling."""
    _kind = "patch_merger_vit"

    def __init__(self, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, patch_merge_layer=None,
                 patch_merge_num_tokens=8, dim_head=64, dropout=0.0, emb_dropout=0.0, *, precision="bf16", device=0, seed=None):
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, 'Image dimensions must be divisible by the patch size.'
        self.num_classes, self.dim = num_classes, dim
        self._dropout_rates = (dropout, emb_dropout)
        index = (patch_merge_layer if patch_merge_layer is not None else depth // 2) - 1   # vit_with_patch_merger.py:108
        self._create(precision, device, image_h=image_height, image_w=image_width, patch_h=patch_height, patch_w=patch_width,
                     num_classes=num_classes, dim=dim, depth=depth, heads=heads, dim_head=dim_head, mlp_dim=mlp_dim, pool=1,
                     patch_merge_layer_index=index, patch_merge_num_tokens=patch_merge_num_tokens)
        self.init_weights(seed)
        self._attach_stage_attributes()


class PatchMerger:
    """vit_with_patch_merger.py:42-55 as a standalone layer: `PatchMerger(dim, num_tokens_out)(x [b, n, dim]) -> [b, num_tokens_out, dim]`.
    Weights: `queries` N(0,1) [num_tokens_out, dim] (:47), `norm_gamma` / `norm_beta` (LayerNormalization :46)."""

    def __init__(self, dim, num_tokens_out, *, precision="bf16", seed=None):
        self.dim, self.num_tokens_out, self.precision = dim, num_tokens_out, precision
        self.queries = np.random.default_rng(seed).standard_normal((num_tokens_out, dim)).astype(np.float32)
        self.norm_gamma = np.ones(dim, np.float32)
        self.norm_beta = np.zeros(dim, np.float32)

    def __call__(self, x, training=True):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 3 or x.shape[2] != self.dim:
            raise ValueError(f"x must be [b, n, {self.dim}]")
        return _as_tensor(_lib.op_patch_merger(x, self.norm_gamma, self.norm_beta, self.queries, precision=self.precision)[0])

    call = __call__


class EfficientViT(_EngineModel):
    """efficient.py:12-55 (`efficient.ViT`): the ViT shell around an injected `transformer` -- any callable
    `transformer(tokens [b, n + 1, dim], training=...) -> [b, n + 1, dim]` (the reference passes Keras layers from
    efficient-attention libraries; another engine model's `.transformer` works too).  The engine runs the patch embedding,
    cls token and positions (efficient.py:40-45, `vb_forward_embed`) and pooling + mlp_head (:48-55, `vb_forward_head`)."""
    _kind = "vit"

    def __init__(self, image_size, patch_size, num_classes, dim, transformer, pool='cls', *, precision="bf16", device=0, seed=None):
        image_size_h, image_size_w = pair(image_size)
        assert image_size_h % patch_size == 0 and image_size_w % patch_size == 0, 'image dimensions must be divisible by the patch size'
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'
        self.num_classes, self.pool, self.dim = num_classes, pool, dim
        self._dropout_rates = ()
        self._create(precision, device, image_h=image_size_h, image_w=image_size_w, patch_h=patch_size, patch_w=patch_size,
                     num_classes=num_classes, dim=dim, depth=0, heads=1, dim_head=1, mlp_dim=1, pool=0 if pool == 'cls' else 1)
        self.init_weights(seed)
        self._attach_stage_attributes()
        self.transformer = transformer

    def __call__(self, img, training=True, **kwargs):
        x = self.forward_embed(img)                                              # efficient.py:40-45
        x = np.asarray(self.transformer(x, training=training), dtype=np.float32)  # :46
        return self.forward_head(x)                                              # :48-55

    call = __call__


def from_config(cfg: dict, precision="bf16", device=0, seed=None):
    """Build a model from an oracle-style config dict (kind + reference kwargs)."""
    kw = {k: v for k, v in cfg.items() if k not in ("kind", "channels", "image_h", "image_w", "patch_h", "patch_w", "num_patches",
                                                    "patch_merge_layer_index", "t2t_dims")}
    if cfg["kind"] == "patch_merger_vit":
        kw.pop("pool", None)
    cls = {"vit": ViT, "deepvit": DeepViT, "cait": CaiT, "crossvit": CrossViT, "parallel_vit": ParallelViT,
           "patch_merger_vit": PatchMergerViT, "t2t_vit": T2TViT}[cfg["kind"]]
    if cfg["kind"] == "crossvit":
        kw.setdefault("dropout", 0.0)
        kw.setdefault("emb_dropout", 0.0)
    return cls(**kw, precision=precision, device=device, seed=seed)
