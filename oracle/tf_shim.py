"""numpy stand-in for the TensorFlow / Keras entry points the reference's hot-path files call.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  TensorFlow cannot be installed in this image, so the
reference (`/root/reference/vit_tensorflow/*.py`, pure Python over `tensorflow` + `einops`) cannot run as shipped.
This module lets it run UNMODIFIED anyway: it puts a module called `tensorflow` (with `tensorflow.keras`,
`tensorflow.keras.layers`, ...) into `sys.modules` for the duration of a `with installed():` block, implementing in
numpy exactly the primitives those files touch.  The reference's own code -- its einsum strings, Rearrange patterns,
concat order, where the scale / LayerScale / residual sits, how `training` threads through -- then executes as
written, and its logits pin `oracle/spec_numpy.py` (tests/test_reference_shim.py, tests/golden/*__refshim.npz).

What this does NOT pin: the semantics of the TensorFlow primitives themselves, restated here from the public API
documentation (third-party, TensorFlow ">= 2.3.0" per the reference's README):
  * `Dense`: `x @ kernel + bias`, kernel `[in, units]`, glorot-uniform / zeros init, built on first call;
  * `LayerNormalization()`: last axis, `epsilon = 1e-3` (the Keras default), population variance, gamma / beta;
  * `Softmax()` / `tf.nn.softmax`: last axis; `Dropout`: identity unless `training` is true;
  * `tf.einsum` (whitespace ignored), `tf.matmul` (batch dimensions broadcast), `tf.split`, `tf.concat`, `tf.transpose`;
  * `tf.image.extract_patches(..., padding='SAME')`: `ceil(in / stride)` positions, total padding
    `max((out - 1) * stride + k - in, 0)` with the smaller half first, zeros outside, patch vector ordered
    (row, column, channel);
  * Keras call plumbing: `Sequential.call` forwards `training` to layers whose `call` names it; a layer called
    without `training` inherits the enclosing call's value, else its own `call` default.
Those are the assumptions that remain "unpinned at the TensorFlow boundary"; everything above them is the reference.

Tensors are plain numpy arrays (so the real `einops.rearrange / repeat` the reference imports work on them);
`tf.Variable` is an ndarray subclass with `.assign()` / `.numpy()`.  `set_dtype(np.float64)` runs the reference in
double precision, which turns the comparison with the float64 spec into a ~1e-12 identity check.
"""
from __future__ import annotations

import collections
import contextlib
import inspect
import math
import sys
import types

import numpy as np

_DTYPE = [np.dtype(np.float32)]
_RNG = [np.random.default_rng(0)]
_TRAINING_CTX: list = []      # `training` of the enclosing Layer.__call__ frames (Keras call context)


def set_dtype(dt) -> None:
    _DTYPE[0] = np.dtype(dt)


def get_dtype():
    return _DTYPE[0]


def set_seed(seed: int) -> None:
    _RNG[0] = np.random.default_rng(seed)


# ------------------------------------------------------------------------------------------------ tensors
class Tensor(np.ndarray):
    """What tf.* functions and layer calls return: an ndarray (einops and numpy treat it as one) that also answers
    `.numpy()`, which the reference's wrappers call on results (`encoded.numpy()[batch_range, masked_indices]`, simmim.py:119)."""

    def numpy(self):
        return self.view(np.ndarray)


def _t(x):
    """ndarray results (and tuples / lists of them) -> Tensor views; everything else unchanged."""
    if isinstance(x, np.ndarray):
        return x if isinstance(x, Tensor) else x.view(Tensor)
    if isinstance(x, (np.generic,)):
        return np.asarray(x).view(Tensor)
    if isinstance(x, tuple) and not hasattr(x, "_fields"):
        return tuple(_t(e) for e in x)
    if isinstance(x, list):
        return [_t(e) for e in x]
    return x


class Variable(Tensor):
    """tf.Variable: owns its storage, `.assign` overwrites in place (same shape, as Keras requires)."""

    def __new__(cls, initial_value=None, trainable=True, name=None, dtype=None, **_):
        a = np.array(initial_value, dtype=dtype or _DTYPE[0], copy=True)
        return a.view(cls)

    def assign(self, value):
        value = np.asarray(value)
        if value.shape != self.shape:
            raise ValueError(f"Variable.assign: shape {value.shape} does not match {self.shape}")
        np.copyto(self, value.astype(self.dtype))
        return self

def _arr(x):
    return x if isinstance(x, np.ndarray) else np.asarray(x, dtype=_DTYPE[0])


def to_numpy(t):
    return np.asarray(t.numpy() if hasattr(t, "numpy") else t).view(np.ndarray)


# ------------------------------------------------------------------------------------------------ tf.* functions
def _einsum(equation, *operands):
    return np.einsum(equation.replace(" ", ""), *[_arr(o) for o in operands])


def _matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _arr(a), _arr(b)
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return np.matmul(a, b)


def _split(value, num_or_size_splits, axis=0):
    value = _arr(value)
    if isinstance(num_or_size_splits, int):
        if value.shape[axis] % num_or_size_splits:
            raise ValueError("tf.split: dimension not evenly divisible")
        return np.split(value, num_or_size_splits, axis=axis)
    return np.split(value, np.cumsum(num_or_size_splits)[:-1], axis=axis)


def _softmax(x, axis=-1):
    x = _arr(x)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def _log_softmax(x, axis=-1):
    x = _arr(x)
    s = x - x.max(axis=axis, keepdims=True)
    return s - np.log(np.exp(s).sum(axis=axis, keepdims=True))


_erf = np.vectorize(math.erf, otypes=[np.float64])


def _erf_op(x):
    x = _arr(x)
    try:
        from scipy.special import erf as sp_erf
        return sp_erf(x).astype(x.dtype)
    except Exception:       # pragma: no cover
        return _erf(x).astype(x.dtype)


def _extract_patches(images, sizes, strides, rates, padding):
    """tf.image.extract_patches, rates 1 (module docstring).  [b, H, W, C] -> [b, oh, ow, kh*kw*C]."""
    x = _arr(images)
    if list(rates) != [1, 1, 1, 1]:
        raise NotImplementedError("extract_patches: rates != 1")
    _, kh, kw, _ = sizes
    _, sh, sw, _ = strides
    b, H, W, C = x.shape
    if padding == "SAME":
        oh, ow = -(-H // sh), -(-W // sw)
        th, tw = max((oh - 1) * sh + kh - H, 0), max((ow - 1) * sw + kw - W, 0)
        x = np.pad(x, ((0, 0), (th // 2, th - th // 2), (tw // 2, tw - tw // 2), (0, 0)))
    elif padding == "VALID":
        oh, ow = (H - kh) // sh + 1, (W - kw) // sw + 1
    else:
        raise ValueError(padding)
    win = np.lib.stride_tricks.sliding_window_view(x, (kh, kw), axis=(1, 2))      # [b, H', W', C, kh, kw]
    win = win[:, ::sh, ::sw][:, :oh, :ow]
    return np.ascontiguousarray(win.transpose(0, 1, 2, 4, 5, 3)).reshape(b, oh, ow, kh * kw * C)


# ------------------------------------------------------------------------------------------------ keras layers
def _call_params(fn):
    try:
        return inspect.signature(fn).parameters
    except (TypeError, ValueError):     # pragma: no cover
        return {}


class Layer:
    """tf.keras.layers.Layer: `__call__` resolves `training` (explicit > enclosing call > `call`'s own default) and
    forwards to `call`."""

    def __init__(self, name=None, **kwargs):
        self.name = name

    def __call__(self, *args, **kwargs):
        params = _call_params(self.call)
        named = "training" in params
        accepts = named or any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())
        value = kwargs.get("training")
        if value is None and _TRAINING_CTX:
            value = _TRAINING_CTX[-1]
        if value is None and named and params["training"].default is not inspect.Parameter.empty:
            value = params["training"].default
        if accepts and value is not None:
            kwargs["training"] = value
        elif "training" in kwargs and (kwargs["training"] is None or not accepts):
            del kwargs["training"]       # keras drops a `training` the layer's call does not take (simmim.py:88 passes it to Rearrange)
        _TRAINING_CTX.append(value)
        try:
            return _t(self.call(*args, **kwargs))
        finally:
            _TRAINING_CTX.pop()

    def call(self, inputs, *args, **kwargs):
        return inputs

    @property
    def weights(self):
        """The layer's own variables in creation order (`patch_to_emb.weights[0].shape[0]`, mae.py:38)."""
        return [v for v in vars(self).values() if isinstance(v, Variable)]


class Model(Layer):
    def build(self, input_shape):
        """keras.Model.build on a subclassed model runs `call` on a placeholder of `input_shape` so that every sub-layer creates
        its variables (mae.py:32, simmim.py:74, mpp.py:149)."""
        self(np.zeros(tuple(input_shape), dtype=_DTYPE[0]))


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self._seq = list(layers) if layers else []

    @property
    def layers(self):
        return self._seq

    def add(self, layer):
        self._seq.append(layer)

    def call(self, inputs, training=None, mask=None):
        x = inputs
        for layer in self._seq:
            kw = {}
            if "training" in _call_params(layer.call):       # keras Sequential.call: only to layers whose call names it
                kw["training"] = training
            x = layer(x, **kw)
        return x


class _Weighted(Layer):
    _order: tuple = ()

    def get_weights(self):
        return [to_numpy(getattr(self, n)) for n in self._order if getattr(self, n, None) is not None]

    def set_weights(self, values):
        names = [n for n in self._order if getattr(self, n, None) is not None]
        if len(values) != len(names):
            raise ValueError(f"{type(self).__name__}.set_weights: expected {len(names)} arrays, got {len(values)}")
        for n, v in zip(names, values):
            getattr(self, n).assign(v)


class Dense(_Weighted):
    _order = ("kernel", "bias")

    def __init__(self, units, activation=None, use_bias=True, name=None, **kwargs):
        super().__init__(name=name)
        if activation is not None:
            raise NotImplementedError("Dense(activation=...) is not used by the reference's hot path")
        self.units, self.use_bias = int(units), bool(use_bias)
        self.kernel = self.bias = None

    def call(self, inputs):
        x = _arr(inputs)
        if self.kernel is None:                                   # build on first call: glorot_uniform / zeros
            fan_in = x.shape[-1]
            lim = math.sqrt(6.0 / (fan_in + self.units))
            self.kernel = Variable(_RNG[0].uniform(-lim, lim, size=(fan_in, self.units)))
            if self.use_bias:
                self.bias = Variable(np.zeros(self.units))
        y = np.matmul(x, self.kernel.view(np.ndarray))
        if self.use_bias:
            y = y + self.bias.view(np.ndarray)
        return y


class LayerNormalization(_Weighted):
    _order = ("gamma", "beta")

    def __init__(self, axis=-1, epsilon=1e-3, center=True, scale=True, name=None, **kwargs):
        super().__init__(name=name)
        if axis != -1 or not center or not scale:
            raise NotImplementedError("LayerNormalization: only the default axis / center / scale")
        self.epsilon = epsilon
        self.gamma = self.beta = None

    def call(self, inputs):
        x = _arr(inputs)
        if self.gamma is None:
            self.gamma = Variable(np.ones(x.shape[-1]))
            self.beta = Variable(np.zeros(x.shape[-1]))
        mean = x.mean(axis=-1, keepdims=True)
        var = np.square(x - mean).mean(axis=-1, keepdims=True)
        return (x - mean) / np.sqrt(var + self.epsilon) * self.gamma.view(np.ndarray) + self.beta.view(np.ndarray)


class Softmax(Layer):
    def __init__(self, axis=-1, name=None, **kwargs):
        super().__init__(name=name)
        self.axis = axis

    def call(self, inputs, mask=None):
        if mask is not None:
            raise NotImplementedError
        return _softmax(inputs, self.axis)


class Activation(Layer):
    def __init__(self, activation, name=None, **kwargs):
        super().__init__(name=name)
        if not callable(activation):
            raise NotImplementedError("Activation: only callables")
        self.activation = activation

    def call(self, inputs):
        return self.activation(inputs)


class Dropout(Layer):
    def __init__(self, rate, name=None, **kwargs):
        super().__init__(name=name)
        self.rate = float(rate)

    def call(self, inputs, training=None):
        if not training or self.rate == 0.0:
            return inputs
        x = _arr(inputs)
        keep = _RNG[0].uniform(size=x.shape) >= self.rate
        return x * keep / (1.0 - self.rate)


class Embedding(Layer):
    """keras.layers.Embedding (mae.py:44): table `[input_dim, output_dim]`, uniform(-0.05, 0.05), built on first call."""

    def __init__(self, input_dim, output_dim, name=None, **kwargs):
        super().__init__(name=name)
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.embeddings = None

    def call(self, inputs):
        if self.embeddings is None:
            self.embeddings = Variable(_RNG[0].uniform(-0.05, 0.05, size=(self.input_dim, self.output_dim)))
        return self.embeddings.view(np.ndarray)[np.asarray(inputs)]


def _top_k(input, k=1, sorted=True, **_):
    x = _arr(input)
    idx = np.argsort(-x, axis=-1, kind="stable")[..., :k].astype(np.int32)
    return _TopKV2(_t(np.take_along_axis(x, idx, axis=-1)), _t(idx))


_TopKV2 = collections.namedtuple("TopKV2", ["values", "indices"])


def _categorical_crossentropy(y_true, y_pred, from_logits=False, **_):
    y_true, y_pred = _arr(y_true), _arr(y_pred)
    logp = _log_softmax(y_pred, -1) if from_logits else np.log(np.clip(y_pred / y_pred.sum(-1, keepdims=True), 1e-7, 1.0))
    return -(y_true * logp).sum(-1)


class _KLDivergence:
    """keras.losses.KLDivergence: sum(y_true * log(y_true / y_pred)) over the last axis, both clipped to [1e-7, 1]."""

    def __init__(self, reduction="auto", name=None):
        self.reduction = reduction

    def __call__(self, y_true, y_pred):
        y_true, y_pred = np.clip(_arr(y_true), 1e-7, 1.0), np.clip(_arr(y_pred), 1e-7, 1.0)
        per = (y_true * np.log(y_true / y_pred)).sum(-1)
        return _t(per if self.reduction == "none" else per.mean())


def _returns_tensor(f):
    def wrapped(*a, **k):
        return _t(f(*a, **k))
    wrapped.__name__ = getattr(f, "__name__", "op")
    return wrapped


# ------------------------------------------------------------------------------------------------ module objects
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__file__ = __file__
    return m


def _build_modules():
    tf = _module(
        "tensorflow", __version__="0.0-numpy-shim", __path__=[],
        Variable=Variable, einsum=_einsum, matmul=_matmul, split=_split,
        concat=lambda values, axis, name=None: np.concatenate([_arr(v) for v in values], axis=axis),
        transpose=lambda a, perm=None, **_: np.transpose(_arr(a), perm),
        cast=lambda x, dtype: np.asarray(x).astype(dtype),
        tanh=lambda x: np.tanh(_arr(x)), pow=lambda x, y: np.power(_arr(x), y),
        reduce_mean=lambda x, axis=None, keepdims=False: _arr(x).mean(axis=axis, keepdims=keepdims),
        reduce_sum=lambda x, axis=None, keepdims=False: _arr(x).sum(axis=axis, keepdims=keepdims),
        identity=lambda x: x, stop_gradient=lambda x: x,
        fill=lambda dims, value: np.full(tuple(dims), value, dtype=_DTYPE[0]),
        argmax=lambda x, axis=None, **_: np.argmax(_arr(x), axis=axis),
        one_hot=lambda idx, depth, axis=-1, **_: np.eye(depth, dtype=_DTYPE[0])[np.asarray(idx)],
        float32=np.float32, float64=np.float64, int32=np.int32, int64=np.int64,
    )
    tf.random = _module(
        "tensorflow.random",
        normal=lambda shape, mean=0.0, stddev=1.0, dtype=None, **_: (mean + stddev * _RNG[0].standard_normal(tuple(shape))).astype(dtype or _DTYPE[0]),
        uniform=lambda shape, minval=0, maxval=None, dtype=None, **_: _uniform(shape, minval, maxval, dtype),
    )
    tf.math = _module("tensorflow.math", erf=_erf_op, tanh=tf.tanh, pow=tf.pow, top_k=_top_k)
    tf.nn = _module("tensorflow.nn", softmax=_softmax, log_softmax=_log_softmax,
                    softmax_cross_entropy_with_logits=lambda labels, logits, axis=-1, **_: -(_arr(labels) * _log_softmax(logits, axis)).sum(axis=axis))
    tf.image = _module("tensorflow.image", extract_patches=_extract_patches)
    # --- beyond the hot path: what the training-loss wrappers (mae.py, simmim.py, mpp.py, distill.py) call around the encoder.
    # Present so that those wrappers RUN over a drop-in encoder in tests/test_wrappers_over_dropin.py; no parity claim rests on them.
    tf.__dict__.update(
        range=lambda start, limit=None, delta=1, dtype=None, **_: np.arange(*((0, start) if limit is None else (start, limit)), delta, dtype=dtype or np.int32),
        argsort=lambda values, axis=-1, direction="ASCENDING", **_: (np.argsort(_arr(values) if direction == "ASCENDING" else -_arr(values), axis=axis, kind="stable")).astype(np.int32),
        square=lambda x, *_, **__: np.square(_arr(x)), abs=lambda x, **_: np.abs(_arr(x)),
        zeros=lambda shape, dtype=None, **_: np.zeros(tuple(shape), dtype=dtype or _DTYPE[0]),
        where=lambda condition, x=None, y=None, **_: np.where(_arr(condition), _arr(x), _arr(y)),
        expand_dims=lambda x, axis, **_: np.expand_dims(_arr(x), axis), reshape=lambda x, shape, **_: np.reshape(_arr(x), tuple(shape)),
        convert_to_tensor=lambda v, dtype=None, **_: np.asarray(v, dtype=dtype or _DTYPE[0]),
        clip_by_value=lambda t, clip_value_min, clip_value_max, **_: np.clip(_arr(t), clip_value_min, clip_value_max),
        reduce_min=lambda x, axis=None, keepdims=False: _arr(x).min(axis=axis, keepdims=keepdims), bool=np.bool_)
    tf.compat = _module("tensorflow.compat", v1=_module("tensorflow.compat.v1", raw_ops=_module(
        "tensorflow.compat.v1.raw_ops", Bucketize=lambda input, boundaries, **_: _t(np.digitize(_arr(input), np.asarray(boundaries), right=False).astype(np.int32)))))
    for mod in (tf, tf.random, tf.math, tf.nn, tf.image):
        for k, f in list(vars(mod).items()):
            if isinstance(f, types.FunctionType):
                setattr(mod, k, _returns_tensor(f))
    losses = _module("tensorflow.keras.losses", categorical_crossentropy=_returns_tensor(_categorical_crossentropy),
                     KLDivergence=_KLDivergence, Reduction=types.SimpleNamespace(NONE="none", SUM="sum", AUTO="auto"))
    layers = _module("tensorflow.keras.layers", Layer=Layer, Dense=Dense, LayerNormalization=LayerNormalization,
                     Softmax=Softmax, Activation=Activation, Dropout=Dropout, Embedding=Embedding)
    keras = _module("tensorflow.keras", __path__=[], Model=Model, Sequential=Sequential, layers=layers, losses=losses)
    tf.keras = keras
    return {"tensorflow": tf, "tensorflow.random": tf.random, "tensorflow.math": tf.math, "tensorflow.nn": tf.nn,
            "tensorflow.image": tf.image, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
            "tensorflow.keras.losses": losses}


def _uniform(shape, minval, maxval, dtype):
    dt = np.dtype(dtype or _DTYPE[0])
    if np.issubdtype(dt, np.integer):
        return _RNG[0].integers(minval, maxval, size=tuple(shape)).astype(dt)
    return _RNG[0].uniform(minval, 1.0 if maxval is None else maxval, size=tuple(shape)).astype(dt)


def _einops_tf_layers():
    """`einops.layers.tensorflow` (Rearrange / Reduce as Keras layers) over the shim's Layer: the real module subclasses
    the real keras Layer; the work itself is einops' own `rearrange` / `reduce` on the numpy arrays either way."""
    import einops

    class Rearrange(Layer):
        def __init__(self, pattern, **axes_lengths):
            super().__init__()
            self.pattern, self.axes_lengths = pattern, axes_lengths

        def call(self, inputs):
            return einops.rearrange(_arr(inputs), self.pattern, **self.axes_lengths)

    class Reduce(Layer):
        def __init__(self, pattern, reduction, **axes_lengths):
            super().__init__()
            self.pattern, self.reduction, self.axes_lengths = pattern, reduction, axes_lengths

        def call(self, inputs):
            return einops.reduce(_arr(inputs), self.pattern, self.reduction, **self.axes_lengths)

    return _module("einops.layers.tensorflow", Rearrange=Rearrange, Reduce=Reduce)


REFERENCE_MODULES = ("vit", "deepvit", "cait", "cross_vit", "t2t", "parallel_vit", "vit_with_patch_merger", "efficient", "distill",
                     "mae", "simmim", "mpp")


@contextlib.contextmanager
def installed(reference_dir=None):
    """Within the block `import tensorflow` resolves to the numpy stand-in and (optionally) `reference_dir` -- the
    directory holding the reference's flat modules (`vit.py`, `t2t.py` does `from vit import Transformer`) -- is first on
    `sys.path`.  On exit every module this put into `sys.modules` (the stand-in and the reference's) is removed again, so
    the rest of the process (e.g. `transformers`, which probes for TensorFlow) never sees a fake `tensorflow`."""
    import einops
    einops.rearrange(np.zeros((1, 2)), "a b -> b a")          # make sure einops' numpy backend is the one initialised
    mods = _build_modules()
    mods["einops.layers.tensorflow"] = _einops_tf_layers()
    names = list(mods) + list(REFERENCE_MODULES)
    saved = {n: sys.modules.get(n) for n in names}
    for n in REFERENCE_MODULES:
        sys.modules.pop(n, None)
    sys.modules.update(mods)
    if reference_dir is not None:
        sys.path.insert(0, reference_dir)
    try:
        yield mods["tensorflow"]
    finally:
        if reference_dir is not None and reference_dir in sys.path:
            sys.path.remove(reference_dir)
        for n in names:
            if saved[n] is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = saved[n]
