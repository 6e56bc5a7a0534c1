"""numpy "spec" restatement of the reference forward pass (float64 by default).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED at the TensorFlow
boundary (no TF in this image, the reference holds no golden vectors).

Each function cites the reference lines it follows.  Keras/TF semantics:
Dense = x @ kernel[in,out] + bias; LayerNormalization eps = 1e-3, biased variance;
Softmax axis -1; GELU exact erf form; Dropout is identity (parity is defined for rate 0).
einops itself (installed) performs the patch Rearrange, so the patch-vector order is the
reference's by construction.
"""
from __future__ import annotations

import math
import numpy as np
from einops import rearrange

try:  # scipy is in the image; math.erf fallback keeps the spec importable without it
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

LN_EPS = 1e-3  # tf.keras.layers.LayerNormalization default epsilon (vit.py:18)


# ----------------------------------------------------------------------------- primitives
def dense(x, w, name, bias=True):
    y = x @ w[name + ".kernel"]
    if bias:
        y = y + w[name + ".bias"]
    return y


def layer_norm(x, w, name):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + LN_EPS) * w[name + ".gamma"] + w[name + ".beta"]


def softmax(x):
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=-1, keepdims=True)


def gelu(x):
    # vit.py:29-34, approximate=False branch
    return 0.5 * x * (1.0 + _erf(x / 1.4142135623730951))


def mlp(x, w, pre):
    # vit.py:38-44 (cait.py:72-79, cross_vit.py:40-47, deepvit.py identical)
    return dense(gelu(dense(x, w, pre + "fc1")), w, pre + "fc2")


def patch_embed(img, w, name, ph, pw):
    # vit.py:141-144: Rearrange('b (h p1) (w p2) c -> b (h w) (p1 p2 c)') + Dense(dim)
    x = rearrange(img, 'b (h p1) (w p2) c -> b (h w) (p1 p2 c)', p1=ph, p2=pw)
    return dense(x, w, name)


def _split_heads(t, h):
    return rearrange(t, 'b n (h d) -> b h n d', h=h)


# ----------------------------------------------------------------------------- attention variants
def attention_vit(x, w, pre, heads, dim_head, *, deepvit=False):
    """vit.py:71-85 / deepvit.py:73-91."""
    qkv = dense(x, w, pre + "to_qkv", bias=False)
    q, k, v = (_split_heads(t, heads) for t in np.split(qkv, 3, axis=-1))
    dots = np.einsum('bhid,bhjd->bhij', q, k) * dim_head ** -0.5
    attn = softmax(dots)
    if deepvit:
        # deepvit.py:83-84: head mix then LayerNorm across the head axis
        attn = np.einsum('bhij,hg->bgij', attn, w[pre + "reattn_weights"])
        a = rearrange(attn, 'b h i j -> b i j h')
        a = layer_norm(a, w, pre + "reattn_norm")
        attn = rearrange(a, 'b i j h -> b h i j')
    out = np.einsum('bhij,bhjd->bhid', attn, v)
    out = rearrange(out, 'b h n d -> b n (h d)')
    if (pre + "to_out.kernel") in w:  # absent iff heads==1 and dim_head==dim (vit.py:53)
        out = dense(out, w, pre + "to_out")
    return out


def attention_qkv(x, w, pre, heads, dim_head, *, context=None, talking_heads=False):
    """cait.py:107-131 (talking_heads=True; context concatenated after x, :109-112) and
    cross_vit.py:69-93 (kv_include_self=True is the same concat, :75-76)."""
    ctx = x if context is None else np.concatenate([x, context], axis=1)
    q = dense(x, w, pre + "to_q", bias=False)
    kv = dense(ctx, w, pre + "to_kv", bias=False)
    k, v = np.split(kv, 2, axis=-1)
    q, k, v = (_split_heads(t, heads) for t in (q, k, v))
    dots = np.einsum('bhid,bhjd->bhij', q, k) * dim_head ** -0.5
    if talking_heads:
        dots = np.einsum('bhij,hg->bgij', dots, w[pre + "mix_pre"])      # cait.py:123
    attn = softmax(dots)
    if talking_heads:
        attn = np.einsum('bhij,hg->bgij', attn, w[pre + "mix_post"])     # cait.py:125
    out = np.einsum('bhij,bhjd->bhid', attn, v)
    out = rearrange(out, 'b h n d -> b n (h d)')
    return dense(out, w, pre + "to_out")


# ----------------------------------------------------------------------------- models
def transformer_vit(x, w, cfg, prefix="layers."):
    """vit.py:99-104 / deepvit.py:105-110 -- also the entry the L3 wrappers use with any n."""
    deep = cfg["kind"] == "deepvit"
    for L in range(cfg["depth"]):
        pre = f"{prefix}{L}."
        x = attention_vit(layer_norm(x, w, pre + "attn_norm"), w, pre, cfg["heads"], cfg["dim_head"], deepvit=deep) + x
        x = mlp(layer_norm(x, w, pre + "ff_norm"), w, pre) + x
    return x


def transformer_parallel(x, w, cfg):
    """parallel_vit.py:114-117 with Parallel :41-42: every branch has its own PreNorm; branch outputs are summed."""
    for L in range(cfg["depth"]):
        pres = [f"layers.{L}.branch{i}." for i in range(cfg["num_parallel_branches"])]
        x = sum(attention_vit(layer_norm(x, w, p + "attn_norm"), w, p, cfg["heads"], cfg["dim_head"]) for p in pres) + x
        x = sum(mlp(layer_norm(x, w, p + "ff_norm"), w, p) for p in pres) + x
    return x


def forward_vit(img, w, cfg):
    """vit.py:159-177 / deepvit.py:139-157 / parallel_vit.py:167-185."""
    x = patch_embed(img, w, "patch", cfg["patch_h"], cfg["patch_w"])
    b, n, _ = x.shape
    cls = np.broadcast_to(w["cls_token"], (b, 1, x.shape[-1]))
    x = np.concatenate([cls, x], axis=1)
    x = x + w["pos_embedding"][:, :n + 1]
    x = transformer_parallel(x, w, cfg) if cfg["kind"] == "parallel_vit" else transformer_vit(x, w, cfg)
    x = x.mean(axis=1) if cfg["pool"] == "mean" else x[:, 0]
    return dense(layer_norm(x, w, "head_norm"), w, "head")


def transformer_cait(x, w, cfg, stack, depth, context=None):
    """cait.py:146-153 with layer_dropout == 0 (cait.py:18-19)."""
    for L in range(depth):
        pre = f"{stack}.layers.{L}."
        a = attention_qkv(layer_norm(x, w, pre + "attn_norm"), w, pre, cfg["heads"], cfg["dim_head"],
                          context=context, talking_heads=True)
        x = a * w[pre + "attn_scale"] + x                                # LayerScale cait.py:48, residual :150
        x = mlp(layer_norm(x, w, pre + "ff_norm"), w, pre) * w[pre + "ff_scale"] + x
    return x


def forward_cait(img, w, cfg):
    """cait.py:180-194."""
    x = patch_embed(img, w, "patch", cfg["patch_h"], cfg["patch_w"])
    b, n, d = x.shape
    x = x + w["pos_embedding"][:, :n]
    x = transformer_cait(x, w, cfg, "patch_transformer", cfg["depth"])
    cls = np.broadcast_to(w["cls_token"], (b, 1, d))
    x = transformer_cait(cls, w, cfg, "cls_transformer", cfg["cls_depth"], context=x)
    return dense(layer_norm(x[:, 0], w, "head_norm"), w, "head")


def _crossvit_encoder(x, w, cfg, pre, br):
    """cross_vit.py:108-115 (Transformer with trailing LayerNorm)."""
    h, dh = cfg[f"{br}_enc_heads"], cfg[f"{br}_enc_dim_head"]
    for L in range(cfg[f"{br}_enc_depth"]):
        p = f"{pre}layers.{L}."
        x = attention_qkv(layer_norm(x, w, p + "attn_norm"), w, p, h, dh) + x
        x = mlp(layer_norm(x, w, p + "ff_norm"), w, p) + x
    return layer_norm(x, w, pre + "final_norm")


def _cross_attend(cls, ctx, w, cfg, pre):
    """ProjectInOut(PreNorm(Attention)) cross_vit.py:128-138,24-25,69-93 (+ residual :159-160)."""
    x = cls
    proj = (pre + "project_in.kernel") in w
    if proj:
        x = dense(x, w, pre + "project_in")
    x = attention_qkv(layer_norm(x, w, pre + "norm"), w, pre, cfg["cross_attn_heads"],
                      cfg["cross_attn_dim_head"], context=ctx)
    if proj:
        x = dense(x, w, pre + "project_out")
    return x + cls


def forward_crossvit(img, w, cfg):
    """cross_vit.py:290-303."""
    toks = {}
    for br in ("sm", "lg"):
        p = cfg[f"{br}_patch_size"]
        x = patch_embed(img, w, f"{br}_embed.patch", p, p)                # cross_vit.py:219-229
        b, n, d = x.shape
        cls = np.broadcast_to(w[f"{br}_embed.cls_token"], (b, 1, d))
        x = np.concatenate([cls, x], axis=1) + w[f"{br}_embed.pos_embedding"][:, :n + 1]
        toks[br] = x
    sm, lg = toks["sm"], toks["lg"]
    for D in range(cfg["depth"]):                                         # cross_vit.py:190-196
        sm = _crossvit_encoder(sm, w, cfg, f"blocks.{D}.sm_enc.", "sm")
        lg = _crossvit_encoder(lg, w, cfg, f"blocks.{D}.lg_enc.", "lg")
        sm_cls, sm_p, lg_cls, lg_p = sm[:, :1], sm[:, 1:], lg[:, :1], lg[:, 1:]   # :154
        for R in range(cfg["cross_attn_depth"]):                          # :156-158
            sm_cls = _cross_attend(sm_cls, lg_p, w, cfg, f"blocks.{D}.cross.{R}.sm_attend_lg.")
            lg_cls = _cross_attend(lg_cls, sm_p, w, cfg, f"blocks.{D}.cross.{R}.lg_attend_sm.")
        sm = np.concatenate([sm_cls, sm_p], axis=1)
        lg = np.concatenate([lg_cls, lg_p], axis=1)
    sm_logits = dense(layer_norm(sm[:, 0], w, "sm_head_norm"), w, "sm_head")
    lg_logits = dense(layer_norm(lg[:, 0], w, "lg_head_norm"), w, "lg_head")
    return sm_logits + lg_logits                                          # :301


def patch_merger(x, w, pre="patch_merger."):
    """PatchMerger.call vit_with_patch_merger.py:49-55: LN -> sim = queries . (x^T * dim^-0.5) -> softmax over the tokens ->
    attn . x; [b, n, dim] -> [b, num_tokens_out, dim]."""
    d = x.shape[-1]
    x = layer_norm(x, w, pre + "norm")                                                      # :50
    sim = np.einsum('td,bnd->btn', w[pre + "queries"], x * d ** -0.5)                       # :51 (scale :45)
    return np.einsum('btn,bnd->btd', softmax(sim), x)                                       # :52-53


def forward_patch_merger_vit(img, w, cfg):
    """vit_with_patch_merger.py:174-185 (ViT.call) with Transformer.call :118-126."""
    x = patch_embed(img, w, "patch", cfg["patch_h"], cfg["patch_w"])                        # :175
    n = x.shape[1]
    x = x + w["pos_embedding"][:, :n]                                                       # :178 (no cls token)
    for L in range(cfg["depth"]):
        pre = f"layers.{L}."
        x = attention_vit(layer_norm(x, w, pre + "attn_norm"), w, pre, cfg["heads"], cfg["dim_head"]) + x   # :120
        x = mlp(layer_norm(x, w, pre + "ff_norm"), w, pre) + x                              # :121
        if L == cfg["patch_merge_layer_index"]:                                             # :123-124
            x = patch_merger(x, w)
    x = x.mean(axis=1)                                                                      # Reduce('b n d -> b d', 'mean') :169
    return dense(layer_norm(x, w, "head_norm"), w, "head")


def extract_patches_same(x, k, stride):
    """tf.image.extract_patches(x, sizes=[1,k,k,1], strides=[1,s,s,1], rates=[1,1,1,1], padding='SAME') (t2t.py:43).
    TF semantics (third-party, restated): out = ceil(in / s); total padding max((out-1)*s + k - in, 0) with the smaller
    half BEFORE; out-of-image taps read 0; the patch vector is ordered (k_row, k_col, channel) with channel fastest.
    x [b, H, W, C] -> [b, oh, ow, k*k*C]."""
    b, H, W, C = x.shape
    oh, ow = -(-H // stride), -(-W // stride)
    ph, pw = max((oh - 1) * stride + k - H, 0), max((ow - 1) * stride + k - W, 0)
    xp = np.zeros((b, H + ph, W + pw, C), x.dtype)
    xp[:, ph // 2:ph // 2 + H, pw // 2:pw // 2 + W] = x
    out = np.empty((b, oh, ow, k, k, C), x.dtype)
    for i in range(k):
        for j in range(k):
            out[:, :, :, i, j] = xp[:, i:i + (oh - 1) * stride + 1:stride, j:j + (ow - 1) * stride + 1:stride]
    return out.reshape(b, oh, ow, k * k * C)


def t2t_tokens(img, w, cfg):
    """T2TViT.patch_embedding (t2t.py:58-74): per soft-split layer RearrangeUnfoldTransformer.call (:39-48), then Dense(dim)."""
    x = img
    last = len(cfg["t2t_layers"]) - 1
    for i, (k, st) in enumerate(cfg["t2t_layers"]):
        if i > 0:                                                                           # :40-41
            hh = int(math.sqrt(x.shape[1]))
            x = rearrange(x, 'b (h w) c -> b h w c', h=hh)
        x = extract_patches_same(x, k, st)                                                  # :43
        x = rearrange(x, 'b h w c -> b (h w) c')                                            # :44
        if i != last:                                                                       # :45-46: Transformer(dim, heads=1, depth=1, dim_head=dim, mlp_dim=dim)
            d = x.shape[-1]
            sub = dict(kind="vit", depth=1, heads=1, dim_head=d)
            x = transformer_vit(x, w, sub, prefix=f"t2t.{i}.layers.")
    return dense(x, w, "patch")                                                             # :73


def forward_t2t_vit(img, w, cfg):
    """t2t.py:96-116."""
    x = t2t_tokens(img, w, cfg)                                                             # :97
    b, n, d = x.shape
    x = np.concatenate([np.broadcast_to(w["cls_token"], (b, 1, d)), x], axis=1)             # :100-101
    x = x + w["pos_embedding"][:, :n + 1]                                                   # :102
    x = transformer_vit(x, w, cfg)                                                          # :105
    x = x.mean(axis=1) if cfg["pool"] == "mean" else x[:, 0]                                # :107-110
    return dense(layer_norm(x, w, "head_norm"), w, "head")                                  # :112


def forward(img, weights, cfg, dtype=np.float64):
    """Logits [b, num_classes] for an NHWC image batch, computed in `dtype`."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    img = np.asarray(img, dtype=dtype)
    kind = cfg["kind"]
    if kind in ("vit", "deepvit", "parallel_vit"):
        return forward_vit(img, w, cfg)
    if kind == "cait":
        return forward_cait(img, w, cfg)
    if kind == "crossvit":
        return forward_crossvit(img, w, cfg)
    if kind == "patch_merger_vit":
        return forward_patch_merger_vit(img, w, cfg)
    if kind == "t2t_vit":
        return forward_t2t_vit(img, w, cfg)
    raise ValueError(kind)


def forward_distill(img, distill_token, weights, cfg, dtype=np.float64):
    """DistillMixin.call with a distillation token (distill.py:16-45) on a ViT: returns (logits, distill_tokens)."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    img = np.asarray(img, dtype=dtype)
    x = patch_embed(img, w, "patch", cfg["patch_h"], cfg["patch_w"])                       # :18
    b, n, d = x.shape
    x = np.concatenate([np.broadcast_to(w["cls_token"], (b, 1, d)), x], axis=1)             # :21-22
    x = x + w["pos_embedding"][:, :n + 1]                                                   # :23
    tok = np.broadcast_to(np.asarray(distill_token, dtype=dtype).reshape(1, 1, d), (b, 1, d))
    x = np.concatenate([x, tok], axis=1)                                                    # :26-27
    x = transformer_vit(x, w, cfg)                                                          # :29 (_attend, dropout 0)
    x, dist = x[:, :-1], x[:, -1]                                                           # :32
    x = x.mean(axis=1) if cfg["pool"] == "mean" else x[:, 0]                                # :34-37
    return dense(layer_norm(x, w, "head_norm"), w, "head"), dist                            # :39-42


def embed_tokens(img, weights, cfg, dtype=np.float64):
    """Everything `call` does before `self.transformer`: patch embedding, cls token (where the model has one), positions
    (vit.py:160-166, cait.py:181-184, vit_with_patch_merger.py:175-179, t2t.py:97-103, efficient.py:40-45)."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    img = np.asarray(img, dtype=dtype)
    x = t2t_tokens(img, w, cfg) if cfg["kind"] == "t2t_vit" else patch_embed(img, w, "patch", cfg["patch_h"], cfg["patch_w"])
    b, n, d = x.shape
    if "cls_token" in w and cfg["kind"] != "cait":
        x = np.concatenate([np.broadcast_to(w["cls_token"], (b, 1, d)), x], axis=1)
    return x + w["pos_embedding"][:, :x.shape[1]]


def head_logits(x, weights, cfg, dtype=np.float64):
    """Everything `call` does after `self.transformer`: pooling + mlp_head (vit.py:170-175)."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    x = np.asarray(x, dtype=dtype)
    x = x.mean(axis=1) if cfg.get("pool", "cls") == "mean" else x[:, 0]
    return dense(layer_norm(x, w, "head_norm"), w, "head")


def transformer_tokens(x, weights, cfg, dtype=np.float64):
    """`model.transformer(tokens)` for ViT/DeepViT with arbitrary n (mae.py:69, simmim.py:116)."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    return transformer_vit(np.asarray(x, dtype=dtype), w, cfg)
