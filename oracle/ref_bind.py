"""Instantiate the REAL reference classes and load the oracle's named weights into them by attribute path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works against real TensorFlow (tools/ref_tf_dump.py, for anyone who
has it) and against the numpy stand-in of oracle/tf_shim.py (tests/test_reference_shim.py,
tests/golden/make_ref_golden.py) -- the attribute paths are the reference's own:

    model.patch_embedding.layers[1]                      'patch'            vit.py:131-134
    model.transformer.layers[L] = [PreNorm(Attention), PreNorm(MLP)]        vit.py:95-99
        .norm / .fn.to_qkv / .fn.to_out.layers[0] / .fn.net.layers[0|3]    vit.py:17,58,62,38-44
    model.mlp_head.layers[-2:]                           'head_norm','head' vit.py:144-147

The reference's modules are flat files importing each other by bare name (`from vit import Transformer`, t2t.py:6), so
the reference DIRECTORY goes on sys.path; nothing is copied out of it.
"""
from __future__ import annotations

import importlib

import numpy as np

# reference module and class per oracle kind
CLASSES = {"vit": ("vit", "ViT"), "deepvit": ("deepvit", "DeepViT"), "cait": ("cait", "CaiT"), "crossvit": ("cross_vit", "CrossViT"),
           "parallel_vit": ("parallel_vit", "ViT"), "patch_merger_vit": ("vit_with_patch_merger", "ViT"), "t2t_vit": ("t2t", "T2TViT")}


def to_numpy(t):
    return np.asarray(t.numpy() if hasattr(t, "numpy") else t)


def _set_dense(layer, w, name):
    layer.set_weights([w[name + ".kernel"]] + ([w[name + ".bias"]] if name + ".bias" in w else []))


def _set_ln(layer, w, name):
    layer.set_weights([w[name + ".gamma"], w[name + ".beta"]])


def _set_layer(attn_prenorm, ff_prenorm, w, pre, kind):
    """One [PreNorm(Attention), PreNorm(MLP)] pair.  kind: which Attention class this is (vit.py:48, deepvit.py:48,
    cait.py:85, cross_vit.py:52)."""
    _set_ln(attn_prenorm.norm, w, pre + "attn_norm")
    a = attn_prenorm.fn
    if kind in ("vit", "deepvit"):
        _set_dense(a.to_qkv, w, pre + "to_qkv")
    else:
        _set_dense(a.to_q, w, pre + "to_q")
        _set_dense(a.to_kv, w, pre + "to_kv")
    if kind == "deepvit":
        a.reattn_weights.assign(w[pre + "reattn_weights"])
        _set_ln(a.reattn_norm.layers[1], w, pre + "reattn_norm")
    if kind == "cait":
        a.mix_heads_pre_attn.assign(w[pre + "mix_pre"])
        a.mix_heads_post_attn.assign(w[pre + "mix_post"])
    if pre + "to_out.kernel" in w:
        _set_dense(a.to_out.layers[0], w, pre + "to_out")
    _set_ln(ff_prenorm.norm, w, pre + "ff_norm")
    _set_dense(ff_prenorm.fn.net.layers[0], w, pre + "fc1")
    _set_dense(ff_prenorm.fn.net.layers[3], w, pre + "fc2")


def _set_transformer(tr, w, prefix, kind):
    for L, (attn, ff) in enumerate(tr.layers):
        _set_layer(attn, ff, w, f"{prefix}{L}.", kind)


def _set_crossvit(model, w, cfg):
    for br in ("sm", "lg"):
        emb = getattr(model, f"{br}_image_embedder")                                   # cross_vit.py:254-255, 206-229
        _set_dense(emb.patch_embedding.layers[1], w, f"{br}_embed.patch")
        emb.pos_embedding.assign(w[f"{br}_embed.pos_embedding"])
        emb.cls_token.assign(w[f"{br}_embed.cls_token"])
        head = getattr(model, f"{br}_mlp_head").layers                                 # :280-288
        _set_ln(head[0], w, f"{br}_head_norm")
        _set_dense(head[1], w, f"{br}_head")
    for D, (sm_enc, lg_enc, cross) in enumerate(model.multi_scale_encoder.layers):     # :187-193
        for br, enc in (("sm", sm_enc), ("lg", lg_enc)):
            _set_transformer(enc, w, f"blocks.{D}.{br}_enc.layers.", "crossvit")
            _set_ln(enc.norm, w, f"blocks.{D}.{br}_enc.final_norm")                    # :100
        for R, pair in enumerate(cross.layers):                                        # :147-150
            for name, pio in zip(("sm_attend_lg", "lg_attend_sm"), pair):
                pre = f"blocks.{D}.cross.{R}.{name}."
                if pio.need_projection:                                                # :124-127
                    _set_dense(pio.project_in, w, pre + "project_in")
                    _set_dense(pio.project_out, w, pre + "project_out")
                _set_ln(pio.fn.norm, w, pre + "norm")
                a = pio.fn.fn
                _set_dense(a.to_q, w, pre + "to_q")
                _set_dense(a.to_kv, w, pre + "to_kv")
                _set_dense(a.to_out.layers[0], w, pre + "to_out")


def ctor_kwargs(case: dict) -> tuple:
    """tests/cases.py entry -> (oracle kind, the reference constructor's kwargs)."""
    kw = dict(case)
    kind = kw.pop("kind")
    if kind == "crossvit":
        kw.setdefault("dropout", 0.0)          # the reference defaults to 0.1 (cross_vit.py:251-252); inference ignores it
        kw.setdefault("emb_dropout", 0.0)
    return kind, kw


def build_model(kind: str, kw: dict, weights: dict, img, cls=None):
    """Construct the reference class for `kind` with kwargs `kw` (modules resolved through sys.path / sys.modules as they
    are NOW: call inside `tf_shim.installed(ref_dir)` or with real TensorFlow and the reference directory on sys.path),
    run it once so that Keras builds its variables, and overwrite every variable with `weights`."""
    if cls is None:
        mod, name = CLASSES[kind]
        cls = getattr(importlib.import_module(mod), name)
    model = cls(**kw)
    model(img, training=False)                       # builds the lazily-created Dense / LayerNormalization variables
    load_weights(model, kind, weights)
    return model


def load_weights(model, kind: str, w: dict) -> None:
    if kind == "crossvit":
        _set_crossvit(model, w, None)
        return
    model.pos_embedding.assign(w["pos_embedding"])
    if "cls_token" in w:
        model.cls_token.assign(w["cls_token"])
    if kind == "t2t_vit":
        # t2t.py:58-74: Sequential([RearrangeUnfoldTransformer, ..., Dense]); every soft split but the last owns
        # `transformer_layer` = vit.Transformer(depth=1)
        stages = model.patch_embedding.layers
        for i, st in enumerate(stages[:-2]):
            _set_transformer(st.transformer_layer, w, f"t2t.{i}.layers.", "vit")
        _set_dense(stages[-1], w, "patch")
    else:
        _set_dense(model.patch_embedding.layers[1], w, "patch")
    if kind == "parallel_vit":
        # parallel_vit.py:109-112: layers[L] = [Parallel(attention fns), Parallel(feed-forward fns)]
        for L, (attns, ffs) in enumerate(model.transformer.layers):
            for i, (attn, ff) in enumerate(zip(attns.fns, ffs.fns)):
                _set_layer(attn, ff, w, f"layers.{L}.branch{i}.", "vit")
    elif kind == "patch_merger_vit":
        _set_transformer(model.transformer, w, "layers.", "vit")
        pm = model.transformer.patch_merger                         # vit_with_patch_merger.py:42-47
        pm.queries.assign(w["patch_merger.queries"])
        _set_ln(pm.norm, w, "patch_merger.norm")
    elif kind == "cait":
        for stack in ("patch_transformer", "cls_transformer"):      # cait.py:172-173
            for L, (attn, ff) in enumerate(getattr(model, stack).layers):
                pre = f"{stack}.layers.{L}."
                attn.scale.assign(w[pre + "attn_scale"])            # LayerScale.scale cait.py:43-44
                ff.scale.assign(w[pre + "ff_scale"])
                _set_layer(attn.fn, ff.fn, w, pre, "cait")
    else:   # vit, deepvit, t2t_vit: vit.Transformer / deepvit.Transformer layers
        _set_transformer(model.transformer, w, "layers.", "vit" if kind == "t2t_vit" else kind)
    head = model.mlp_head.layers                                    # patch-merger ViT: [Reduce, LayerNormalization, Dense]
    _set_ln(head[-2], w, "head_norm")
    _set_dense(head[-1], w, "head")
