"""Golden-vector hook for anyone WITH TensorFlow: run the REAL reference and dump {logits} next to the oracle's.

    python tools/ref_tf_dump.py /path/to/vit-tensorflow  [out_dir]

TensorFlow is not installed in this image (and there is no network), so this script cannot run here; it is the
documented way to pin the oracle at the TF boundary.  It instantiates the reference models
with the same kwargs as tests/cases.py, copies the oracle's seeded weights INTO the Keras variables by attribute
path, runs `model(img, training=False)` and writes `<case>__tf.npz`; compare with
tests/golden/<case>__*.npz (expected agreement: fp32 round-off, ~1e-5).
"""
import json
import os
import sys

import numpy as np


def _set_dense(layer, w, name):
    vals = [w[name + ".kernel"]] + ([w[name + ".bias"]] if name + ".bias" in w else [])
    layer.set_weights(vals)


def _set_ln(layer, w, name):
    layer.set_weights([w[name + ".gamma"], w[name + ".beta"]])


def _set_vit_layer(attn_prenorm, ff_prenorm, w, pre, kind):
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


def main():
    ref_root = sys.argv[1]
    out_dir = sys.argv[2] if len(sys.argv) > 2 else "tests/golden"
    sys.path.insert(0, os.path.join(ref_root, "vit_tensorflow"))   # flat sibling imports, no __init__.py
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, repo)
    sys.path.insert(0, os.path.join(repo, "tests"))
    import importlib
    import oracle
    from cases import SMALL
    # reference module and class per oracle kind (flat imports: `from vit import Transformer` inside t2t.py etc.)
    classes = {"vit": ("vit", "ViT"), "deepvit": ("deepvit", "DeepViT"), "cait": ("cait", "CaiT"),
               "parallel_vit": ("parallel_vit", "ViT"), "patch_merger_vit": ("vit_with_patch_merger", "ViT"), "t2t_vit": ("t2t", "T2TViT")}

    for name, d in SMALL.items():
        kw = dict(d)
        kind = kw.pop("kind")
        if kind not in classes:
            print(name, "skipped (CrossViT: same recipe; attribute paths in SURVEY.md App. B)")
            continue
        cfg = oracle.make_config(kind, **kw)
        w = oracle.stress_weights(cfg, 11)
        img = oracle.make_image(cfg, 2, 12)
        mod, cls = classes[kind]
        model = getattr(importlib.import_module(mod), cls)(**kw)
        model(img, training=False)  # build variables
        model.pos_embedding.assign(w["pos_embedding"])
        if "cls_token" in w:
            model.cls_token.assign(w["cls_token"])
        if kind == "t2t_vit":
            # t2t.py:58-74: Sequential([RearrangeUnfoldTransformer, ..., Dense]); every soft split but the last owns
            # `transformer_layer` = vit.Transformer(depth=1)
            stages = model.patch_embedding.layers
            for i, st in enumerate(stages[:-2]):
                attn, ff = st.transformer_layer.layers[0]
                _set_vit_layer(attn, ff, w, f"t2t.{i}.layers.0.", "vit")
            _set_dense(stages[-1], w, "patch")
        else:
            _set_dense(model.patch_embedding.layers[1], w, "patch")
        if kind == "parallel_vit":
            # parallel_vit.py:109-112: layers[L] = [Parallel(attention fns), Parallel(feed-forward fns)]
            for L, (attns, ffs) in enumerate(model.transformer.layers):
                for i, (attn, ff) in enumerate(zip(attns.fns, ffs.fns)):
                    _set_vit_layer(attn, ff, w, f"layers.{L}.branch{i}.", "vit")
        elif kind == "patch_merger_vit":
            for L, (attn, ff) in enumerate(model.transformer.layers):
                _set_vit_layer(attn, ff, w, f"layers.{L}.", "vit")
            pm = model.transformer.patch_merger                         # vit_with_patch_merger.py:42-47
            pm.queries.assign(w["patch_merger.queries"])
            _set_ln(pm.norm, w, "patch_merger.norm")
        elif kind == "cait":
            for stack in ("patch_transformer", "cls_transformer"):
                for L, (attn, ff) in enumerate(getattr(model, stack).layers):
                    pre = f"{stack}.layers.{L}."
                    attn.scale.assign(w[pre + "attn_scale"])
                    ff.scale.assign(w[pre + "ff_scale"])
                    _set_vit_layer(attn.fn, ff.fn, w, pre, "cait")
        else:   # vit, deepvit, t2t_vit: vit.Transformer / deepvit.Transformer layers
            for L, (attn, ff) in enumerate(model.transformer.layers):
                _set_vit_layer(attn, ff, w, f"layers.{L}.", "vit" if kind == "t2t_vit" else kind)
        head = model.mlp_head.layers                                    # patch-merger ViT: [Reduce, LayerNormalization, Dense]
        _set_ln(head[-2], w, "head_norm")
        _set_dense(head[-1], w, "head")
        logits = model(img, training=False).numpy()
        np.savez(os.path.join(out_dir, f"{name}__tf.npz"), logits_tf=logits,
                 meta=json.dumps(dict(config=d, weights="stress_weights", weight_seed=11, image_seed=12, batch=2)))
        ref = oracle.forward_numpy(img, w, cfg)
        print(name, "max |tf - oracle| =", float(np.abs(logits - ref).max()))


if __name__ == "__main__":
    main()
