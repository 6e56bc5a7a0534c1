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
    import oracle
    from cases import SMALL
    from vit import ViT
    from deepvit import DeepViT
    from cait import CaiT

    for name, d in SMALL.items():
        kw = dict(d)
        kind = kw.pop("kind")
        if kind == "crossvit":
            continue  # same recipe; attribute paths in SURVEY.md App. B
        cfg = oracle.make_config(kind, **kw)
        w = oracle.stress_weights(cfg, 11)
        img = oracle.make_image(cfg, 2, 12)
        model = {"vit": ViT, "deepvit": DeepViT, "cait": CaiT}[kind](**kw)
        model(img, training=False)  # build variables
        model.pos_embedding.assign(w["pos_embedding"])
        model.cls_token.assign(w["cls_token"])
        _set_dense(model.patch_embedding.layers[1], w, "patch")
        if kind == "cait":
            for stack in ("patch_transformer", "cls_transformer"):
                for L, (attn, ff) in enumerate(getattr(model, stack).layers):
                    pre = f"{stack}.layers.{L}."
                    attn.scale.assign(w[pre + "attn_scale"])
                    ff.scale.assign(w[pre + "ff_scale"])
                    _set_vit_layer(attn.fn, ff.fn, w, pre, "cait")
        else:
            for L, (attn, ff) in enumerate(model.transformer.layers):
                _set_vit_layer(attn, ff, w, f"layers.{L}.", kind)
        _set_ln(model.mlp_head.layers[0], w, "head_norm")
        _set_dense(model.mlp_head.layers[1], w, "head")
        logits = model(img, training=False).numpy()
        np.savez(os.path.join(out_dir, f"{name}__tf.npz"), logits_tf=logits,
                 meta=json.dumps(dict(config=d, weights="stress_weights", weight_seed=11, image_seed=12, batch=2)))
        ref = oracle.forward_numpy(img, w, cfg)
        print(name, "max |tf - oracle| =", float(np.abs(logits - ref).max()))


if __name__ == "__main__":
    main()
