"""Independent third-party anchor for the oracle: Hugging Face `transformers.ViTForImageClassification` (PyTorch, installed
in this image) is the same architecture as the reference's `ViT(pool='cls')` (vit.py:106-177: patch projection, cls token,
learned positions on every row, pre-norm blocks with exact-erf GELU, final LayerNorm + Dense on the cls row) when
dim_head = dim / heads, qkv_bias = False and layer_norm_eps = 1e-3.  Mapping the oracle's Keras-layout weights into it and
comparing logits checks the oracle's restatement against code neither written here nor derived from the reference.

This does NOT pin the TensorFlow boundary (Keras' own LayerNormalization / Dense semantics stay as restated in SURVEY.md App. A
and oracle/tf_shim.py); the reference's own model code is pinned separately, by running it (tests/test_reference_shim.py)."""
import numpy as np
import pytest

import oracle

transformers = pytest.importorskip("transformers")
torch = pytest.importorskip("torch")


def _hf_model(cfg, w):
    from transformers import ViTConfig, ViTForImageClassification
    dim, heads = cfg["dim"], cfg["heads"]
    assert cfg["dim_head"] * heads == dim
    hc = ViTConfig(hidden_size=dim, num_hidden_layers=cfg["depth"], num_attention_heads=heads, intermediate_size=cfg["mlp_dim"],
                   hidden_act="gelu", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=1e-3,
                   image_size=(cfg["image_h"], cfg["image_w"]), patch_size=(cfg["patch_h"], cfg["patch_w"]), num_channels=3,
                   qkv_bias=False, num_labels=cfg["num_classes"])
    m = ViTForImageClassification(hc).eval()
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32)   # noqa: E731
    sd = {}
    ph, pw = cfg["patch_h"], cfg["patch_w"]
    sd["vit.embeddings.cls_token"] = t(w["cls_token"])
    sd["vit.embeddings.position_embeddings"] = t(w["pos_embedding"])
    # Dense kernel rows are ordered (p1 p2 c) (vit.py:142) -> Conv2d weight [dim, c, p1, p2]
    sd["vit.embeddings.patch_embeddings.projection.weight"] = t(w["patch.kernel"].reshape(ph, pw, 3, dim).transpose(3, 2, 0, 1))
    sd["vit.embeddings.patch_embeddings.projection.bias"] = t(w["patch.bias"])
    inner = dim
    for L in range(cfg["depth"]):
        p, q = f"layers.{L}.", f"vit.encoder.layer.{L}."
        qkv = w[p + "to_qkv.kernel"]                                           # [dim, 3 * inner], columns [q | k | v] (vit.py:72-73)
        for i, name in enumerate(("query", "key", "value")):
            sd[q + f"attention.attention.{name}.weight"] = t(qkv[:, i * inner:(i + 1) * inner].T)
        sd[q + "attention.output.dense.weight"] = t(w[p + "to_out.kernel"].T)
        sd[q + "attention.output.dense.bias"] = t(w[p + "to_out.bias"])
        sd[q + "layernorm_before.weight"], sd[q + "layernorm_before.bias"] = t(w[p + "attn_norm.gamma"]), t(w[p + "attn_norm.beta"])
        sd[q + "layernorm_after.weight"], sd[q + "layernorm_after.bias"] = t(w[p + "ff_norm.gamma"]), t(w[p + "ff_norm.beta"])
        sd[q + "intermediate.dense.weight"], sd[q + "intermediate.dense.bias"] = t(w[p + "fc1.kernel"].T), t(w[p + "fc1.bias"])
        sd[q + "output.dense.weight"], sd[q + "output.dense.bias"] = t(w[p + "fc2.kernel"].T), t(w[p + "fc2.bias"])
    sd["vit.layernorm.weight"], sd["vit.layernorm.bias"] = t(w["head_norm.gamma"]), t(w["head_norm.beta"])
    sd["classifier.weight"], sd["classifier.bias"] = t(w["head.kernel"].T), t(w["head.bias"])
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("bias" in k and ("query" in k or "key" in k or "value" in k) for k in missing), missing   # qkv_bias=False leaves none
    return m


@pytest.mark.parametrize("kw", [
    dict(image_size=64, patch_size=16, num_classes=10, dim=64, depth=2, heads=4, mlp_dim=128, dim_head=16),
    dict(image_size=(48, 64), patch_size=(8, 16), num_classes=7, dim=96, depth=3, heads=3, mlp_dim=160, dim_head=32),
    dict(image_size=224, patch_size=16, num_classes=1000, dim=192, depth=1, heads=3, mlp_dim=768),          # BASELINE configs[0]
])
def test_oracle_vit_matches_huggingface_vit(kw):
    cfg = oracle.make_config("vit", **kw)
    w = oracle.stress_weights(cfg, 5)
    img = oracle.make_image(cfg, 2, 6)
    ref = oracle.forward_numpy(img, w, cfg)
    m = _hf_model(cfg, w)
    with torch.no_grad():
        got = m(pixel_values=torch.tensor(img).permute(0, 3, 1, 2).contiguous()).logits.numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-5)
