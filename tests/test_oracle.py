"""CPU tests of the oracle: the two independent restatements agree, shapes match the reference's usage
docstrings, FLOP accounting matches SURVEY.md App. C, and the committed golden fixtures reproduce."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import ref_torch, spec_numpy
from cases import SMALL, cfg_of

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(SMALL))
@pytest.mark.parametrize("gen", ["init", "stress"])
def test_spec_vs_torch(name, gen):
    cfg = cfg_of(name)
    w = (oracle.init_weights if gen == "init" else oracle.stress_weights)(cfg, 3)
    img = oracle.make_image(cfg, 2, 5)
    a = oracle.forward_numpy(img, w, cfg)
    b = ref_torch.forward(img, w, cfg)
    assert a.shape == (2, cfg["num_classes"])
    np.testing.assert_allclose(b, a, rtol=1e-4, atol=2e-5)


def test_patch_order_is_einops_order():
    # vit.py:142: patch vector order (p1 p2 c) with c fastest == reshape/transpose restatement
    rng = np.random.default_rng(0)
    img = rng.standard_normal((2, 12, 8, 3)).astype(np.float32)
    from einops import rearrange
    a = rearrange(img, 'b (h p1) (w p2) c -> b (h w) (p1 p2 c)', p1=4, p2=2)
    b = img.reshape(2, 3, 4, 4, 2, 3).transpose(0, 1, 3, 2, 4, 5).reshape(2, 12, 24)
    np.testing.assert_array_equal(a, b)


def test_usage_docstring_shapes():
    # vit.py:179-196 (scaled down): [1, H, W, 3] -> (1, num_classes)
    cfg = oracle.make_config("vit", image_size=64, patch_size=32, num_classes=1000, dim=64, depth=1, heads=2, mlp_dim=64, dim_head=16)
    y = ref_torch.forward(oracle.make_image(cfg, 1), oracle.init_weights(cfg), cfg)
    assert y.shape == (1, 1000)


def test_fewer_patches_than_pos_embedding():
    # vit.py:165 pos_embedding[:, :n+1] allows smaller images (README.md:909-934)
    cfg = cfg_of("vit_small")
    w = oracle.init_weights(cfg, 0)
    img = oracle.make_image(cfg, 2, 0, h=32, w=48)
    a = oracle.forward_numpy(img, w, cfg)
    b = ref_torch.forward(img, w, cfg)
    np.testing.assert_allclose(b, a, rtol=1e-4, atol=2e-5)


def test_flops_match_survey_appendix_c():
    c1 = cfg_of("vit_c1")
    c2 = oracle.make_config("vit", image_size=224, patch_size=16, num_classes=1000, dim=768, depth=12, heads=12, mlp_dim=3072)
    c3 = oracle.make_config("deepvit", image_size=224, patch_size=16, num_classes=1000, dim=1024, depth=24, heads=16, mlp_dim=4096)
    c4 = oracle.make_config("cait", image_size=224, patch_size=16, num_classes=1000, dim=384, depth=36, cls_depth=2, heads=8, mlp_dim=1536, dim_head=48)
    c5 = oracle.make_config("vit", image_size=384, patch_size=16, num_classes=1000, dim=1024, depth=24, heads=16, mlp_dim=4096)
    for cfg, g in ((c1, 0.2623), (c2, 35.128), (c3, 123.59), (c4, 27.80), (c5, 382.13)):
        assert abs(oracle.flops_per_image(cfg) / 1e9 - g) < 5e-3 * g


def test_transformer_tokens_any_n():
    cfg = cfg_of("vit_small")
    w = oracle.init_weights(cfg, 0)
    x = np.random.default_rng(0).standard_normal((2, 7, cfg["dim"])).astype(np.float32)
    a = spec_numpy.transformer_tokens(x, w, cfg)
    b = ref_torch.TorchReference(w, cfg).transformer(x)
    np.testing.assert_allclose(b, a, rtol=1e-4, atol=2e-5)


def test_distill_forward_reduces_to_vit_and_tokens():
    """distill.py:16-45 restated: logits of the distill call equal mlp_head over the first n+1 rows of the transformer
    run on [cls; patches; token]; the two oracle restatements of the pieces agree."""
    cfg = cfg_of("vit_small")
    w = oracle.stress_weights(cfg, 1)
    img = oracle.make_image(cfg, 2, 5)
    tok = np.random.default_rng(1).standard_normal((1, 1, cfg["dim"]))
    logits, dist = spec_numpy.forward_distill(img, tok, w, cfg)
    wd = {k: np.asarray(v, np.float64) for k, v in w.items()}
    x = spec_numpy.patch_embed(np.asarray(img, np.float64), wd, "patch", cfg["patch_h"], cfg["patch_w"])
    b, n, d = x.shape
    x = np.concatenate([np.broadcast_to(wd["cls_token"], (b, 1, d)), x], 1) + wd["pos_embedding"][:, :n + 1]
    x = np.concatenate([x, np.broadcast_to(tok, (b, 1, d))], 1)
    y = ref_torch.TorchReference(w, cfg).transformer(x.astype(np.float32))
    np.testing.assert_allclose(dist, y[:, -1], rtol=1e-4, atol=5e-5)
    assert logits.shape == (2, cfg["num_classes"]) and np.isfinite(logits).all()
    # a zero-influence check: the token changes the logits (it attends with every row)
    l2, _ = spec_numpy.forward_distill(img, tok * 2.0, w, cfg)
    assert np.abs(l2 - logits).max() > 1e-6


def test_config_asserts_match_reference_messages():
    with pytest.raises(AssertionError, match="Image dimensions must be divisible by the patch size."):
        oracle.make_config("vit", image_size=30, patch_size=16, num_classes=2, dim=8, depth=1, heads=1, mlp_dim=8)
    with pytest.raises(AssertionError, match="pool type must be either cls"):
        oracle.make_config("vit", image_size=32, patch_size=16, num_classes=2, dim=8, depth=1, heads=1, mlp_dim=8, pool="max")


@pytest.mark.parametrize("name", sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.endswith("__refshim.npz")) if os.path.isdir(GOLDEN) else [])
def test_golden_fixture(name):
    """Committed fixtures (tests/golden/make_golden.py): oracle output pinned so that later edits cannot drift."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    kw = dict(meta["config"])
    cfg = oracle.make_config(kw.pop("kind"), **{k: (tuple(v) if isinstance(v, list) else v) for k, v in kw.items()})
    w = getattr(oracle, meta["weights"])(cfg, meta["weight_seed"])
    img = oracle.make_image(cfg, meta["batch"], meta["image_seed"])
    np.testing.assert_allclose(oracle.forward_numpy(img, w, cfg), z["logits_f64"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(ref_torch.forward(img, w, cfg), z["logits_f64"], rtol=1e-4, atol=2e-5)
