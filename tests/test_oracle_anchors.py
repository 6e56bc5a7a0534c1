"""Second, independent anchors for the oracle's building blocks: the numpy-float64 spec against PyTorch's own
implementations of the same published operators, on random inputs.  None of this is TensorFlow (absent from this image: parity stays
"unpinned" at the Keras boundary), but it is third-party code the spec did not come from:

  attention core (vit.py:77-82)            F.scaled_dot_product_attention
  fused-QKV multi-head attention            torch.nn.MultiheadAttention with mapped weights (vit.py:59,63,72-84)
  extract_patches SAME (t2t.py:43)          F.pad with TF's documented SAME split + F.unfold (window extraction, stride handling)
  LayerNormalization eps 1e-3 (vit.py:18)   F.layer_norm
  exact-erf GELU (vit.py:29-34)             F.gelu(approximate='none')
  Rearrange patches (vit.py:142)            F.unfold with kernel = stride = patch (window order) -- besides einops itself
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import spec_numpy as S


def test_attention_core_vs_torch_sdpa():
    rng = np.random.default_rng(0)
    b, h, n, d = 2, 3, 37, 16
    x = rng.standard_normal((b, n, h * d))
    w = {"l.to_qkv.kernel": rng.standard_normal((h * d, 3 * h * d)) / np.sqrt(h * d)}
    got = S.attention_vit(x, w, "l.", h, d)                          # no to_out in w: the attention core, heads merged
    qkv = torch.from_numpy(x @ w["l.to_qkv.kernel"])
    q, k, v = (t.reshape(b, n, h, d).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, h * d).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12)


def test_fused_qkv_attention_vs_torch_multihead_attention():
    rng = np.random.default_rng(1)
    b, h, n, d = 2, 4, 19, 8
    e = h * d
    x = rng.standard_normal((b, n, e))
    w = {"l.to_qkv.kernel": rng.standard_normal((e, 3 * e)) / np.sqrt(e), "l.to_out.kernel": rng.standard_normal((e, e)) / np.sqrt(e),
         "l.to_out.bias": rng.standard_normal(e)}
    got = S.attention_vit(x, w, "l.", h, d)
    mha = torch.nn.MultiheadAttention(e, h, bias=True, batch_first=True, dtype=torch.float64)
    with torch.no_grad():
        mha.in_proj_weight.copy_(torch.from_numpy(w["l.to_qkv.kernel"].T))      # rows [q | k | v] = tf.split(qkv, 3, -1) (vit.py:73)
        mha.in_proj_bias.zero_()
        mha.out_proj.weight.copy_(torch.from_numpy(w["l.to_out.kernel"].T))
        mha.out_proj.bias.copy_(torch.from_numpy(w["l.to_out.bias"]))
        xt = torch.from_numpy(x)
        ref, _ = mha(xt, xt, xt, need_weights=False)
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("H,W,k,s", [(224, 224, 7, 4), (56, 56, 3, 2), (28, 28, 3, 2), (9, 11, 3, 2), (10, 7, 5, 3), (8, 8, 2, 2)])
def test_extract_patches_same_vs_torch_unfold(H, W, k, s):
    rng = np.random.default_rng(H + k)
    b, C = 2, 3
    x = rng.standard_normal((b, H, W, C))
    got = S.extract_patches_same(x, k, s)                                        # [b, oh, ow, (k_row, k_col, c)]
    oh, ow = -(-H // s), -(-W // s)
    ph, pw = max((oh - 1) * s + k - H, 0), max((ow - 1) * s + k - W, 0)          # TF SAME: smaller half first
    xt = F.pad(torch.from_numpy(x).permute(0, 3, 1, 2), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    u = F.unfold(xt, kernel_size=k, stride=s)                                    # [b, C*k*k, L], rows ordered (c, k_row, k_col)
    assert u.shape[-1] == oh * ow
    ref = u.reshape(b, C, k, k, oh, ow).permute(0, 4, 5, 2, 3, 1).reshape(b, oh, ow, k * k * C).numpy()
    np.testing.assert_array_equal(got, ref)


def test_patch_rearrange_vs_torch_unfold():
    rng = np.random.default_rng(2)
    b, H, W, C, p1, p2 = 2, 24, 32, 3, 8, 16
    img = rng.standard_normal((b, H, W, C))
    w = {"patch.kernel": np.eye(p1 * p2 * C), "patch.bias": np.zeros(p1 * p2 * C)}
    got = S.patch_embed(img, w, "patch", p1, p2)                                 # identity Dense: the Rearrange alone
    u = F.unfold(torch.from_numpy(img).permute(0, 3, 1, 2), kernel_size=(p1, p2), stride=(p1, p2))      # rows (c, p1, p2)
    ref = u.reshape(b, C, p1, p2, -1).permute(0, 4, 2, 3, 1).reshape(b, -1, p1 * p2 * C).numpy()        # -> (p1 p2 c)
    np.testing.assert_array_equal(got, ref)


def test_layernorm_and_gelu_vs_torch():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 7, 48)) * 3 + 1
    w = {"n.gamma": rng.uniform(0.5, 1.5, 48), "n.beta": rng.standard_normal(48)}
    ref = F.layer_norm(torch.from_numpy(x), (48,), torch.from_numpy(w["n.gamma"]), torch.from_numpy(w["n.beta"]), eps=1e-3).numpy()
    np.testing.assert_allclose(S.layer_norm(x, w, "n"), ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(S.gelu(x), F.gelu(torch.from_numpy(x), approximate="none").numpy(), rtol=1e-12, atol=1e-14)
