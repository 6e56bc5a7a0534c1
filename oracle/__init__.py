"""CPU oracle for the ViT-family forward path of taki0112/vit-tensorflow.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
arm, never as the thing shipped.  The product path (``vit_tensorflow_b200``) never
imports this package and fails loudly when its CUDA library is missing.

PARITY UNPINNED: the reference (TensorFlow/Keras + einops, un-vendored, version
">= 2.3.0") cannot be imported in this image (no TensorFlow), and the reference
ships no tests / golden vectors.  The oracle is therefore a
restatement, made trustworthy by (1) two independent implementations -- a
numpy-float64 "spec" (``spec_numpy``) and a torch-CPU-float32 restatement
(``ref_torch``) -- that must agree, (2) einops itself (installed here) used as the
ground truth for the patch ``Rearrange``, (3) ``tools/ref_tf_dump.py``, a hook
that dumps golden vectors from the real reference wherever TensorFlow exists, and
(4) ``tests/test_oracle_vs_hf_vit.py``: the ViT restatement reproduces Hugging Face's
independent PyTorch ``ViTForImageClassification`` on mapped weights (architecture and
layouts, not Keras op semantics).
"""
from .weights import (make_config, weight_specs, init_weights, stress_weights,  # noqa: F401
                      make_image, flops_per_image, t2t_token_grid)
from .spec_numpy import forward as forward_numpy  # noqa: F401
