"""Short ncu target: a few forwards of one model through the C-ABI with host buffers (no torch in the process).

    python tools/ncu_target.py vit_b16 [batch] [forwards]     ViT-B/16 224^2 (BASELINE configs[1]), default batch 256
    python tools/ncu_target.py t2t [batch] [forwards]         T2TViT 224^2 (soft-split kernels), default batch 8
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from vit_tensorflow_b200 import ViT, T2TViT  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "vit_b16"
if which == "vit_b16":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    m = ViT(image_size=224, patch_size=16, num_classes=1000, dim=768, depth=12, heads=12, mlp_dim=3072, seed=0)
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    m = T2TViT(image_size=224, num_classes=1000, dim=512, depth=2, heads=8, mlp_dim=512, seed=0)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
img = np.random.default_rng(0).standard_normal((B, 224, 224, 3), dtype=np.float32)
for _ in range(n):
    out = m(img, training=False)
print(which, B, out.shape, float(np.abs(out).mean()), m.last_launch_count)
