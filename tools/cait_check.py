import sys, os, numpy as np
sys.path.insert(0, '/root/repo')
import oracle
from vit_tensorflow_b200 import from_config
for depth in (2, 6):
    cfgc = oracle.make_config("cait", image_size=224, patch_size=16, num_classes=100, dim=384, depth=depth, cls_depth=1, heads=8, mlp_dim=768, dim_head=48)
    for gen in ("stress_weights", "init_weights"):
        wc = getattr(oracle, gen)(cfgc, 5)
        imgc = oracle.make_image(cfgc, 2, 9)
        mc = from_config(cfgc, precision="bf16", device=0)
        mc.set_weights_dict(wc)
        got = mc(imgc, training=False)
        ref = oracle.forward_numpy(imgc, wc, cfgc)
        m32 = from_config(cfgc, precision="fp32", device=0); m32.set_weights_dict(wc)
        g32 = m32(imgc, training=False)
        print(os.environ.get("VB_NO_ATTN_MIX"), depth, gen, "bf16 max err %.4f  fp32 max err %.2e  |ref| max %.2f std %.2f" % (np.abs(got-ref).max(), np.abs(g32-ref).max(), np.abs(ref).max(), ref.std()), flush=True)
