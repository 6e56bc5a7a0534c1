"""Golden fixtures produced by the REFERENCE'S OWN CODE (imported from /root/reference, unmodified) running over the numpy
stand-in for TensorFlow in oracle/tf_shim.py.

    python tests/golden/make_ref_golden.py [--force] [--reference /root/reference]

For every case of tests/cases.py (SMALL), every BASELINE.json configuration at its own width / depth / heads (FULL, batch 2) and the
reference README's CrossViT / T2TViT usage examples at their own size (README, batch 2)
this instantiates the reference class (vit.ViT, deepvit.DeepViT, cait.CaiT, cross_vit.CrossViT, parallel_vit.ViT,
vit_with_patch_merger.ViT, t2t.T2TViT), loads the oracle's seeded weights into its Keras variables by attribute path
(oracle/ref_bind.py), calls `model(img, training=False)` and stores the logits:

    logits_ref_f32   the reference computing in float32 (what TensorFlow's default dtype would do, up to summation order)
    logits_ref_f64   the reference computing in float64 (SMALL cases only; compared with the float64 spec at 1e-12)

What these fixtures pin and what they do not is spelled out in oracle/tf_shim.py: the model code is the reference's, the
~35 TensorFlow / Keras primitives under it are numpy restatements.  /root/reference does not exist on the GPU box, which is
why the outputs are committed; weights and images are regenerated there from the seeds in `meta`.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from oracle import ref_bind, tf_shim  # noqa: E402
from cases import SMALL, FULL, README  # noqa: E402

WEIGHT_SEED, IMAGE_SEED, BATCH = 11, 12, 2


def run_reference(case, wname, dtype, ref_dir):
    kind, kw = ref_bind.ctor_kwargs(case)
    cfg = oracle.make_config(kind, **{k: v for k, v in case.items() if k != "kind"})
    w = {k: v.astype(dtype) for k, v in getattr(oracle, wname)(cfg, WEIGHT_SEED).items()}
    img = oracle.make_image(cfg, BATCH, IMAGE_SEED).astype(dtype)
    tf_shim.set_dtype(dtype)
    try:
        with tf_shim.installed(ref_dir):
            model = ref_bind.build_model(kind, kw, w, img)
            return ref_bind.to_numpy(model(img, training=False))
    finally:
        tf_shim.set_dtype(np.float32)


def main():
    ref_root = sys.argv[sys.argv.index("--reference") + 1] if "--reference" in sys.argv else "/root/reference"
    ref_dir = os.path.join(ref_root, "vit_tensorflow")
    if not os.path.isdir(ref_dir):
        raise SystemExit(f"{ref_dir} not found: these fixtures can only be generated where the reference checkout exists")
    for group, cases in (("small", SMALL), ("full", FULL), ("readme", README)):
        for name, case in cases.items():
            for wname in ("init_weights", "stress_weights"):
                path = os.path.join(HERE, f"{name}__{wname}__refshim.npz")
                if os.path.exists(path) and "--force" not in sys.argv:      # committed fixtures are never rewritten silently
                    continue
                t0 = time.time()
                out = dict(logits_ref_f32=run_reference(case, wname, np.float32, ref_dir))
                if group == "small":
                    out["logits_ref_f64"] = run_reference(case, wname, np.float64, ref_dir)
                meta = dict(config=case, weights=wname, weight_seed=WEIGHT_SEED, image_seed=IMAGE_SEED, batch=BATCH,
                            generator="tests/golden/make_ref_golden.py", reference="vit_tensorflow (unmodified) over oracle/tf_shim.py",
                            numpy=np.__version__)
                np.savez(path, meta=json.dumps(meta), **out)
                print(f"{name} {wname}: {out['logits_ref_f32'].shape}, |logits| mean {np.abs(out['logits_ref_f32']).mean():.3f} "
                      f"({time.time() - t0:.1f} s)")


if __name__ == "__main__":
    main()
