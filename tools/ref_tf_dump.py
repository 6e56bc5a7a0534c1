"""Golden-vector hook for anyone WITH TensorFlow: run the REAL reference and dump {logits} next to the oracle's.

    python tools/ref_tf_dump.py /path/to/vit-tensorflow  [out_dir]          # real TensorFlow
    python tools/ref_tf_dump.py /path/to/vit-tensorflow  [out_dir] --shim   # numpy stand-in (oracle/tf_shim.py), runs here

TensorFlow is not installed in this image (and there is no network), so without --shim this script cannot run here; it is
the documented way to pin the oracle at the TF boundary.  It instantiates the reference models
with the same kwargs as tests/cases.py, copies the oracle's seeded weights INTO the Keras variables by attribute
path (oracle/ref_bind.py), runs `model(img, training=False)` and writes `<case>__tf.npz`; compare with
tests/golden/<case>__*.npz (expected agreement: fp32 round-off, ~1e-5).  With --shim the very same code path runs over the
numpy stand-in (that is how tests/golden/*__refshim.npz were made, by tests/golden/make_ref_golden.py): a maintainer with
TensorFlow only has to drop the flag to replace the one remaining assumption -- the primitives' semantics -- by the real thing.
"""
import contextlib
import json
import os
import sys

import numpy as np


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    shim = "--shim" in sys.argv
    ref_root = args[0]
    out_dir = args[1] if len(args) > 1 else "tests/golden"
    ref_dir = os.path.join(ref_root, "vit_tensorflow")                # flat sibling imports, no __init__.py
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, repo)
    sys.path.insert(0, os.path.join(repo, "tests"))
    import oracle
    from oracle import ref_bind
    from cases import SMALL
    if shim:
        from oracle import tf_shim
        env = tf_shim.installed(ref_dir)
    else:
        sys.path.insert(0, ref_dir)
        env = contextlib.nullcontext()
    suffix = "shim" if shim else "tf"
    with env:
        for name, d in SMALL.items():
            kind, kw = ref_bind.ctor_kwargs(d)
            cfg = oracle.make_config(kind, **{k: v for k, v in d.items() if k != "kind"})
            w = oracle.stress_weights(cfg, 11)
            img = oracle.make_image(cfg, 2, 12)
            model = ref_bind.build_model(kind, kw, w, img)
            logits = ref_bind.to_numpy(model(img, training=False))
            np.savez(os.path.join(out_dir, f"{name}__{suffix}.npz"), **{f"logits_{suffix}": logits},
                     meta=json.dumps(dict(config=d, weights="stress_weights", weight_seed=11, image_seed=12, batch=2)))
            ref = oracle.forward_numpy(img, w, cfg)
            print(name, f"max |{suffix} - oracle| =", float(np.abs(logits - ref).max()))


if __name__ == "__main__":
    main()
