"""Generate the golden fixtures under tests/golden/ from the numpy-float64 spec oracle.

    PYTHONPATH=. python tests/golden/make_golden.py

These vectors pin the ORACLE's own output (so that later edits cannot drift).  The vectors produced by the reference's
own code are the `*__refshim.npz` files next to them (tests/golden/make_ref_golden.py: unmodified reference modules over
the numpy TensorFlow stand-in of oracle/tf_shim.py); `tools/ref_tf_dump.py` regenerates them with real TensorFlow
wherever it is available; all three must agree to fp32 round-off.
Fixtures hold config + seeds + float64 logits only (weights/images are regenerated from the seeds)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from cases import SMALL  # noqa: E402

for name, d in SMALL.items():
    for wname in ("init_weights", "stress_weights"):
        kw = dict(d)
        cfg = oracle.make_config(kw.pop("kind"), **kw)
        path = os.path.join(HERE, f"{name}__{wname}.npz")
        if os.path.exists(path) and "--force" not in sys.argv:   # committed fixtures are never rewritten silently
            continue
        meta = dict(config=d, weights=wname, weight_seed=11, image_seed=12, batch=2)
        w = getattr(oracle, wname)(cfg, 11)
        img = oracle.make_image(cfg, 2, 12)
        logits = oracle.forward_numpy(img, w, cfg)
        np.savez(path, meta=json.dumps(meta), logits_f64=logits)
        print(name, wname, logits.shape, float(np.abs(logits).mean()))
