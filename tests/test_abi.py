"""CPU tests of the boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
include/vitb200.h declares, refuses to compute without a GPU, and the Python host classes validate kwargs with
the reference's assertion messages (vit.py:136,139; deepvit.py:117; cait.py:160; cross_vit.py:207)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vitb200.h")).read()
    return sorted(set(re.findall(r"VB_API[^;(]*?\b(vb_\w+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    from vit_tensorflow_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vitb200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes SIGNATURES must cover exactly the declared ABI"
    assert lib.vb_abi_version() == 3


def test_config_struct_layout_matches_header():
    from vit_tensorflow_b200 import _lib
    src = open(os.path.join(ROOT, "include", "vitb200.h")).read()
    body = src[src.index("typedef struct vb_config {"):src.index("} vb_config;")]
    fields = []
    for line in body.splitlines():
        line = line.split("/*")[0].strip()
        if line.startswith("int32_t"):
            fields += [f.strip() for f in line[len("int32_t"):].rstrip(";").split(",")]
    assert [f for f, _ in _lib.VbConfig._fields_] == fields


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from vit_tensorflow_b200 import _lib, ViT
    with pytest.raises(_lib.VbError, match="no CUDA device"):
        _lib.op_linear(np.zeros((4, 8), np.float32), np.zeros((8, 64), np.float32))
    with pytest.raises(_lib.VbError, match="no CUDA device"):
        ViT(image_size=32, patch_size=16, num_classes=4, dim=32, depth=1, heads=2, mlp_dim=32, dim_head=16)


def test_reference_assertions(lib):
    from vit_tensorflow_b200 import ViT, DeepViT, CaiT, CrossViT
    kw = dict(num_classes=4, dim=32, depth=1, heads=2, mlp_dim=32)
    with pytest.raises(AssertionError, match="Image dimensions must be divisible by the patch size."):
        ViT(image_size=30, patch_size=16, **kw)
    with pytest.raises(AssertionError, match="pool type must be either cls"):
        ViT(image_size=32, patch_size=16, pool="max", **kw)
    with pytest.raises(AssertionError, match="Image dimensions must be divisible by the patch size."):
        DeepViT(image_size=30, patch_size=16, **kw)
    with pytest.raises(AssertionError, match="Image dimensions must be divisible by the patch size."):
        CaiT(image_size=30, patch_size=16, cls_depth=1, **kw)
    with pytest.raises(AssertionError, match="Image dimensions must be divisible by the patch size."):
        CrossViT(image_size=30, num_classes=4, sm_dim=32, lg_dim=32)
    with pytest.raises(NotImplementedError, match="layer_dropout"):
        CaiT(image_size=32, patch_size=16, cls_depth=1, layer_dropout=0.05, **kw)


def test_drop_in_import_names():
    # README.md:47,148,177,325 of the reference
    from vit_tensorflow import ViT
    from vit_tensorflow.deepvit import DeepViT
    from vit_tensorflow.cait import CaiT
    from vit_tensorflow.cross_vit import CrossViT
    import vit_tensorflow_b200 as vb
    assert (ViT, DeepViT, CaiT, CrossViT) == (vb.ViT, vb.DeepViT, vb.CaiT, vb.CrossViT)


def test_built_for_sm100a_with_tcgen05(lib):
    """The shipped library contains sm_100a SASS with tcgen05 (UTCHMMA), TMA (UTMALDG/UTMASTG) and TMEM loads."""
    import shutil
    import subprocess
    from vit_tensorflow_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, mnemonic
