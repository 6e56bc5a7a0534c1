from vit_tensorflow_b200 import CaiT  # noqa: F401
