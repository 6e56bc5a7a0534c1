from vit_tensorflow_b200 import ViT  # noqa: F401
