from vit_tensorflow_b200 import CrossViT  # noqa: F401
