from vit_tensorflow_b200 import DeepViT  # noqa: F401
