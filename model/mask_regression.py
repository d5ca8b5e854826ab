"""`from model.mask_regression import *` (reference model/resnet_generator_app_v2.py:5) -> the B200-native module."""
from layout2img_b200.model.mask_regression import MaskRegressNetv2  # noqa: F401

__all__ = ["MaskRegressNetv2"]
