"""`from model.norm_module import *` (reference model/resnet_generator_app_v2.py:4) -> the B200-native module."""
from layout2img_b200.model.norm_module import SpatialAdaptiveSynBatchNorm2d  # noqa: F401

__all__ = ["SpatialAdaptiveSynBatchNorm2d"]
