"""`from model.resnet_generator_app_v2 import *` (reference train_context_app_v2.py:18) -> the B200-native modules."""
from layout2img_b200.model.resnet_generator_app_v2 import *  # noqa: F401,F403
from layout2img_b200.model.resnet_generator_app_v2 import __all__  # noqa: F401
