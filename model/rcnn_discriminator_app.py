"""`from model.rcnn_discriminator_app import *` (reference train_context_app_v2.py:19) -> the B200-native modules."""
from layout2img_b200.model.rcnn_discriminator_app import *  # noqa: F401,F403
from layout2img_b200.model.rcnn_discriminator_app import __all__  # noqa: F401
