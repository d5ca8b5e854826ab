"""Drop-in import path of the reference's `model` package (train_context_app_v2.py:18-19, test_context_app_v2.py:10):

    from model.resnet_generator_app_v2 import *
    from model.rcnn_discriminator_app import *

resolve to the B200-native modules in layout2img_b200.model when this repository's root is on sys.path, so the
reference's scripts need no edited import lines.
"""
