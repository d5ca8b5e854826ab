"""`from model.sync_batchnorm import SynchronizedBatchNorm2d, DataParallelWithCallback` (reference
train_context_app_v2.py:22, model/resnet_generator_app_v2.py:6).  One process per GPU replaces the reference's
single-process DataParallel replication, so DataParallelWithCallback is the identity wrapper here."""
from layout2img_b200.model.layers import SynchronizedBatchNorm2d  # noqa: F401


def DataParallelWithCallback(module, *args, **kwargs):
    """The reference replicates the generator across GPUs inside one process (sync_batchnorm/replicate.py); this
    framework runs one process per GPU (layout2img_b200.train.GradBuckets), so the module is returned as is."""
    return module


__all__ = ["SynchronizedBatchNorm2d", "DataParallelWithCallback"]
