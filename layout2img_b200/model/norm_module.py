"""Drop-in for the reference's model/norm_module.py:152-189."""
from __future__ import annotations

import torch.nn as nn

from .layers import SynchronizedBatchNorm2d


class SpatialAdaptiveSynBatchNorm2d(nn.Module):
    """ISLA norm.  Holds the reference's parameters (spectral-normed gamma/beta projections, BN running
    statistics); the arithmetic runs inside the fused generator block (functional.GBlockFn, csrc/isla.cu)."""

    def __init__(self, num_features, num_w=512, batchnorm_func=SynchronizedBatchNorm2d, eps=1e-5, momentum=0.1,
                 affine=False, track_running_stats=True):
        super().__init__()
        self.num_features = num_features
        self.weight_proj = nn.utils.spectral_norm(nn.Linear(num_w, num_features))
        self.bias_proj = nn.utils.spectral_norm(nn.Linear(num_w, num_features))
        self.batch_norm2d = batchnorm_func(num_features, eps=eps, momentum=momentum, affine=affine)

    def forward(self, x, vector, bbox):
        raise RuntimeError("the ISLA arithmetic is fused into the generator block (functional.g_block)")

    def __repr__(self):
        return self.__class__.__name__ + '(' + str(self.num_features) + ')'
