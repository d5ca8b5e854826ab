"""Drop-in for the reference's model/norm_module.py:152-189."""
from __future__ import annotations

import torch.nn as nn

from .. import functional as L
from .layers import SynchronizedBatchNorm2d


class SpatialAdaptiveSynBatchNorm2d(nn.Module):
    """ISLA norm.  Holds the reference's parameters (spectral-normed gamma/beta projections, BN running
    statistics); `operands()` produces what the fused norm+ReLU+conv kernel consumes."""

    def __init__(self, num_features, num_w=512, batchnorm_func=SynchronizedBatchNorm2d, eps=1e-5, momentum=0.1,
                 affine=False, track_running_stats=True):
        super().__init__()
        self.num_features = num_features
        self.weight_proj = nn.utils.spectral_norm(nn.Linear(num_w, num_features))
        self.bias_proj = nn.utils.spectral_norm(nn.Linear(num_w, num_features))
        self.batch_norm2d = batchnorm_func(num_features, eps=eps, momentum=momentum, affine=affine)

    def operands(self, x, vector, bbox):
        """x NHWC (b,h,w,c); vector (b*o, num_w); bbox (b,o,hm,wm) -> (bn, mask_pm, gamma, beta)."""
        b, o = bbox.shape[:2]
        h, w = x.shape[1:3]
        mask_pm = L.mask_resize(bbox, h, w, True)          # bilinear if sizes differ (norm_module.py:173-176)
        gamma = self.weight_proj(vector).view(b, o, -1)
        beta = self.bias_proj(vector).view(b, o, -1)
        return self.batch_norm2d, mask_pm, gamma, beta

    def forward(self, x, vector, bbox):
        raise RuntimeError("use operands(): the ISLA arithmetic is fused into the following convolution")

    def __repr__(self):
        return self.__class__.__name__ + '(' + str(self.num_features) + ')'
