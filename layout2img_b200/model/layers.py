"""Building blocks shared by the drop-in generator and discriminator.

Activations inside the modules are contiguous fp32 (N,H,W,C) tensors; the public forward()s of
the top-level modules take and return the reference's NCHW shapes.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as L


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 2, 3, 1).contiguous()


def to_nchw_view(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 1, 2)


class Conv2d(nn.Conv2d):
    """nn.Conv2d whose arithmetic is the tcgen05 implicit-GEMM kernel (csrc/conv_tc.cu).  Keeps
    nn.Conv2d's parameters/state_dict so nn.utils.spectral_norm wraps it exactly like the
    reference's conv2d() helper (resnet_generator_app_v2.py:681-686).  Input/outputs are NHWC.

    forward(x, residual=None, relu_in=False, up2_in=False, res_up2=False, norm=None)
      norm = (bn_module, mask_pm, gamma, beta): fuse batch-norm/ISLA + ReLU (+ nearest x2) in front.
    """

    def forward(self, x, residual=None, relu_in=False, up2_in=False, res_up2=False, norm=None, chan_scale=None):
        if self.stride != (1, 1) or self.kernel_size not in ((3, 3), (1, 1)) or self.dilation != (1, 1) or self.groups != 1:
            raise ValueError("layout2img_b200 Conv2d supports 3x3/pad 1 and 1x1/pad 0, stride 1 only")
        if norm is None:
            return L.conv2d(x, self.weight, self.bias, residual, relu_in, up2_in, res_up2)   # plain (non-SN) use
        bn, mask_pm, gamma, beta = norm
        w, b, sn = L._sn_of(self)             # weight_orig + (u, v, eps, training) when spectrally normalised
        return L.norm_conv(x, w, b, bn.running_mean, bn.running_var, bn.training,
                           mask_pm=mask_pm, gamma=gamma, beta=beta, aff_w=bn.weight if bn.affine else None,
                           aff_b=bn.bias if bn.affine else None, residual=residual, up2=up2_in, res_up2=res_up2,
                           momentum=bn.momentum, eps=bn.eps, sn=sn, chan_scale=chan_scale)


class SynchronizedBatchNorm2d(nn.BatchNorm2d):
    """State holder with nn.BatchNorm2d's keys (reference model/sync_batchnorm/batchnorm.py).  The
    arithmetic runs inside the fused norm+conv kernels; the single-device reference branch calls
    F.batch_norm directly and therefore never advances num_batches_tracked -- neither do we.
    Called directly (NHWC input) it is a plain batch norm through the same kernels."""

    def forward(self, x):
        raise RuntimeError("SynchronizedBatchNorm2d is consumed by the fused norm+conv kernels")


BatchNorm = SynchronizedBatchNorm2d


def conv2d(in_feat, out_feat, kernel_size=3, stride=1, pad=1, spectral_norm=True):
    conv = Conv2d(in_feat, out_feat, kernel_size, stride, pad)
    if spectral_norm:
        return nn.utils.spectral_norm(conv, eps=1e-4)
    return conv
