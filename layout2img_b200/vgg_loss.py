"""VGG19 perceptual loss of the generator step (reference utils/util.py:49-94, used at train_context_app_v2.py:185-187):

    loss = sum_i w_i * L1(vgg_i(fake), vgg_i(real).detach()),   w = (1/32, 1/16, 1/8, 1/4, 1)

with vgg_i the ReLU outputs after torchvision VGG19 `features` 1, 6, 11, 20, 29.  The 13 convolutions run on the
tensor-core convolution kernel (functional.conv2d), the four max-poolings on csrc/roi_align.cu maxpool2_*.  The reference
downloads ImageNet weights (`models.vgg19(pretrained=True)`); there is no network here, so the weights come from a local
torchvision checkpoint file (`VGGLoss(weights_path=...)`, keys `features.N.weight / bias`) -- without one the layers keep
their random initialisation (useful for parity tests only).  The module is optional: bench.py's metric excludes it
(SURVEY.md section 8d)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as L
from .model.layers import Conv2d, to_nhwc

# torchvision vgg19 "E" configuration up to features[29]
_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512]
_TAPS = (0, 5, 10, 19, 28)          # conv indices whose ReLU outputs are h_relu1 .. h_relu5


class Vgg19(nn.Module):
    """state_dict-compatible with torchvision's vgg19().features[:30] (conv parameters at the same indices)."""

    def __init__(self, requires_grad: bool = False):
        super().__init__()
        layers: List[nn.Module] = []
        cin = 3
        for v in _CFG:
            if v == "M":
                layers.append(nn.Identity())                 # index of the MaxPool2d
            else:
                layers += [Conv2d(cin, v, kernel_size=3, padding=1), nn.Identity()]   # conv, index of its ReLU
                cin = v
        self.features = nn.Sequential(*layers)
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, X):                                     # X (b,3,H,W) NCHW in [-1,1] -> 5 NHWC feature maps (post-ReLU)
        f = self.features
        x = to_nhwc(X)
        outs, pre, relu_in = [], x, False
        for i, m in enumerate(f):
            if isinstance(m, Conv2d):
                pre = L.conv2d(pre, m.weight, m.bias, relu_in=relu_in)      # keeps the PRE-ReLU output; ReLU fused into the consumer
                relu_in = True
                if i in _TAPS:
                    outs.append(F.relu(pre))
                    if i == _TAPS[-1]:
                        break
            elif i in (4, 9, 18, 27):
                pre = L.maxpool2(pre)                         # max commutes with ReLU: pool the pre-activation, ReLU in the next conv
        return outs


class VGGLoss(nn.Module):
    def __init__(self, weights_path: Optional[str] = None):
        super().__init__()
        self.vgg = Vgg19()
        if weights_path is not None:
            sd = torch.load(weights_path, map_location="cpu")
            own = self.vgg.state_dict()
            self.vgg.load_state_dict({k: v for k, v in sd.items() if k in own})
        self.weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]

    def forward(self, x, y):
        x_vgg = self.vgg(x)
        with torch.no_grad():
            y_vgg = self.vgg(y)
        loss = 0
        for w, a, b in zip(self.weights, x_vgg, y_vgg):
            loss = loss + w * (a - b).abs().mean()
        return loss
