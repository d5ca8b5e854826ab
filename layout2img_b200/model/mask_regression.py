"""Drop-in for the reference's model/mask_regression.py:58-102 (MaskRegressNetv2)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as L
from .layers import Conv2d, to_nchw_view, to_nhwc


class MaskRegressNetv2(nn.Module):
    def __init__(self, obj_feat=128, mask_size=16, map_size=64):
        super().__init__()
        self.mask_size = mask_size
        self.map_size = map_size
        self.fc = nn.utils.spectral_norm(nn.Linear(obj_feat, 256 * 4 * 4))
        self.conv1 = nn.Sequential(nn.utils.spectral_norm(Conv2d(256, 256, 3, 1, 1)), nn.InstanceNorm2d(256), nn.ReLU())
        self.conv2 = nn.Sequential(nn.utils.spectral_norm(Conv2d(256, 256, 3, 1, 1)), nn.InstanceNorm2d(256), nn.ReLU())
        self.conv3 = nn.Sequential(nn.utils.spectral_norm(Conv2d(256, 256, 3, 1, 1)), nn.InstanceNorm2d(256), nn.ReLU(),
                                   nn.utils.spectral_norm(Conv2d(256, 1, 1, 1)), nn.Sigmoid())

    def forward(self, obj_feat, bbox):
        b, num_o, _ = bbox.size()
        x = L.sn_linear(self.fc, obj_feat.view(b * num_o, -1))
        x = to_nhwc(x.view(b * num_o, 256, 4, 4))
        # three (conv3x3 -> InstanceNorm -> ReLU [-> bilinear x2]) stages + the 1x1 logit conv: one fused node
        x = torch.sigmoid(L.mask_trunk(x, self.conv1[0], self.conv2[0], self.conv3[0], self.conv3[3]))   # (b*o,16,16,1)
        x = x.view(b, num_o, self.mask_size, self.mask_size)
        return L.masks_to_layout(x, bbox.to(x.device).float(), self.map_size)
