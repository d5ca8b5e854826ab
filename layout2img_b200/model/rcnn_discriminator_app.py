"""Drop-in for the reference's model/rcnn_discriminator_app.py (classes CombineDiscriminator128_app,
ResnetDiscriminator128_app, OptimizedBlock, ResBlock; helper conv2d).  Same constructors, forward
signatures and state_dict keys; arithmetic in libl2i.so (sm_100a)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as L
from ..sn_group import prepare_network
from .layers import conv2d, to_nhwc

__all__ = ["CombineDiscriminator128_app", "ResnetDiscriminator128_app", "OptimizedBlock", "ResBlock", "conv2d"]


class OptimizedBlock(nn.Module):
    """reference :294-314: pool(conv2(relu(conv1 x))) + c_sc(pool x) -- one fused autograd node (functional.DBlockFn)."""

    def __init__(self, in_ch, out_ch, ksize=3, pad=1, downsample=False):
        super().__init__()
        self.conv1 = conv2d(in_ch, out_ch, ksize, 1, pad)
        self.conv2 = conv2d(out_ch, out_ch, ksize, 1, pad)
        self.c_sc = conv2d(in_ch, out_ch, 1, 1, 0)
        self.activation = nn.ReLU()
        self.downsample = downsample

    def forward(self, in_feat):                      # NHWC
        # spectral norm (power iteration, 1/sigma) of the three convs runs inside the fused node
        return L.d_block(in_feat, self.conv1, self.conv2, self.c_sc, down=self.downsample, optimized=True)


class ResBlock(nn.Module):
    """reference :317-344: pool?(conv2(relu(conv1(relu x)))) + pool?(c_sc x).  The 1x1 shortcut commutes with
    average pooling, so it runs on the pooled input and is added after the pooling in conv2's epilogue
    (functional.DBlockFn)."""

    def __init__(self, in_ch, out_ch, ksize=3, pad=1, downsample=False):
        super().__init__()
        self.conv1 = conv2d(in_ch, out_ch, ksize, 1, pad)
        self.conv2 = conv2d(out_ch, out_ch, ksize, 1, pad)
        self.activation = nn.ReLU()
        self.downsample = downsample
        self.learnable_sc = (in_ch != out_ch) or downsample
        if self.learnable_sc:
            self.c_sc = conv2d(in_ch, out_ch, 1, 1, 0)

    def forward(self, in_feat):                      # NHWC
        return L.d_block(in_feat, self.conv1, self.conv2, self.c_sc if self.learnable_sc else None,
                         down=self.downsample, optimized=False)


class ResnetDiscriminator128_app(nn.Module):
    """reference :84-168."""

    def __init__(self, num_classes=0, input_dim=3, ch=64):
        super().__init__()
        self.num_classes = num_classes
        self.block1 = OptimizedBlock(3, ch, downsample=True)
        self.block2 = ResBlock(ch, ch * 2, downsample=True)
        self.block3 = ResBlock(ch * 2, ch * 4, downsample=True)
        self.block4 = ResBlock(ch * 4, ch * 8, downsample=True)
        self.block5 = ResBlock(ch * 8, ch * 16, downsample=True)
        self.block6 = ResBlock(ch * 16, ch * 16, downsample=False)
        self.l7 = nn.utils.spectral_norm(nn.Linear(ch * 16, 1))
        self.activation = nn.ReLU()
        self.block_obj3 = ResBlock(ch * 2, ch * 4, downsample=False)
        self.block_obj4 = ResBlock(ch * 4, ch * 8, downsample=False)
        self.block_obj5 = ResBlock(ch * 8, ch * 16, downsample=True)
        self.l_obj = nn.utils.spectral_norm(nn.Linear(ch * 16, 1))
        self.l_y = nn.utils.spectral_norm(nn.Embedding(num_classes, ch * 16))
        self.app_conv = ResBlock(ch * 8, ch * 8, downsample=False)
        self.l_y_app = nn.utils.spectral_norm(nn.Embedding(num_classes, ch * 8))
        self.app = nn.utils.spectral_norm(nn.Linear(ch * 16, 1))

    def forward(self, x, y=None, bbox=None):         # x NHWC; bbox (K,5) rois in pixels
        prepare_network(self)          # spectral norm of all 33 modules + every conv's operand pairs: one grouped call
        x = self.block1(x)
        x1 = self.block2(x)
        x2 = self.block3(x1)
        x = self.block4(x2)
        x = self.block5(x)
        x = self.block6(x)
        out_im = L.proj_head(x, self.l7)                               # sum_hw relu -> SN linear (:125-127), one kernel

        # small / large object paths (:131-146); order = all large then all small
        s_idx = ((bbox[:, 3] - bbox[:, 1]) < 64) * ((bbox[:, 4] - bbox[:, 2]) < 64)
        bbox_l, bbox_s = bbox[~s_idx], bbox[s_idx]
        y_l, y_s = y[~s_idx], y[s_idx]
        obj_feat_s = self.block_obj3(x1)
        obj_feat_s = self.block_obj4(obj_feat_s)
        obj_feat_s = L.roi_align(obj_feat_s, bbox_s, 1.0 / 4.0)
        obj_feat_l = self.block_obj4(x2)
        obj_feat_l = L.roi_align(obj_feat_l, bbox_l, 1.0 / 8.0)
        obj_feat = torch.cat([obj_feat_l, obj_feat_s], dim=0)          # (K,8,8,512) NHWC
        y = torch.cat([y_l, y_s], dim=0)

        # appearance head (:148-157): mean_i Linear([Gram_i, e_y]) without forming the (K,512,512) Gram matrix or the
        # (K,512,1024) concat (csrc/heads.cu):  sum_i Gram[i,:] . w1 = (1/C) sum_p (sum_i F[i,p]) (sum_j F[j,p] w1[j])
        out_app = L.gram_proj(self.app_conv(obj_feat), self.app, self.l_y_app, y)

        # object head (:160-166): sum_hw relu -> SN linear + <SN embedding[y], .>
        out_obj = L.proj_head(self.block_obj5(obj_feat), self.l_obj, self.l_y, y)
        return out_im, out_obj, out_app


class CombineDiscriminator128_app(nn.Module):
    """reference :396-421.  images (b,3,128,128) NCHW; bbox (b,o,4) xywh in [0,1]; label (b,o[,1])."""

    def __init__(self, num_classes=81):
        super().__init__()
        self.obD = ResnetDiscriminator128_app(num_classes=num_classes, input_dim=3)

    def forward(self, images, bbox, label, mask=None):
        if not images.is_cuda:
            raise RuntimeError("layout2img_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        dev = images.device
        b, o = bbox.size(0), bbox.size(1)
        idx = torch.arange(start=0, end=b, device=dev).view(b, 1, 1).expand(-1, o, -1).float()
        bbox = bbox.to(dev).float().clone()          # the reference edits a GPU-resident bbox in place (:408-409)
        bbox[:, :, 2] = bbox[:, :, 2] + bbox[:, :, 0]
        bbox[:, :, 3] = bbox[:, :, 3] + bbox[:, :, 1]
        bbox = bbox * images.size(2)
        bbox = torch.cat((idx, bbox), dim=2).view(-1, 5)
        label = label.to(dev).view(-1)
        keep = (label != 0).nonzero().view(-1)
        bbox = bbox[keep]
        label = label[keep]
        return self.obD(to_nhwc(images), label, bbox)
