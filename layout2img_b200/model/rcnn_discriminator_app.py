"""Drop-in for the reference's model/rcnn_discriminator_app.py (classes CombineDiscriminator128_app,
ResnetDiscriminator128_app, OptimizedBlock, ResBlock; helper conv2d).  Same constructors, forward
signatures and state_dict keys; arithmetic in libl2i.so (sm_100a)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as L
from .. import ops
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

    def forward(self, x, y=None, bbox=None, level=None):   # x NHWC; bbox (K,5) rois in pixels, level (K,) 0 large / 1 small / 2 dropped
        prepare_network(self)          # spectral norm of all 33 modules + every conv's operand pairs: one grouped call
        x = self.block1(x)
        x1 = self.block2(x)
        x2 = self.block3(x1)
        x = self.block4(x2)
        x = self.block5(x)
        x = self.block6(x)
        out_im = L.proj_head(x, self.l7)                               # sum_hw relu -> SN linear (:125-127), one kernel

        # small / large object paths (:131-146): rois arrive ordered [all large, all small] with their level; one launch
        # samples both feature maps into the (K,8,8,512) stack the reference builds with two RoIAligns and a cat
        obj_feat_s = self.block_obj4(self.block_obj3(x1))
        obj_feat_l = self.block_obj4(x2)
        obj_feat = L.roi_align2(obj_feat_l, obj_feat_s, bbox, level)

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
        # False (default): outputs have one row per valid object, as the reference returns them (one host sync per call).
        # True: fixed b*o rows, dropped objects as zero-feature rows flagged in `valid_mask` -- no host sync, CUDA-graph safe.
        self.static_shapes = False
        self.valid_mask = None
        self.valid_count = None

    def forward(self, images, bbox, label, mask=None):
        if not images.is_cuda:
            raise RuntimeError("layout2img_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        dev = images.device
        b, o = bbox.size(0), bbox.size(1)
        # :402-417 and the small / large partition of :131-134 on the device (csrc/roi_align.cu roi_prepare): bit-exact
        # rois in the reference's order [large ..., small ...], dropped (label 0) rows last, counts on the device
        rois, y, level, _, counts = ops.roi_prepare(bbox.to(dev).float().contiguous(),
                                                    label.to(dev).to(torch.int64).reshape(-1).contiguous(), float(images.size(2)))
        if self.static_shapes:
            # graph-capturable form: all b*o rows are processed (dropped rows as zero features), no host round trip;
            # `valid_mask` (b*o,) marks the rows the losses may use, `valid_count` their number (device scalars)
            self.valid_mask = level < 2
            self.valid_count = counts.sum()
        else:
            k = int(counts.sum())                     # the one host synchronisation of the call: output rows = valid objects
            rois, y, level = rois[:k], y[:k], level[:k]
            self.valid_mask = self.valid_count = None
        return self.obD(to_nhwc(images), y, rois, level)
