"""Drop-in for the reference's model/resnet_generator_app_v2.py (classes ResnetGenerator128_context,
ResnetGenerator128, ResBlock, BoxMultiHeadedAttention, PSPModule; helpers conv2d, bbox_mask,
batched_index_select, BatchNorm).  Same constructors, forward signatures, state_dict keys and
initialisation order; the arithmetic runs in libl2i.so (sm_100a) through layout2img_b200.functional.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as L
from .. import ops
from ..sn_group import prepare_network
from .layers import BatchNorm, Conv2d, SynchronizedBatchNorm2d, conv2d, to_nchw_view, to_nhwc
from .mask_regression import MaskRegressNetv2
from .norm_module import SpatialAdaptiveSynBatchNorm2d

__all__ = ["ResnetGenerator128_context", "ResnetGenerator128", "ResBlock", "BoxMultiHeadedAttention", "PSPModule",
           "conv2d", "bbox_mask", "batched_index_select", "BatchNorm", "SynchronizedBatchNorm2d"]


def clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class BoxMultiHeadedAttention(nn.Module):
    """reference :123-214 (h = 1).  Projections (csrc/linear.cu) and residual LayerNorms are libl2i kernels; the relational
    embedding, geometry gate, masked softmax and PV product are one kernel (csrc/attention.cu)."""

    def __init__(self, h, d_model, trignometric_embedding=True, legacy_extra_skip=False, dropout=0.1):
        super().__init__()
        assert d_model % h == 0
        if h != 1 or not trignometric_embedding:
            raise ValueError("layout2img_b200 implements the configuration the generator uses: h=1, trigonometric embedding")
        self.h, self.d_k, self.d_v, self.dim_g = h, d_model // h, d_model // h, 64
        self.linears = clones(nn.Linear(d_model, d_model), 4)
        self.WGs = clones(nn.Linear(self.dim_g, 1, bias=True), h)
        self.layer_norm = nn.LayerNorm(d_model)
        self.layer_norm0 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, input_query, input_key, input_value, input_box, mask=None):
        b, o, d = input_query.shape
        q, k, v = [L.linear(x, l.weight, l.bias) for l, x in zip(self.linears, (input_query, input_key, input_value))]
        if mask is None:
            mask = torch.ones((b, o), dtype=torch.int64, device=q.device)
        x = L.box_attention(q, k, v, input_box.to(q.device).float(), mask.to(torch.int64).contiguous(),
                            self.WGs[0].weight, self.WGs[0].bias)
        # reference :197-198: transpose then *view* -- reinterprets the (d,o) matrix as (o,d); kept.
        x = x.transpose(1, 2).contiguous().view(b, -1, self.h * self.d_k)
        output = L.add_layer_norm(x, input_query, self.layer_norm0)           # LayerNorm(x + residual), one kernel
        new_residual = output
        output = self.dropout(L.linear(output, self.linears[-1].weight, self.linears[-1].bias))
        return L.add_layer_norm(output, new_residual, self.layer_norm)


class PSPModule(nn.Module):
    """reference :724-752.  Pools, stage 1x1 convolutions, their BatchNorm + ReLU, the up-sampling / concatenation, the
    528 -> 100 3x3 bottleneck convolution and its BatchNorm / ReLU / Dropout2d all run in libl2i.so (csrc/psp.cu,
    csrc/linear.cu, csrc/isla.cu affine form, csrc/conv_tc.cu); only the Dropout2d keep-mask is drawn by torch's RNG."""

    def __init__(self, features, out_features=512, sizes=(1, 2, 3, 6)):
        super().__init__()
        if tuple(sizes) != (1, 2, 3, 6):
            raise ValueError("layout2img_b200 PSPModule implements the reference's pyramid sizes (1, 2, 3, 6)")
        self.sizes = tuple(sizes)
        self.stages = nn.ModuleList([self._make_stage(features, out_features, size) for size in sizes])
        self.bottleneck = nn.Sequential(
            Conv2d(features + len(sizes) * out_features, out_features, kernel_size=3, padding=1, dilation=1, bias=False),
            BatchNorm(out_features), nn.ReLU(), nn.Dropout2d(0.1))
        self.dropout_mask = None     # tests may pin the (b, C) keep-mask (already scaled by 1/(1-p))

    def _make_stage(self, features, out_features, size):
        prior = nn.AdaptiveAvgPool2d(output_size=(size, size))
        conv = nn.Conv2d(features, out_features, kernel_size=1, bias=False)
        bn = nn.BatchNorm2d(out_features)
        return nn.Sequential(prior, conv, bn, nn.ReLU())

    def _bottleneck(self, feats):
        """-> (bottleneck conv output (b,h,w,100) before its norm, Dropout2d keep-mask (b,100) scaled by 1/(1-p) or None)."""
        b, h, w, c = feats.shape
        pooled = L.psp_pool(feats)                                      # (b, 50, c): all four adaptive pools
        priors, off = [], 0
        for stage, s in zip(self.stages, self.sizes):
            p = L.linear(pooled[:, off:off + s * s], stage[1].weight.view(stage[1].out_channels, c))   # 1x1 conv (csrc/linear.cu)
            # nn.BatchNorm2d over the b*s*s samples + ReLU (csrc/isla.cu affine form); the reference never synchronises
            # these plain BatchNorm2d layers across GPUs
            priors.append(L.bn_relu(p, stage[2], sync=False))
            if stage[2].training and stage[2].track_running_stats:
                stage[2].num_batches_tracked.add_(1)
            off += s * s
        x = L.psp_bottleneck(feats, torch.cat(priors, dim=1), self.bottleneck[0].weight)   # (b,h,w,100)
        keep = None
        if self.training:
            if self.dropout_mask is not None:
                keep = self.dropout_mask.to(x).view(b, -1).contiguous()
            else:
                keep = (torch.rand((b, x.shape[-1]), device=x.device) >= 0.1).to(x.dtype) / 0.9
        return x, keep

    def forward(self, feats):                       # feats NHWC -> BatchNorm + ReLU + Dropout2d of the bottleneck (b,h,w,100)
        x, keep = self._bottleneck(feats)
        return L.bn_relu(x, self.bottleneck[1], chan_scale=keep)        # relu(y) * k == relu(y * k) for k >= 0

    def forward_head(self, feats, conv):
        """PSP head + the following 1x1 convolution (`conv_mask[1]`): the bottleneck's BatchNorm / ReLU / Dropout2d are
        applied while the convolution's operand pair is written (functional.NormConvFn)."""
        x, keep = self._bottleneck(feats)
        return conv.forward(x, norm=(self.bottleneck[1], None, None, None), chan_scale=keep)


class ResBlock(nn.Module):
    """reference :628-678.  b1/b2 + ReLU + nearest-x2 are fused in front of conv1/conv2; the 1x1
    shortcut runs at the low resolution (it commutes with nearest up-sampling) and is added in
    conv2's epilogue."""

    def __init__(self, in_ch, out_ch, h_ch=None, ksize=3, pad=1, upsample=False, num_w=128, predict_mask=True,
                 psp_module=False):
        super().__init__()
        self.upsample = upsample
        self.h_ch = h_ch if h_ch else out_ch
        self.conv1 = conv2d(in_ch, self.h_ch, ksize, pad=pad)
        self.conv2 = conv2d(self.h_ch, out_ch, ksize, pad=pad)
        self.b1 = SpatialAdaptiveSynBatchNorm2d(in_ch, num_w=num_w, batchnorm_func=BatchNorm)
        self.b2 = SpatialAdaptiveSynBatchNorm2d(self.h_ch, num_w=num_w, batchnorm_func=BatchNorm)
        self.learnable_sc = in_ch != out_ch or upsample
        if self.learnable_sc:
            self.c_sc = conv2d(in_ch, out_ch, 1, 1, 0)
        self.activation = nn.ReLU()
        self.predict_mask = predict_mask
        self.psp = psp_module
        if self.predict_mask:
            if psp_module:
                self.conv_mask = nn.Sequential(PSPModule(out_ch, 100), Conv2d(100, 184, kernel_size=1))
            else:
                self.conv_mask = nn.Sequential(Conv2d(out_ch, 100, 3, 1, 1), BatchNorm(100), nn.ReLU(),
                                               Conv2d(100, 184, 1, 1, 0, bias=True))
            self.head_conv._l2i_gathered_head = True

    def forward(self, in_feat, w, bbox, head_features=False):   # in_feat NHWC, bbox (b,o,hm,wm)
        if not (self.upsample and self.learnable_sc):
            raise ValueError("layout2img_b200 ResBlock implements the generator's configuration: upsample=True")
        b, h, wd, _ = in_feat.shape
        o = bbox.shape[1]
        # ISLA operands (norm_module.py:173-180): masks bilinearly resized to each norm's resolution, pixel-major;
        # gamma / beta = spectrally normalised linear projections of the object latents
        m1 = L.mask_resize(bbox, h, wd, True)
        m2 = L.mask_resize(bbox, 2 * h, 2 * wd, True)
        g1, be1 = L.sn_linear(self.b1.weight_proj, w).view(b, o, -1), L.sn_linear(self.b1.bias_proj, w).view(b, o, -1)
        g2, be2 = L.sn_linear(self.b2.weight_proj, w).view(b, o, -1), L.sn_linear(self.b2.bias_proj, w).view(b, o, -1)
        out_feat = L.g_block(in_feat, m1, g1, be1, m2, g2, be2, self.conv1, self.conv2, self.c_sc,
                             self.b1.batch_norm2d, self.b2.batch_norm2d)
        if not self.predict_mask:
            return out_feat, None
        if head_features:
            # the mask head up to (not including) its final 1x1 convolution: (b,h,w,100) after BatchNorm / ReLU (/ Dropout2d);
            # the generator forms only the o class channels it consumes (functional.class_mix with `self.head_conv`)
            if self.psp:
                return out_feat, self.conv_mask[0](out_feat)
            return out_feat, L.bn_relu(self.conv_mask[0](out_feat), self.conv_mask[1])
        if self.psp:
            mask = self.conv_mask[0].forward_head(out_feat, self.conv_mask[1])
        else:
            t = self.conv_mask[0](out_feat)
            mask = self.conv_mask[3](t, norm=(self.conv_mask[1], None, None, None))
        return out_feat, mask                         # mask NHWC (b,h,w,184)

    @property
    def head_conv(self):
        """The final 1x1 convolution (100 -> 184) of the mask head."""
        return self.conv_mask[1] if self.psp else self.conv_mask[3]


def batched_index_select(input, dim, index):
    expanse = list(input.shape)
    expanse[0] = -1
    expanse[dim] = -1
    return torch.gather(input, dim, index.expand(expanse))


def bbox_mask(x, bbox, H, W):
    """reference :697-721 -- bit-exact {0,1} box map (csrc/layout_ops.cu)."""
    return ops.bbox_mask(bbox.to(x.device).float().contiguous(), H, W)


class _GeneratorBase(nn.Module):
    context_attention = True

    def __init__(self, ch=64, z_dim=128, num_classes=10, output_dim=3):
        super().__init__()
        self.num_classes = num_classes
        self.label_embedding = nn.Embedding(num_classes, 180)
        num_w = 128 + 180
        if self.context_attention:
            self.context = BoxMultiHeadedAttention(1, num_w, dropout=0.0)
        self.fc = nn.utils.spectral_norm(nn.Linear(z_dim, 4 * 4 * 16 * ch))
        self.res1 = ResBlock(ch * 16, ch * 16, upsample=True, num_w=num_w)
        self.res2 = ResBlock(ch * 16, ch * 8, upsample=True, num_w=num_w)
        self.res3 = ResBlock(ch * 8, ch * 4, upsample=True, num_w=num_w)
        self.res4 = ResBlock(ch * 4, ch * 2, upsample=True, num_w=num_w, psp_module=True)
        self.res5 = ResBlock(ch * 2, ch * 1, upsample=True, num_w=num_w, predict_mask=False)
        self.final = nn.Sequential(BatchNorm(ch), nn.ReLU(), conv2d(ch, output_dim, 3, 1, 1), nn.Tanh())
        self.mapping = nn.Sequential()
        self.alpha1 = nn.Parameter(torch.zeros(1, 184, 1))
        self.alpha2 = nn.Parameter(torch.zeros(1, 184, 1))
        self.alpha3 = nn.Parameter(torch.zeros(1, 184, 1))
        self.alpha4 = nn.Parameter(torch.zeros(1, 184, 1))
        self.sigmoid = nn.Sigmoid()
        self.mask_regress = MaskRegressNetv2(num_w)
        self.init_parameter()

    def forward(self, z, bbox, z_im=None, y=None):
        b, o = z.size(0), z.size(1)
        dev = z.device
        if not z.is_cuda:
            raise RuntimeError("layout2img_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        bbox = bbox.to(dev).float().contiguous()
        y = y.to(dev).to(torch.int64).contiguous()
        prepare_network(self)          # spectral norm of all 42 modules + every conv's operand pairs: one grouped call
        label_embedding = self.label_embedding(y)
        latent_vector = torch.cat((z.reshape(b * o, -1), label_embedding.view(b * o, -1)), dim=1).view(b, o, -1)
        w = self.mapping(latent_vector.view(b * o, -1)).view(b, o, -1)
        if self.context_attention:
            w = self.context(w, w, w, bbox, y)
        w = w.reshape(b * o, -1)
        bmask = self.mask_regress(w, bbox)
        if z_im is None:
            z_im = torch.randn((b, 128), device=dev)
        hard = bbox_mask(z, bbox, 64, 64)
        x = to_nhwc(L.sn_linear(self.fc, z_im).view(b, -1, 4, 4))
        # stage masks (:466-470): only the o class channels of each image's objects are formed (gathered 1x1 head fused with
        # the mixing, csrc/layout_ops.cu class_mix_*) instead of the reference's 184-channel stage mask + gather
        stage_bbox = bmask
        for res, alpha in ((self.res1, self.alpha1), (self.res2, self.alpha2), (self.res3, self.alpha3), (self.res4, self.alpha4)):
            x, t = res(x, w, stage_bbox, head_features=True)
            stage_bbox = L.class_mix(t, res.head_conv, alpha, bmask, y, hard)
        x, _ = self.res5(x, w, stage_bbox)
        x = self.final[2].forward(x, norm=(self.final[0], None, None, None))   # BN -> ReLU -> conv fused (.forward: the
        # spectral-norm hook is bypassed, the normalisation runs in csrc/specnorm.cu via L.sn_weight)
        return to_nchw_view(torch.tanh(x))

    def init_parameter(self):
        for k in self.named_parameters():
            if k[1].dim() > 1:
                torch.nn.init.orthogonal_(k[1])
            if k[0][-4:] == 'bias':
                torch.nn.init.constant_(k[1], 0)


class ResnetGenerator128_context(_GeneratorBase):
    """reference :400-506."""
    context_attention = True


class ResnetGenerator128(_GeneratorBase):
    """reference :299-397 (same network without the object-context attention)."""
    context_attention = False
