"""torch.autograd.Function wrappers: the only place where the C-ABI kernels meet autograd.

Each Function's forward/backward is a short sequence of libl2i.so launches (ops.py) plus views;
the formulas are in the kernels' headers.  Activations are contiguous fp32 (N,H,W,C) tensors.
"""
from __future__ import annotations

import torch

from . import ops
from .sn_group import take_prepared


def _c(t):
    return t if t is None or t.is_contiguous() else t.contiguous()


def _sigma(weight, sn):
    """SNState of a spectrally normalised weight (sn = (u, v, eps, training)) or None: taken from the network-level
    grouped launch (sn_group.SNGroup.prepare) when this forward prepared it, else computed by the per-module kernels."""
    pre = take_prepared(weight)
    if pre is not None and (pre[0] is not None) == (sn is not None):
        return pre[0]
    return ops.sn_sigma(_c(weight), sn[0], sn[1], training=sn[3], eps=sn[2]) if sn else None


def small_in(weight) -> bool:
    """3x3 convolution with <= 4 input channels: run as a 1x1 convolution over the 9C-channel im2col tensor (csrc/im2col.cu)."""
    return weight.shape[1] <= 4 and tuple(weight.shape[2:]) == (3, 3)


def small_out(weight) -> bool:
    """3x3 convolution with <= 4 output channels: 1x1 convolution to 9 * Cout channels + col2im gather (csrc/im2col.cu)."""
    return weight.shape[0] <= 4 and tuple(weight.shape[2:]) == (3, 3) and weight.shape[1] > 4


def _sigma_and_prep(weight, sn, need_dgrad, as_1x1=False):
    """(SNState | None, WeightPair) of a convolution weight, from the grouped launch when available.  The power
    iteration must run exactly once per module call: a prepared state is always used; only missing operand pairs
    (e.g. the data-gradient pair after a no-grad preparation) are produced here.  as_1x1: the (Cout, C, 3, 3) weight of
    a small-input convolution is prepared as the (Cout, 9C, 1, 1) weight of the im2col form (same memory)."""
    w = _c(weight)
    if as_1x1 or w.dim() == 2:
        w = w.view(w.shape[0], -1, 1, 1)
    pre = take_prepared(weight)
    if pre is not None and (pre[0] is not None) == (sn is not None):
        st, wp = pre
        if wp is None or (need_dgrad and wp.d_hi is None) or wp.taps != w.shape[2] * w.shape[3]:
            wp = ops.conv_weight_prep(w, st.sigma if st else None, need_dgrad=need_dgrad)
        return st, wp
    st = ops.sn_sigma(_c(weight), sn[0], sn[1], training=sn[3], eps=sn[2]) if sn else None
    return st, ops.conv_weight_prep(w, st.sigma if st else None, need_dgrad=need_dgrad)


def _dw_to_torch(dw, cout, cin, taps):
    k = 3 if taps == 9 else 1
    return dw.view(cout, k, k, cin).permute(0, 3, 1, 2)


def _sum2x2(t):
    n, h2, w2, c = t.shape
    return t.view(n, h2 // 2, 2, w2 // 2, 2, c).sum(dim=(2, 4))


class ConvFn(torch.autograd.Function):
    """y = conv(up2?(relu?(x)), W) + bias + residual   (3x3 pad 1 or 1x1, stride 1).
    reference: conv2d() helpers, resnet_generator_app_v2.py:681-686, rcnn_discriminator_app.py:10-15.
    sn = None (weight is the weight itself) or (u, v, eps, training): weight is weight_orig and the spectral
    normalisation runs in csrc/specnorm.cu."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, relu_in, up2_in, res_up2, sn):
        x = _c(x)
        cout, cin, kh, kw = weight.shape
        taps = kh * kw
        im2col = small_in(weight) and not relu_in and not up2_in
        if im2col:                       # 3 -> Cout: one K chunk over the 27-channel im2col pair instead of nine padded taps
            xp, _ = ops.im2col3(x, 1)
            st, wp = _sigma_and_prep(weight, sn, ctx.needs_input_grad[0], as_1x1=True)
            taps = 1
        else:
            xp = ops.act_split(x, relu=relu_in, up2=up2_in)
            st, wp = _sigma_and_prep(weight, sn, ctx.needs_input_grad[0])
        out, _ = ops.conv2d_fwd(xp, wp.f_hi, wp.f_lo, cout, taps, bias=_c(bias), residual=_c(residual), res_up2=res_up2)
        ctx.save_for_backward(xp.hi, xp.lo, wp.d_hi, wp.d_lo, weight, *(st or (None, None, None)))
        ctx.meta = (cout, cin, taps, relu_in, up2_in, res_up2, bias is not None, residual is not None, im2col)
        return out

    @staticmethod
    def backward(ctx, dout):
        xhi, xlo, dhi, dlo, weight, sg, u, v = ctx.saved_tensors
        cout, cin, taps, relu_in, up2_in, res_up2, has_bias, has_res, im2col = ctx.meta
        dout = _c(dout)
        # one read of dout: operand pair + per-channel sums (bias gradient)
        dyp, _, colsum = ops.grad_split(dout, want_lo=True, up=False)
        dx = dw = db = dres = None
        kin = 9 * cin if im2col else cin                  # channels of the saved operand pair
        if ctx.needs_input_grad[0]:
            # ReLU derivative from the saved (ReLU'd) input pair, 2x2 sum = backward of the nearest x2, both in the epilogue
            dx, _ = ops.conv2d_fwd(dyp, dhi, dlo, kin, taps, mask_hi=xhi if relu_in else None, pool=2 if up2_in else 0)
            if im2col:
                dx = ops.col2im3(dx, cin, 1)
        if ctx.needs_input_grad[1]:
            g = ops.conv2d_wgrad(dyp, ops.Pair(xhi, xlo, kin), taps)
            dw = (ops.sn_weight_grad(g, weight, ops.SNState(sg, u, v)) if sg is not None
                  else _dw_to_torch(g, cout, kin, taps).reshape(weight.shape))
        if has_bias and ctx.needs_input_grad[2]:
            db = colsum
        if has_res and ctx.needs_input_grad[3]:
            dres = _sum2x2(dout) if res_up2 else dout
        return dx, dw, db, dres, None, None, None, None


def conv2d(x, weight, bias=None, residual=None, relu_in=False, up2_in=False, res_up2=False, sn=None):
    return ConvFn.apply(x, weight, bias, residual, relu_in, up2_in, res_up2, sn)


class DBlockFn(torch.autograd.Function):
    """A whole discriminator residual block (reference rcnn_discriminator_app.py:294-344) as one autograd node:

        OptimizedBlock:  pool(conv2(relu(conv1(x)))) + c_sc(pool(x))
        ResBlock:        pool?(conv2(relu(conv1(relu(x))))) + pool?(c_sc(x))   |   ... + x  (no learnable shortcut)

    Forward = 1 operand-preparation kernel + 3 tensor-core convolutions; the 1x1 shortcut runs on the pooled
    input (it commutes with average pooling) and its result is added after the 2x2 pooling in conv2's
    epilogue; the intermediate activation only ever exists as the ReLU'd bf16 pair conv2 consumes.
    Backward = 1 gradient-split kernel (+ bias gradients), 3 weight-gradient and 3 data-gradient launches;
    the ReLU derivatives are applied in the data-gradient epilogues from the saved pairs."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, wsc, bsc, down, optimized, sn1, sn2, snsc):
        # sn* = None (w* is the weight itself) or (u, v, eps, training): w* is weight_orig and the spectral
        # normalisation (power iteration + 1/sigma) happens here, fused into the operand preparation
        x = _c(x)
        has_sc = wsc is not None
        if down and not has_sc:
            raise ValueError("a down-sampling block needs its 1x1 shortcut")
        need_dx = ctx.needs_input_grad[0]
        c1, cin = w1.shape[0], w1.shape[1]
        c2 = w2.shape[0]
        im2col = optimized and small_in(w1)       # the image block: conv1 as a 1x1 convolution over the 27-channel im2col pair
        b_mode = ((2 if down else 1) if has_sc else 0)
        if im2col:
            a0, _ = ops.im2col3(x, 1)
            _, s0 = ops.act_split2(x, relu_a=False, b_mode=b_mode, want_a=False) if b_mode else (None, None)
        else:
            a0, s0 = ops.act_split2(x, relu_a=not optimized, b_mode=b_mode)
        st1, wp1 = _sigma_and_prep(w1, sn1, need_dx, as_1x1=im2col)
        st2, wp2 = _sigma_and_prep(w2, sn2, True)
        _, a1 = ops.conv2d_fwd(a0, wp1.f_hi, wp1.f_lo, c1, 1 if im2col else 9, bias=_c(b1), want_f32=False, want_pair=True,
                               relu_pair=True)
        if has_sc:
            stsc, wps = _sigma_and_prep(wsc, snsc, need_dx)
            sc, _ = ops.conv2d_fwd(s0, wps.f_hi, wps.f_lo, c2, 1, bias=_c(bsc))
        else:
            stsc, wps, sc = None, None, x
        out, _ = ops.conv2d_fwd(a1, wp2.f_hi, wp2.f_lo, c2, 9, bias=_c(b2), residual=sc, pool=1 if down else 0)
        none3 = (None, None, None)
        ctx.save_for_backward(a0.hi, a0.lo, s0.hi if has_sc else None, s0.lo if has_sc else None, a1.hi, a1.lo,
                              wp1.d_hi, wp1.d_lo, wp2.d_hi, wp2.d_lo, wps.d_hi if has_sc else None,
                              wps.d_lo if has_sc else None, w1, w2, wsc, *(st1 or none3), *(st2 or none3),
                              *(stsc or none3))
        ctx.meta = (cin, c1, c2, down, optimized, has_sc, im2col)
        return out

    @staticmethod
    def backward(ctx, dout):
        (a0h, a0l, s0h, s0l, a1h, a1l, d1h, d1l, d2h, d2l, dsh, dsl, w1, w2, wsc,
         sg1, u1, v1, sg2, u2, v2, sgs, us, vs) = ctx.saved_tensors
        cin, c1, c2, down, optimized, has_sc, im2col = ctx.meta
        need = ctx.needs_input_grad
        dout = _c(dout)
        g_lo, g_up, colsum = ops.grad_split(dout, want_lo=has_sc or not down, up=down, up_scale=0.25)
        g_full = g_up if down else g_lo                      # gradient at conv2's output resolution
        a0, a1 = ops.Pair(a0h, a0l, 9 * cin if im2col else cin), ops.Pair(a1h, a1l, c1)

        def wgrad(dy, xin, taps, w, sg, u, v):
            g = ops.conv2d_wgrad(dy, xin, taps)              # (Cout, taps, Cin)
            if sg is not None:
                return ops.sn_weight_grad(g, w, ops.SNState(sg, u, v))
            return _dw_to_torch(g, g.shape[0], g.shape[2], taps).reshape(w.shape)

        dw1 = db1 = dw2 = db2 = dwsc = dbsc = dx = None
        if need[3]:
            dw2 = wgrad(g_full, a1, 9, w2, sg2, u2, v2)
        if need[4]:
            db2 = colsum
        if has_sc and need[5]:
            dwsc = wgrad(g_lo, ops.Pair(s0h, s0l, cin), 1, wsc, sgs, us, vs)
        if has_sc and need[6]:
            dbsc = colsum.clone() if need[4] else colsum
        if need[0] or need[1] or need[2]:
            _, d1 = ops.conv2d_fwd(g_full, d2h, d2l, c1, 9, mask_hi=a1h, want_f32=False, want_pair=True)
            if need[1]:
                dw1 = wgrad(d1, a0, 1 if im2col else 9, w1, sg1, u1, v1)
            if need[2]:
                db1 = ops.pair_colsum(d1)
            if need[0]:
                if has_sc:
                    r, _ = ops.conv2d_fwd(g_lo, dsh, dsl, cin, 1)
                else:
                    r = dout
                if im2col:        # 64 -> 27 im2col-channel gradient, gathered back to the 3 image channels with the shortcut's term
                    dcol, _ = ops.conv2d_fwd(d1, d1h, d1l, 9 * cin, 1)
                    dx = ops.col2im3(dcol, cin, 1, residual=r, res_up2=down, res_scale=0.25 if down else 1.0)
                else:
                    dx, _ = ops.conv2d_fwd(d1, d1h, d1l, cin, 9, mask_hi=None if optimized else a0h, residual=r,
                                           res_up2=down, res_scale=0.25 if down else 1.0)
        return dx, dw1, db1, dw2, db2, dwsc, dbsc, None, None, None, None, None


def _sn_of(conv):
    """(weight tensor to differentiate, bias, (u, v, eps, training) | None) of a (possibly spectrally normalised)
    conv module.  For a normalised module the hook is bypassed: its arithmetic runs in csrc/specnorm.cu."""
    from torch.nn.utils.spectral_norm import SpectralNorm
    for hook in conv._forward_pre_hooks.values():
        if isinstance(hook, SpectralNorm):
            if hook.n_power_iterations != 1 or hook.dim != 0:
                raise ValueError("layout2img_b200 implements spectral_norm(n_power_iterations=1, dim=0)")
            return conv.weight_orig, getattr(conv, "bias", None), (conv.weight_u, conv.weight_v, hook.eps, conv.training)
    return conv.weight, getattr(conv, "bias", None), None


class LinearFn(torch.autograd.Function):
    """y = x @ W^T + b (W = weight_orig / sigma when spectrally normalised) for every nn.Linear of the path -- the ISLA
    gamma / beta projections (norm_module.py:158-159), mask_regression.py:64, the generator's fc
    (resnet_generator_app_v2.py:409), the attention projections (:148-151,208-212) and the PSP stages' 1x1 convolutions on
    pooled cells (:741-746) -- on the tensor-core convolution kernel: a linear layer over M rows is a 1x1 convolution over
    M images of 1 x 1 pixels, so forward, dx and dW reuse conv2d_fwd / conv2d_wgrad (fp32-class bf16-pair arithmetic) and
    the weight's operand pairs come from the network's grouped preparation.  sn = None | (u, v, eps, training)."""

    @staticmethod
    def forward(ctx, x, weight, bias, sn):
        n_out, k = weight.shape
        x2 = _c(x.reshape(-1, k))
        m = x2.shape[0]
        xp = ops.act_split(x2.view(m, 1, 1, k))
        st, wp = _sigma_and_prep(weight, sn, ctx.needs_input_grad[0], as_1x1=True)
        y, _ = ops.conv2d_fwd(xp, wp.f_hi, wp.f_lo, n_out, 1, bias=_c(bias))
        ctx.save_for_backward(xp.hi, xp.lo, wp.d_hi, wp.d_lo, weight, *(st or (None, None, None)))
        ctx.meta = (tuple(x.shape), n_out, k, bias is not None)
        return y.view(*x.shape[:-1], n_out)

    @staticmethod
    def backward(ctx, dy):
        xhi, xlo, dhi, dlo, weight, sg, u, v = ctx.saved_tensors
        xshape, n_out, k, has_bias = ctx.meta
        need = ctx.needs_input_grad
        dy2 = _c(dy.reshape(-1, n_out))
        m = dy2.shape[0]
        if ops.pad8(n_out) <= 2048:
            dyp, _, colsum = ops.grad_split(dy2.view(m, 1, 1, n_out), want_lo=True, up=False)     # pair + bias gradient in one read
        else:                                                     # fc (16384 wide), mask-regression fc (4096)
            dyp = ops.act_split(dy2.view(m, 1, 1, n_out))
            colsum = ops.colsum(dy2) if (has_bias and need[2]) else None
        dx = dw = None
        if need[0]:
            dx, _ = ops.conv2d_fwd(dyp, dhi, dlo, k, 1)
            dx = dx.view(xshape)
        if need[1]:
            g = ops.conv2d_wgrad(dyp, ops.Pair(xhi, xlo, k), 1)                 # (n_out, 1, k) = dL/d(W / sigma)
            dw = ops.sn_weight_grad(g, _c(weight), ops.SNState(sg, u, v)) if sg is not None else g.view(n_out, k)
        return dx, dw, (colsum if (has_bias and need[2]) else None), None


def linear(x, w, bias=None):
    return LinearFn.apply(x, w, bias, None)


class AddLayerNormFn(torch.autograd.Function):
    """LayerNorm(a + b) (the two residual LayerNorms of the attention block, resnet_generator_app_v2.py:201-212)."""

    @staticmethod
    def forward(ctx, a, b, w, bias, eps):
        a, b = _c(a), _c(b)
        y, stats = ops.add_layernorm_fwd(a, b, _c(w), _c(bias), eps)
        ctx.save_for_backward(a, b, w, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, b, w, stats = ctx.saved_tensors
        ds, dw, db = ops.add_layernorm_bwd(a, b, _c(w), stats, _c(dy))
        return ds, (ds if b is not None else None), dw, db, None


def add_layer_norm(a, b, ln):
    """ln: an nn.LayerNorm over the last dimension."""
    return AddLayerNormFn.apply(a, b, ln.weight, ln.bias, ln.eps)


class SNWeightFn(torch.autograd.Function):
    """W_orig / sigma as a tensor, for the small spectrally normalised weights that are consumed by library ops
    (the discriminator's class embeddings and appearance projection, rcnn_discriminator_app.py:104-109; the
    generator's RGB conv :418): power iteration / sigma / gradient in csrc/specnorm.cu, one division here."""

    @staticmethod
    def forward(ctx, w_orig, u, v, eps, training):
        st = _sigma(w_orig, (u, v, eps, training))
        ctx.save_for_backward(w_orig, st.sigma, st.u, st.v)
        return w_orig / st.sigma

    @staticmethod
    def backward(ctx, dw):
        w_orig, sigma, u, v = ctx.saved_tensors
        r = w_orig.shape[0]
        g = _c(dw).reshape(r, 1, -1)                      # taps = 1: torch layout == kernel layout
        return ops.sn_weight_grad(g, _c(w_orig), ops.SNState(sigma, u, v)).view_as(w_orig), None, None, None, None


def sn_weight(module):
    """The (normalised) weight of a module without firing its library spectral-norm hook."""
    w, _, sn = _sn_of(module)
    if sn is None:
        return w
    return SNWeightFn.apply(w, sn[0], sn[1], sn[2], sn[3])


def sn_linear(module, x):
    """Apply a (spectrally normalised) nn.Linear through SNLinearFn; plain nn.Linear modules are called as is."""
    w, b, sn = _sn_of(module)
    return LinearFn.apply(x, w, b, sn)


def d_block(x, conv1, conv2, c_sc=None, down=False, optimized=False):
    """Fused residual block of the discriminator from its conv modules (rcnn_discriminator_app.py:294-344)."""
    w1, b1, sn1 = _sn_of(conv1)
    w2, b2, sn2 = _sn_of(conv2)
    wsc, bsc, snsc = _sn_of(c_sc) if c_sc is not None else (None, None, None)
    return DBlockFn.apply(x, w1, b1, w2, b2, wsc, bsc, down, optimized, sn1, sn2, snsc)


def d_block_raw(x, w1, b1, w2, b2, wsc=None, bsc=None, down=False, optimized=False):
    """Same block from plain (already normalised) weight tensors."""
    return DBlockFn.apply(x, w1, b1, w2, b2, wsc, bsc, down, optimized, None, None, None)


class GBlockFn(torch.autograd.Function):
    """A whole up-sampling generator block (reference resnet_generator_app_v2.py:653-678, ResBlock.residual +
    shortcut) as one autograd node:

        out = conv2(relu(b2(conv1(up2(relu(b1(x))))))) + up2(c_sc(x)),   b1/b2 = ISLA norm (norm_module.py:163-186)

    Forward: 2 x (batch statistics, ISLA apply -> ReLU -> [nearest x2] -> bf16 operand pair), 3 tensor-core convs
    (the 1x1 shortcut at the low resolution, added up-sampled in conv2's epilogue).  Backward: every gradient
    tensor is read once -- the 2x2 sums that undo the up-sampling are taken in the data-gradient epilogue (conv1)
    and in the operand-preparation kernel (shortcut); bias gradients come from the same passes."""

    @staticmethod
    def forward(ctx, x, mask1, gamma1, beta1, mask2, gamma2, beta2, w1, b1, w2, b2, wsc, bsc, sn1, sn2, snsc, bn1, bn2):
        # bn* = (running_mean, running_var, training, momentum, eps); sn* as in DBlockFn
        x = _c(x)
        mask1, gamma1, beta1, mask2, gamma2, beta2 = (_c(t) for t in (mask1, gamma1, beta1, mask2, gamma2, beta2))
        cin, ch, cout = w1.shape[1], w1.shape[0], w2.shape[0]

        def stats(t, bn):
            rm, rv, training, momentum, eps = bn
            return ops.bn_batch_stats(t, rm, rv, eps, momentum) if training else ops.bn_eval_stats(rm, rv, eps)

        st1, wp1 = _sigma_and_prep(w1, sn1, True)
        st2, wp2 = _sigma_and_prep(w2, sn2, True)
        stsc, wps = _sigma_and_prep(wsc, snsc, True)
        mi1 = stats(x, bn1)
        _, a0 = ops.isla_fwd(x, mi1, mask1, gamma1, beta1, None, None, relu=True, up2=True)
        h1, _ = ops.conv2d_fwd(a0, wp1.f_hi, wp1.f_lo, ch, 9, bias=_c(b1))
        mi2 = stats(h1, bn2)
        _, a1 = ops.isla_fwd(h1, mi2, mask2, gamma2, beta2, None, None, relu=True, up2=False)
        xs = ops.act_split(x)
        sc, _ = ops.conv2d_fwd(xs, wps.f_hi, wps.f_lo, cout, 1, bias=_c(bsc))
        out, _ = ops.conv2d_fwd(a1, wp2.f_hi, wp2.f_lo, cout, 9, bias=_c(b2), residual=sc, res_up2=True)
        none3 = (None, None, None)
        ctx.save_for_backward(x, mi1, mask1, gamma1, beta1, a0.hi, a0.lo, h1, mi2, mask2, gamma2, beta2, a1.hi, a1.lo,
                              xs.hi, xs.lo, wp1.d_hi, wp1.d_lo, wp2.d_hi, wp2.d_lo, wps.d_hi, wps.d_lo, w1, w2, wsc,
                              *(st1 or none3), *(st2 or none3), *(stsc or none3))
        ctx.meta = (cin, ch, cout, bn1[2], bn2[2])
        return out

    @staticmethod
    def backward(ctx, dout):
        (x, mi1, mask1, gamma1, beta1, a0h, a0l, h1, mi2, mask2, gamma2, beta2, a1h, a1l, xsh, xsl,
         d1h, d1l, d2h, d2l, dsh, dsl, w1, w2, wsc, sg1, u1, v1, sg2, u2, v2, sgs, us, vs) = ctx.saved_tensors
        cin, ch, cout, train1, train2 = ctx.meta
        dout = _c(dout)

        def wgrad(dy, xin, taps, w, sg, u, v):
            g = ops.conv2d_wgrad(dy, xin, taps)
            if sg is not None:
                return ops.sn_weight_grad(g, w, ops.SNState(sg, u, v))
            return _dw_to_torch(g, g.shape[0], g.shape[2], taps)

        # gradient at the block output: full-resolution pair (conv2) + 2x2-summed pair (shortcut) in one read
        g_full, g_sum = ops.act_split2(dout, relu_a=False, b_mode=2, b_scale=1.0)
        db2 = ops.pair_colsum(g_sum)                          # = sum over pixels of dout = d bias of conv2 and c_sc
        dw2 = wgrad(g_full, ops.Pair(a1h, a1l, ch), 9, w2, sg2, u2, v2)
        da1, _ = ops.conv2d_fwd(g_full, d2h, d2l, ch, 9)
        dh1, dmask2, dgamma2, dbeta2, _ = ops.isla_bwd(h1, mi2, mask2, gamma2, beta2, None, None, da1, relu=True,
                                                       up2=False, train=train2)
        dh1_p, _, db1 = ops.grad_split(dh1, want_lo=True, up=False)
        dw1 = wgrad(dh1_p, ops.Pair(a0h, a0l, cin), 9, w1, sg1, u1, v1)
        da0, _ = ops.conv2d_fwd(dh1_p, d1h, d1l, cin, 9, pool=2)        # 2x2 sum = backward of the nearest x2
        dx1, dmask1, dgamma1, dbeta1, _ = ops.isla_bwd(x, mi1, mask1, gamma1, beta1, None, None, da0, relu=True,
                                                       up2=False, train=train1)
        dwsc = wgrad(g_sum, ops.Pair(xsh, xsl, cin), 1, wsc, sgs, us, vs)
        dx, _ = ops.conv2d_fwd(g_sum, dsh, dsl, cin, 1, residual=dx1)
        return (dx, dmask1, dgamma1, dbeta1, dmask2, dgamma2, dbeta2, dw1, db1, dw2, db2, dwsc, db2.clone(),
                None, None, None, None, None)


def g_block(x, mask1, gamma1, beta1, mask2, gamma2, beta2, conv1, conv2, c_sc, bn1, bn2):
    """Fused up-sampling generator block from its modules (resnet_generator_app_v2.py:628-678)."""
    w1, b1, sn1 = _sn_of(conv1)
    w2, b2, sn2 = _sn_of(conv2)
    wsc, bsc, snsc = _sn_of(c_sc)
    pack = lambda bn: (bn.running_mean, bn.running_var, bn.training, bn.momentum, bn.eps)
    return GBlockFn.apply(x, mask1, gamma1, beta1, mask2, gamma2, beta2, w1, b1, w2, b2, wsc, bsc, sn1, sn2, snsc,
                          pack(bn1), pack(bn2))


class NormConvFn(torch.autograd.Function):
    """y = conv(up2?(relu(norm(x))), W) + bias + residual with norm = ISLA (mask_pm given) or affine/plain
    batch norm (mask_pm None).  reference: ResBlock.residual resnet_generator_app_v2.py:653-663,
    SpatialAdaptiveSynBatchNorm2d norm_module.py:163-186, `final` :416-419, mask heads :645-651.
    sn = None (weight is the weight itself) or (u, v, eps, training): weight is weight_orig (spectral norm in-kernel)."""

    @staticmethod
    def forward(ctx, x, mask_pm, gamma, beta, aff_w, aff_b, weight, bias, residual, running_mean, running_var,
                training, momentum, eps, up2, res_up2, sn, chan_scale):
        x = _c(x)
        cout, cin, kh, kw = weight.shape
        taps = kh * kw
        if training:
            mi = ops.bn_batch_stats(x, running_mean, running_var, eps, momentum)
        else:
            mi = ops.bn_eval_stats(running_mean, running_var, eps)
        mask_pm, gamma, beta, chan_scale = _c(mask_pm), _c(gamma), _c(beta), _c(chan_scale)
        _, ap = ops.isla_fwd(x, mi, mask_pm, gamma, beta, _c(aff_w), _c(aff_b), relu=True, up2=up2, chan_scale=chan_scale)
        gather = small_out(weight) and residual is None
        if gather:
            # Cin -> 3 (the RGB head): a 1x1 convolution to the 27 (output channel, tap) products, then each pixel gathers
            # its nine neighbours' terms (col2im) -- instead of nine taps whose N tile would be 95 % padding
            st = _sigma(weight, sn)
            w_eff = weight.detach().permute(0, 2, 3, 1).reshape(cout * 9, cin, 1, 1).contiguous()   # [(s, tap)][l]
            wp = ops.conv_weight_prep(w_eff, st.sigma if st else None, need_dgrad=True)
            prod, _ = ops.conv2d_fwd(ap, wp.f_hi, wp.f_lo, cout * 9, 1)
            out = ops.col2im3(prod, cout, -1, bias=_c(bias))
        else:
            st, wp = _sigma_and_prep(weight, sn, True)
            out, _ = ops.conv2d_fwd(ap, wp.f_hi, wp.f_lo, cout, taps, bias=_c(bias), residual=_c(residual), res_up2=res_up2)
        ctx.save_for_backward(x, mi, mask_pm, gamma, beta, aff_w, aff_b, ap.hi, ap.lo, wp.d_hi, wp.d_lo, weight, chan_scale,
                              *(st or (None, None, None)))
        ctx.meta = (cout, cin, taps, training, up2, res_up2, bias is not None, residual is not None, gather)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, mi, mask_pm, gamma, beta, aff_w, aff_b, ahi, alo, dhi, dlo, weight, chan_scale, sg, u, v = ctx.saved_tensors
        cout, cin, taps, training, up2, res_up2, has_bias, has_res, gather = ctx.meta
        dout = _c(dout)
        if gather:
            # adjoint of the col2im gather: the (s, tap) product gradients as an operand pair, then two 1x1 launches
            dyp, colsum = ops.im2col3(dout, -1, want_colsum=has_bias)
            da, _ = ops.conv2d_fwd(dyp, dhi, dlo, cin, 1)
            g = ops.conv2d_wgrad(dyp, ops.Pair(ahi, alo, cin), 1).view(cout, 9, cin)     # [(s, tap)][l] == (Cout, taps, Cin)
        else:
            dyp, _, colsum = ops.grad_split(dout, want_lo=True, up=False)
            da, _ = ops.conv2d_fwd(dyp, dhi, dlo, cin, taps)
            g = ops.conv2d_wgrad(dyp, ops.Pair(ahi, alo, cin), taps)
        dw = ops.sn_weight_grad(g, weight, ops.SNState(sg, u, v)) if sg is not None else _dw_to_torch(g, cout, cin, taps)
        dx, dmask, dgamma, dbeta, csum = ops.isla_bwd(x, mi, mask_pm, gamma, beta, _c(aff_w), _c(aff_b), da,
                                                      relu=True, up2=up2, train=training, chan_scale=chan_scale)
        daw = dab = None
        if aff_w is not None:
            daw = csum[:, 1].float()
            dab = csum[:, 0].float()
        db = colsum if has_bias else None
        dres = None
        if has_res:
            dres = _sum2x2(dout) if res_up2 else dout
        return dx, dmask, dgamma, dbeta, daw, dab, dw, db, dres, None, None, None, None, None, None, None, None, None


def norm_conv(x, weight, bias, running_mean, running_var, training, mask_pm=None, gamma=None, beta=None,
              aff_w=None, aff_b=None, residual=None, up2=False, res_up2=False, momentum=0.1, eps=1e-5, sn=None,
              chan_scale=None):
    """chan_scale (B, C): affine form only -- the scaled Dropout2d keep-mask between the ReLU and the convolution."""
    return NormConvFn.apply(x, mask_pm, gamma, beta, aff_w, aff_b, weight, bias, residual, running_mean, running_var,
                            training, momentum, eps, up2, res_up2, sn, chan_scale)


class BnReluFn(torch.autograd.Function):
    """relu(batch_norm(x) * w + b) [* chan_scale] on a (..., C) fp32 tensor through the ISLA kernels' affine form
    (csrc/isla.cu, O = 0): the nn.BatchNorm2d + ReLU of the PSP stages (resnet_generator_app_v2.py:741-746, never
    synchronised across ranks: sync=False) and the bottleneck's BatchNorm + ReLU + Dropout2d (:736)."""

    @staticmethod
    def forward(ctx, x, aff_w, aff_b, running_mean, running_var, training, momentum, eps, chan_scale, sync):
        x = _c(x)
        x4 = x if x.dim() == 4 else x.reshape(x.shape[0], -1, 1, x.shape[-1])
        if training:
            mi = ops.bn_batch_stats(x4, running_mean, running_var, eps, momentum, sync=sync)
        else:
            mi = ops.bn_eval_stats(running_mean, running_var, eps)
        chan_scale = _c(chan_scale)
        out, _ = ops.isla_fwd(x4, mi, None, None, None, _c(aff_w), _c(aff_b), relu=False, up2=False, want_f32=True,
                              want_pair=False, chan_scale=chan_scale, relu_f32=True)
        ctx.save_for_backward(x4, mi, aff_w, aff_b, chan_scale)
        ctx.meta = (training, sync, x.shape)
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, dout):
        x4, mi, aff_w, aff_b, chan_scale = ctx.saved_tensors
        training, sync, shape = ctx.meta
        dx, _, _, _, csum = ops.isla_bwd(x4, mi, None, None, None, _c(aff_w), _c(aff_b), _c(dout).view(x4.shape), relu=True,
                                         up2=False, train=training, chan_scale=chan_scale, sync=sync)
        return dx.view(shape), csum[:, 1].float(), csum[:, 0].float(), None, None, None, None, None, None, None


def bn_relu(x, bn, chan_scale=None, sync=True):
    """x (..., C) through an nn.BatchNorm2d-like module `bn` (affine) followed by ReLU."""
    return BnReluFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.training, bn.momentum, bn.eps,
                          chan_scale, sync)


class MaskTrunkFn(torch.autograd.Function):
    """The convolutional trunk of MaskRegressNetv2 (reference model/mask_regression.py:66-99) as one autograd node:

        x (N,4,4,256) -> [conv3x3 -> InstanceNorm -> ReLU -> bilinear x2] x 2 -> conv3x3 -> InstanceNorm -> ReLU
                      -> conv1x1 (256 -> 1) -> logits (N,16,16,1)

    Each InstanceNorm/ReLU/up-sampling stage writes the next convolution's operand pair directly
    (csrc/layout_ops.cu inorm_*); the backward reads every gradient tensor once."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, w4, b4, sn1, sn2, sn3, sn4):
        x = _c(x)
        ws, bs, sns = (w1, w2, w3, w4), (b1, b2, b3, b4), (sn1, sn2, sn3, sn4)
        both = [_sigma_and_prep(w, sn, True) for w, sn in zip(ws, sns)]
        sts, wps = [b[0] for b in both], [b[1] for b in both]
        pair = ops.act_split(x)
        saved_pairs, hs, stats = [pair], [], []
        for i in range(3):
            h, _ = ops.conv2d_fwd(pair, wps[i].f_hi, wps[i].f_lo, ws[i].shape[0], 9, bias=_c(bs[i]))
            pair, st = ops.inorm_relu_fwd(h, up2=i < 2)
            hs.append(h); stats.append(st); saved_pairs.append(pair)
        logit, _ = ops.conv2d_fwd(pair, wps[3].f_hi, wps[3].f_lo, ws[3].shape[0], 1, bias=_c(bs[3]))
        flat = []
        for st in sts:
            flat += list(st) if st else [None, None, None]
        ctx.save_for_backward(*[t for p in saved_pairs for t in (p.hi, p.lo)], *hs, *stats,
                              *[t for wp in wps for t in (wp.d_hi, wp.d_lo)], *ws, *flat)
        ctx.chans = [p.C for p in saved_pairs]
        return logit

    @staticmethod
    def backward(ctx, dlogit):
        t = list(ctx.saved_tensors)
        pairs = [ops.Pair(t[2 * i], t[2 * i + 1], ctx.chans[i]) for i in range(4)]
        hs, stats = t[8:11], t[11:14]
        dws = [(t[14 + 2 * i], t[15 + 2 * i]) for i in range(4)]
        ws = t[22:26]
        sts = [ops.SNState(*t[26 + 3 * i:29 + 3 * i]) if t[26 + 3 * i] is not None else None for i in range(4)]
        need = ctx.needs_input_grad

        def wgrad(dy, xin, taps, i):
            g = ops.conv2d_wgrad(dy, xin, taps)
            return ops.sn_weight_grad(g, ws[i], sts[i]) if sts[i] is not None else _dw_to_torch(g, g.shape[0], g.shape[2], taps)

        grads_w, grads_b = [None] * 4, [None] * 4
        dyp, _, colsum = ops.grad_split(_c(dlogit), want_lo=True, up=False)
        grads_b[3] = colsum
        if need[7]:
            grads_w[3] = wgrad(dyp, pairs[3], 1, 3)
        da, _ = ops.conv2d_fwd(dyp, dws[3][0], dws[3][1], ctx.chans[3], 1)
        dx = None
        for i in (2, 1, 0):
            dh = ops.inorm_relu_bwd(hs[i], stats[i], da, up2=i < 2)
            dyp, _, colsum = ops.grad_split(dh, want_lo=True, up=False)
            grads_b[i] = colsum
            if need[1 + 2 * i]:
                grads_w[i] = wgrad(dyp, pairs[i], 9, i)
            if i > 0 or need[0]:
                da, _ = ops.conv2d_fwd(dyp, dws[i][0], dws[i][1], ctx.chans[i], 9)
                if i == 0:
                    dx = da
        return (dx, grads_w[0], grads_b[0], grads_w[1], grads_b[1], grads_w[2], grads_b[2], grads_w[3], grads_b[3],
                None, None, None, None)


def mask_trunk(x, conv1, conv2, conv3, conv4):
    """x (N,4,4,256) NHWC -> mask logits (N,16,16,1) through the four conv modules of MaskRegressNetv2."""
    args, sns = [], []
    for m in (conv1, conv2, conv3, conv4):
        w, b, sn = _sn_of(m)
        args += [w, b]
        sns.append(sn)
    return MaskTrunkFn.apply(x, *args, *sns)


class PspPoolFn(torch.autograd.Function):
    """The four AdaptiveAvgPool2d of PSPModule (resnet_generator_app_v2.py:741-746) in one pass: (B,H,W,C) -> (B,50,C)."""

    @staticmethod
    def forward(ctx, feats):
        feats = _c(feats)
        ctx.shape = tuple(feats.shape)
        return ops.psp_pool_fwd(feats)

    @staticmethod
    def backward(ctx, dpooled):
        return ops.psp_pool_bwd(_c(dpooled), ctx.shape)


def psp_pool(feats):
    return PspPoolFn.apply(feats)


class PspBottleneckFn(torch.autograd.Function):
    """conv3x3(cat[up(priors_1..6), feats]) of PSPModule (resnet_generator_app_v2.py:736,748-751): the concat is
    written directly as the convolution's bf16 operand pair; priors (B,50,CP) are the ReLU'd stage outputs."""

    @staticmethod
    def forward(ctx, feats, priors, weight):
        feats, priors = _c(feats), _c(priors)
        cout = weight.shape[0]
        pair = ops.psp_concat_fwd(feats, priors)
        _, wp = _sigma_and_prep(weight, None, True)
        out, _ = ops.conv2d_fwd(pair, wp.f_hi, wp.f_lo, cout, 9)
        ctx.save_for_backward(pair.hi, pair.lo, wp.d_hi, wp.d_lo)
        ctx.meta = (cout, pair.C, priors.shape[2])
        return out

    @staticmethod
    def backward(ctx, dout):
        phi, plo, dhi, dlo = ctx.saved_tensors
        cout, ctot, cp = ctx.meta
        dyp = ops.act_split(_c(dout))
        dcat, _ = ops.conv2d_fwd(dyp, dhi, dlo, ctot, 9)
        dw = _dw_to_torch(ops.conv2d_wgrad(dyp, ops.Pair(phi, plo, ctot), 9), cout, ctot, 9)
        dpriors = ops.psp_concat_bwd(dcat, cp)
        return dcat[..., 4 * cp:], dpriors, dw


def psp_bottleneck(feats, priors, weight):
    return PspBottleneckFn.apply(feats, priors, weight)


class AvgPool2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.avgpool2_fwd(_c(x))

    @staticmethod
    def backward(ctx, dout):
        return ops.avgpool2_bwd(_c(dout))


def avgpool2(x):
    return AvgPool2Fn.apply(x)


class MaxPool2Fn(torch.autograd.Function):
    """F.max_pool2d(x, 2) on NHWC (VGG19 feature extractor of the perceptual loss, reference utils/util.py:49-94)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        ctx.save_for_backward(x)
        return ops.maxpool2_fwd(x)

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        return ops.maxpool2_bwd(x, _c(dout))


def maxpool2(x):
    return MaxPool2Fn.apply(x)


class RoiAlignFn(torch.autograd.Function):
    """torchvision.ops.RoIAlign((8,8), scale, 0) on NHWC features (rcnn_discriminator_app.py:139,143)."""

    @staticmethod
    def forward(ctx, feat, rois, scale):
        feat = _c(feat)
        rois = _c(rois.float())
        ctx.save_for_backward(rois)
        ctx.meta = (scale, tuple(feat.shape))
        return ops.roi_align_fwd(feat, rois, scale)

    @staticmethod
    def backward(ctx, dout):
        (rois,) = ctx.saved_tensors
        scale, shape = ctx.meta
        return ops.roi_align_bwd(_c(dout), rois, scale, shape), None, None


def roi_align(feat, rois, scale):
    return RoiAlignFn.apply(feat, rois, scale)


class RoiAlign2Fn(torch.autograd.Function):
    """The object path's two RoIAligns + concatenation (rcnn_discriminator_app.py:137-146) as one launch: row k of the
    (K,8,8,C) stack samples the large-object map (level 0, scale 1/8), the small-object map (level 1, scale 1/4), or is
    zero (level 2: a dropped object in the fixed-shape form)."""

    @staticmethod
    def forward(ctx, feat_l, feat_s, rois, level, scale_l, scale_s):
        feat_l, feat_s = _c(feat_l), _c(feat_s)
        ctx.save_for_backward(rois, level)
        ctx.meta = (tuple(feat_l.shape), tuple(feat_s.shape), scale_l, scale_s)
        return ops.roi_align2_fwd(feat_l, scale_l, feat_s, scale_s, rois, level)

    @staticmethod
    def backward(ctx, dout):
        rois, level = ctx.saved_tensors
        shape_l, shape_s, scale_l, scale_s = ctx.meta
        dl, ds = ops.roi_align2_bwd(_c(dout), rois, level, shape_l, scale_l, shape_s, scale_s)
        return dl, ds, None, None, None, None


def roi_align2(feat_l, feat_s, rois, level, scale_l=1.0 / 8.0, scale_s=1.0 / 4.0):
    return RoiAlign2Fn.apply(feat_l, feat_s, rois, level, scale_l, scale_s)


class MasksToLayoutFn(torch.autograd.Function):
    """utils/bilinear.py:137-158 (grid_sample paste of per-object masks into the layout map)."""

    @staticmethod
    def forward(ctx, masks, bbox, size):
        masks, bbox = _c(masks), _c(bbox)
        ctx.save_for_backward(bbox)
        ctx.m = masks.shape[-1]
        return ops.masks_to_layout_fwd(bbox, masks, size)

    @staticmethod
    def backward(ctx, dout):
        (bbox,) = ctx.saved_tensors
        return ops.masks_to_layout_bwd(bbox, _c(dout), ctx.m), None, None


def masks_to_layout(masks, bbox, size):
    return MasksToLayoutFn.apply(masks, bbox, size)


class MaskResizeFn(torch.autograd.Function):
    """F.interpolate(mask, size=(h,w), mode='bilinear') of norm_module.py:176, optionally emitting the
    pixel-major (B,h,w,O) layout the ISLA kernels read."""

    @staticmethod
    def forward(ctx, mask, h, w, pixel_major):
        mask = _c(mask)
        ctx.meta = (mask.shape[2], mask.shape[3], pixel_major)
        return ops.mask_resize_fwd(mask, h, w, pixel_major)

    @staticmethod
    def backward(ctx, dout):
        hi, wi, pm = ctx.meta
        return ops.mask_resize_bwd(_c(dout), hi, wi, pm), None, None, None


def mask_resize(mask, h, w, pixel_major=True):
    return MaskResizeFn.apply(mask, h, w, pixel_major)


class StageMixFn(torch.autograd.Function):
    """resnet_generator_app_v2.py:466-470: stage_bbox = bilinear(bmask)*(1-a) + sigmoid(gather(stage_mask,y))*nearest(hard)*a."""

    @staticmethod
    def forward(ctx, stage, alpha, bmask, y, hard):
        stage, bmask = _c(stage), _c(bmask)
        alpha_flat = _c(alpha.reshape(-1))
        ctx.save_for_backward(stage, alpha_flat, bmask, y, hard)
        ctx.alpha_shape = alpha.shape
        return ops.stage_mix_fwd(stage, y, alpha_flat, bmask, hard)

    @staticmethod
    def backward(ctx, dout):
        stage, alpha_flat, bmask, y, hard = ctx.saved_tensors
        dstage, dalpha, dsoft = ops.stage_mix_bwd(stage, y, alpha_flat, bmask, hard, _c(dout))
        dbmask = ops.mask_resize_bwd(dsoft, bmask.shape[2], bmask.shape[3], False)
        return dstage, dalpha.view(ctx.alpha_shape), dbmask, None, None


def stage_mix(stage, alpha, bmask, y, hard):
    return StageMixFn.apply(stage, alpha, bmask, y, hard)


class ClassMixFn(torch.autograd.Function):
    """Gathered 1x1 mask head + stage-mask mixing (resnet_generator_app_v2.py:646/651 and :466-470) in one kernel: only the
    o class channels that the image's objects select are computed (the reference forms all 184 and gathers)."""

    @staticmethod
    def forward(ctx, t, weight, bias, alpha, bmask, y, hard):
        t, bmask = _c(t), _c(bmask)
        wc = _c(weight).view(weight.shape[0], -1)
        alpha_flat = _c(alpha.reshape(-1))
        sel, out = ops.class_mix_fwd(t, wc, _c(bias), y, alpha_flat, bmask, hard)
        ctx.save_for_backward(t, weight, alpha_flat, bmask, y, hard, sel)
        ctx.meta = (alpha.shape, bias is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        t, weight, alpha_flat, bmask, y, hard, sel = ctx.saved_tensors
        alpha_shape, has_bias = ctx.meta
        wc = _c(weight).view(weight.shape[0], -1)
        dt, dw, db, dalpha, dsoft = ops.class_mix_bwd(t, wc, y, alpha_flat, bmask, hard, sel, _c(dout), has_bias)
        dbmask = ops.mask_resize_bwd(dsoft, bmask.shape[2], bmask.shape[3], False)
        return dt, dw.view_as(weight), db, dalpha.view(alpha_shape), dbmask, None, None


def class_mix(t, conv, alpha, bmask, y, hard):
    """t: the mask head's features (B,h,w,100); conv: its final 1x1 Conv2d(100, 184) module (plain, not spectrally normalised)."""
    return ClassMixFn.apply(t, conv.weight, conv.bias, alpha, bmask, y, hard)


class BoxAttentionFn(torch.autograd.Function):
    """box_attention + relational embedding + WGs gate (resnet_generator_app_v2.py:17-120,172-192)."""

    @staticmethod
    def forward(ctx, q, k, v, bbox, y, wg_w, wg_b):
        q, k, v, bbox = _c(q), _c(k), _c(v), _c(bbox)
        wg = _c(wg_w.reshape(-1))
        out, p, glin = ops.box_attention_fwd(q, k, v, bbox, y, wg, _c(wg_b))
        ctx.save_for_backward(q, k, v, bbox, y, p, glin)
        ctx.wshape = wg_w.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, bbox, y, p, glin = ctx.saved_tensors
        dq, dk, dv, dwg, dbg = ops.box_attention_bwd(q, k, v, bbox, y, p, glin, _c(dout))
        return dq, dk, dv, None, None, dwg.view(ctx.wshape), dbg


def box_attention(q, k, v, bbox, y, wg_w, wg_b):
    return BoxAttentionFn.apply(q, k, v, bbox, y, wg_w, wg_b)


class ProjHeadFn(torch.autograd.Function):
    """The discriminator's projection heads (rcnn_discriminator_app.py:125-127,160-166):
        out = Linear_sn(sum_hw relu(feat)) [+ <Embedding_sn[y], sum_hw relu(feat)>]
    feat (N,h,w,C) NHWC.  One kernel forward (csrc/heads.cu), one backward + the spectral-norm gradient maps."""

    @staticmethod
    def forward(ctx, feat, w_orig, bias, sn_w, emb_orig, sn_e, y):
        feat = _c(feat)
        n, h, w_, c = feat.shape
        st_w = _sigma(w_orig, sn_w)
        st_e = _sigma(emb_orig, sn_e) if emb_orig is not None else None
        f3 = feat.view(n, h * w_, c)
        s, out = ops.head_fwd(f3, _c(w_orig), st_w.sigma, _c(bias), _c(emb_orig), st_e.sigma if st_e else None, y)
        ctx.save_for_backward(feat, s, w_orig, emb_orig, y, *st_w, *(st_e or (None, None, None)))
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        feat, s, w_orig, emb_orig, y, sg, u, v, sge, ue, ve = ctx.saved_tensors
        n, h, w_, c = feat.shape
        need = ctx.needs_input_grad
        dfeat, gw, gemb, dbias = ops.head_bwd(feat.view(n, h * w_, c), s, _c(dout).view(-1), _c(w_orig), sg, _c(emb_orig), sge, y,
                                              need_dfeat=need[0], need_gw=need[1] or need[4], need_db=ctx.has_bias and need[2])
        dw = demb = None
        if need[1]:
            dw = ops.sn_weight_grad(gw.view(1, 1, c), _c(w_orig), ops.SNState(sg, u, v)).view_as(w_orig)
        if emb_orig is not None and need[4]:
            demb = ops.sn_weight_grad(gemb.view(-1, 1, c), _c(emb_orig), ops.SNState(sge, ue, ve)).view_as(emb_orig)
        return (dfeat.view_as(feat) if dfeat is not None else None, dw, dbias if (ctx.has_bias and need[2]) else None, None,
                demb, None, None)


def proj_head(feat, linear, embedding=None, y=None):
    """feat (N,h,w,C) through a spectrally normalised nn.Linear(C, 1) (+ the class projection with a spectrally normalised
    nn.Embedding(num_classes, C))."""
    w, b, sn = _sn_of(linear)
    if sn is None:
        raise ValueError("proj_head expects a spectrally normalised linear layer")
    if embedding is None:
        return ProjHeadFn.apply(feat, w, b, sn, None, None, None)
    we, _, sne = _sn_of(embedding)
    return ProjHeadFn.apply(feat, w, b, sn, we, sne, y.contiguous())


class GramProjFn(torch.autograd.Function):
    """The appearance head (rcnn_discriminator_app.py:148-157) on x = app_conv(obj_feat) (K,h,w,C): ReLU, Gram matrix,
    concatenation with the class embedding, Linear(2C, 1), mean over rows -- as one kernel that never forms the Gram
    matrix (csrc/heads.cu gram_proj_*)."""

    @staticmethod
    def forward(ctx, x, w_orig, bias, sn_w, emb_orig, sn_e, y):
        x = _c(x)
        k, h, w_, c = x.shape
        if w_orig.shape != (1, 2 * c):
            raise ValueError(f"appearance projection must be (1, {2 * c}), got {tuple(w_orig.shape)}")
        st_w, st_e = _sigma(w_orig, sn_w), _sigma(emb_orig, sn_e)
        colsum, proj, out = ops.gram_proj_fwd(x.view(k, h * w_, c), _c(w_orig), st_w.sigma, _c(bias), _c(emb_orig), st_e.sigma, y)
        ctx.save_for_backward(x, colsum, proj, w_orig, emb_orig, y, *st_w, *st_e)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, colsum, proj, w_orig, emb_orig, y, sg, u, v, sge, ue, ve = ctx.saved_tensors
        k, h, w_, c = x.shape
        dx, gw, gemb, dbias = ops.gram_proj_bwd(x.view(k, h * w_, c), colsum, proj, _c(dout).view(-1), _c(w_orig), sg,
                                                _c(emb_orig), sge, y)
        dw = demb = None
        if ctx.needs_input_grad[1]:
            dw = ops.sn_weight_grad(gw.view(1, 1, 2 * c), _c(w_orig), ops.SNState(sg, u, v)).view_as(w_orig)
        if ctx.needs_input_grad[4]:
            demb = ops.sn_weight_grad(gemb.view(-1, 1, c), _c(emb_orig), ops.SNState(sge, ue, ve)).view_as(emb_orig)
        return (dx.view_as(x) if ctx.needs_input_grad[0] else None, dw, dbias if (ctx.has_bias and ctx.needs_input_grad[2]) else None,
                None, demb, None, None)


def gram_proj(x, linear, embedding, y):
    w, b, sn = _sn_of(linear)
    we, _, sne = _sn_of(embedding)
    if sn is None or sne is None:
        raise ValueError("gram_proj expects spectrally normalised app / l_y_app modules")
    return GramProjFn.apply(x, w, b, sn, we, sne, y.contiguous())
