"""Tensor-level wrappers over the C ABI (include/l2i.h): allocate outputs with torch, pass raw
pointers + the current stream to libl2i.so.  No arithmetic happens here.

Conventions: activations are contiguous fp32 tensors of shape (N, H, W, C) ("NHWC"); a `Pair`
is the (hi, lo) bf16 split of such a tensor with the channel count padded to a multiple of 8,
the operand format of the tensor-core convolutions (csrc/conv_tc.cu).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from ._lib import call


def pad8(c: int) -> int:
    """Channel count of a tensor-core operand pair: a multiple of 8 (TMA's 16-byte stride rule).  Tiny channel
    counts (RGB images, 1-channel mask logits) are padded to a full 64-wide K chunk: a TMA box that is mostly
    out of bounds in the channel dimension runs ~3x slower than the same box over zero-filled memory
    (measured on D.block1.conv1: 970 us -> 330 us), and the padding costs < 0.3 GB of the 180 GB HBM."""
    if c <= 8:
        return 64
    c8 = (c + 7) // 8 * 8
    tail = c8 % 64
    return c8 + (64 - tail) if 0 < tail <= 40 else c8      # e.g. 100 -> 128, 528 -> 576, 184 -> 184


class Pair(NamedTuple):
    hi: torch.Tensor   # (N, H, W, Cpad) bf16
    lo: torch.Tensor
    C: int             # real channel count

    @property
    def cpad(self) -> int:
        return self.hi.shape[-1]


class WeightPair(NamedTuple):
    f_hi: torch.Tensor            # (Cout, taps, CinPad) bf16  forward operand
    f_lo: torch.Tensor
    d_hi: Optional[torch.Tensor]  # (Cin, taps, CoutPad) bf16  data-gradient operand (flipped, transposed)
    d_lo: Optional[torch.Tensor]
    cout: int
    cin: int
    taps: int


def _chk(x: torch.Tensor, dtype=torch.float32):
    if not x.is_cuda:
        raise RuntimeError("layout2img_b200 ops run on CUDA tensors only (no CPU fallback)")
    if x.dtype != dtype or not x.is_contiguous():
        raise ValueError(f"expected contiguous {dtype} tensor, got {x.dtype} contiguous={x.is_contiguous()}")
    return x


def conv_weight_prep(w: torch.Tensor, sigma: Optional[torch.Tensor] = None, need_dgrad: bool = True) -> WeightPair:
    """w (Cout, Cin, kh, kw) fp32 in torch layout -> tensor-core operand pairs."""
    _chk(w)
    cout, cin, kh, kw = w.shape
    taps = kh * kw
    cin_pad, cout_pad = pad8(cin), pad8(cout)
    f = torch.empty((2, cout, taps, cin_pad), dtype=torch.bfloat16, device=w.device)
    d = torch.empty((2, cin, taps, cout_pad), dtype=torch.bfloat16, device=w.device) if need_dgrad else None
    call("l2i_conv_weight_prep", w, sigma, cout, cin, taps, f[0], f[1], cin_pad,
         d[0] if need_dgrad else None, d[1] if need_dgrad else None, cout_pad)
    return WeightPair(f[0], f[1], d[0] if need_dgrad else None, d[1] if need_dgrad else None, cout, cin, taps)


class SNState(NamedTuple):
    """What the backward of a spectrally normalised weight needs: sigma (1,), the u / v it was computed with."""
    sigma: torch.Tensor
    u: torch.Tensor
    v: torch.Tensor


def sn_sigma(w_orig: torch.Tensor, u: torch.Tensor, v: torch.Tensor, training: bool, eps: float) -> SNState:
    """torch.nn.utils.spectral_norm's pre-forward step: in training one power iteration updates u, v IN PLACE;
    returns sigma = u . (W v) and copies of the vectors used."""
    _chk(w_orig); _chk(u); _chk(v)
    r = w_orig.shape[0]
    cc = w_orig.numel() // r
    if u.numel() != r or v.numel() != cc:
        raise ValueError(f"spectral norm vectors ({u.numel()}, {v.numel()}) do not match weight ({r}, {cc})")
    out = torch.empty(1 + 2 * (r + cc), dtype=torch.float32, device=w_orig.device)
    sigma, u_used, v_used, work = out[:1], out[1:1 + r], out[1 + r:1 + r + cc], out[1 + r + cc:]
    call("l2i_sn_sigma", w_orig, r, cc, u, v, int(training), float(eps), u_used, v_used, sigma, work)
    return SNState(sigma, u_used, v_used)


def sn_weight_grad(g_ours: torch.Tensor, w_orig: torch.Tensor, st: SNState) -> torch.Tensor:
    """Tensor-core weight gradient (Cout, taps, Cin) -> gradient of weight_orig (Cout, Cin, kh, kw)."""
    cout, taps, cin = g_ours.shape
    dw = torch.empty_like(w_orig)
    scratch = torch.empty(1, dtype=torch.float32, device=w_orig.device)
    call("l2i_sn_weight_grad", g_ours, w_orig, st.u, st.v, st.sigma, cout, cin, taps, dw, scratch)
    return dw


def act_split(x: torch.Tensor, relu: bool = False, up2: bool = False) -> Pair:
    """x (N,H,W,C) fp32 -> Pair at (N, H<<up2, W<<up2, pad8(C)), optional ReLU first."""
    _chk(x)
    n, h, w, c = x.shape
    s = 2 if up2 else 1
    out = torch.empty((2, n, h * s, w * s, pad8(c)), dtype=torch.bfloat16, device=x.device)
    call("l2i_act_split", x, n, h, w, c, int(relu), int(up2), out[0], out[1], pad8(c))
    return Pair(out[0], out[1], c)


def act_split2(x: torch.Tensor, relu_a: bool, b_mode: int, b_scale: Optional[float] = None, want_a: bool = True):
    """x (N,H,W,C) -> (pair a = relu_a ? relu(x) : x, pair b = None | x | avgpool2(x))  [b_mode 0 | 1 | 2];
    b_scale overrides b's factor (default 1 for b_mode 1, 0.25 = average for b_mode 2; 1.0 there = 2x2 sum)."""
    if b_scale is None:
        b_scale = 0.25 if b_mode == 2 else 1.0
    _chk(x)
    n, h, w, c = x.shape
    cp = pad8(c)
    a = torch.empty((2, n, h, w, cp), dtype=torch.bfloat16, device=x.device) if want_a else None
    b = None
    if b_mode == 1:
        b = torch.empty((2, n, h, w, cp), dtype=torch.bfloat16, device=x.device)
    elif b_mode == 2:
        b = torch.empty((2, n, h // 2, w // 2, cp), dtype=torch.bfloat16, device=x.device)
    call("l2i_act_split2", x, n, h, w, c, int(relu_a), a[0] if want_a else None, a[1] if want_a else None, int(b_mode),
         float(b_scale), b[0] if b is not None else None, b[1] if b is not None else None, cp)
    return (Pair(a[0], a[1], c) if want_a else None), (Pair(b[0], b[1], c) if b is not None else None)


def im2col3(x: torch.Tensor, sign: int = 1, want_colsum: bool = False):
    """x (N,H,W,C<=4) fp32 -> Pair over 9C channels (order c*9+tap), pair[p] = x[p + sign*d(tap)]; optional colsum (C,)."""
    _chk(x)
    n, h, w, c = x.shape
    cp = pad8(9 * c)
    buf = torch.empty((2, n, h, w, cp), dtype=torch.bfloat16, device=x.device)
    colsum = torch.empty((c,), dtype=torch.float32, device=x.device) if want_colsum else None
    call("l2i_im2col3_pair", x, n, h, w, c, int(sign), buf[0], buf[1], cp, colsum)
    return Pair(buf[0], buf[1], 9 * c), colsum


def col2im3(col: torch.Tensor, c: int, sign: int = 1, bias=None, residual=None, res_up2: bool = False, res_scale: float = 1.0):
    """col (N,H,W,ldc>=9c) fp32 -> (N,H,W,c): out[q] = sum_tap col[q - sign*d(tap), c*9+tap] + bias + res_scale*residual."""
    _chk(col)
    n, h, w, ldc = col.shape
    out = torch.empty((n, h, w, c), dtype=torch.float32, device=col.device)
    if residual is not None:
        _chk(residual)
        exp = (n, h // 2, w // 2, c) if res_up2 else (n, h, w, c)
        if tuple(residual.shape) != exp:
            raise ValueError(f"col2im3: residual shape {tuple(residual.shape)} != {exp}")
    call("l2i_col2im3", col, ldc, n, h, w, c, int(sign), bias, residual, int(res_up2), float(res_scale), out)
    return out


def grad_split(g: torch.Tensor, want_lo: bool = True, up: bool = False, up_scale: float = 0.25):
    """g (N,H,W,C) fp32 -> (pair of g | None, pair of up_scale * nearest_x2(g) | None, colsum (C,))."""
    _chk(g)
    n, h, w, c = g.shape
    cp = pad8(c)
    lo = torch.empty((2, n, h, w, cp), dtype=torch.bfloat16, device=g.device) if want_lo else None
    hi = torch.empty((2, n, 2 * h, 2 * w, cp), dtype=torch.bfloat16, device=g.device) if up else None
    colsum = torch.empty((c,), dtype=torch.float32, device=g.device)
    call("l2i_grad_split", g, n, h, w, c, lo[0] if want_lo else None, lo[1] if want_lo else None, float(up_scale),
         hi[0] if up else None, hi[1] if up else None, colsum, cp)
    return (Pair(lo[0], lo[1], c) if want_lo else None), (Pair(hi[0], hi[1], c) if up else None), colsum


def pair_colsum(p: Pair) -> torch.Tensor:
    colsum = torch.empty((p.C,), dtype=torch.float32, device=p.hi.device)
    call("l2i_pair_colsum", p.hi, p.lo, p.hi.numel() // p.cpad, p.C, p.cpad, colsum)
    return colsum


def conv2d_fwd(x: Pair, w_hi: torch.Tensor, w_lo: torch.Tensor, cout: int, taps: int,
               bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
               res_up2: bool = False, res_scale: float = 1.0, out_scale: float = 1.0,
               mask_hi: Optional[torch.Tensor] = None, pool: int = 0, want_f32: bool = True,
               want_pair: bool = False, relu_pair: bool = False):
    """pool((conv(x, w) + bias) * out_scale * [mask_hi > 0]) + res_scale * residual
    -> (fp32 NHWC or None, Pair or None); see include/l2i.h."""
    n, h, w_, cin_pad = x.hi.shape
    # (cout, taps, cin_pad) tensors from conv_weight_prep, or flat slices of the grouped preparation's buffer
    if w_hi.shape != (cout, taps, cin_pad) and not (w_hi.dim() == 1 and w_hi.numel() >= cout * taps * cin_pad):
        raise ValueError(f"weight operand {tuple(w_hi.shape)} does not match ({cout},{taps},{cin_pad})")
    ho, wo = (h // 2, w_ // 2) if pool else (h, w_)
    out = torch.empty((n, ho, wo, cout), dtype=torch.float32, device=x.hi.device) if want_f32 else None
    pair = None
    if want_pair:
        buf = torch.empty((2, n, ho, wo, pad8(cout)), dtype=torch.bfloat16, device=x.hi.device)
        pair = Pair(buf[0], buf[1], cout)
    if residual is not None:
        _chk(residual)
        exp = (n, ho // 2, wo // 2, cout) if res_up2 else (n, ho, wo, cout)
        if tuple(residual.shape) != exp:
            raise ValueError(f"residual shape {tuple(residual.shape)} != {exp}")
    mask_cpad = 0
    if mask_hi is not None:
        if mask_hi.dtype != torch.bfloat16 or tuple(mask_hi.shape[:3]) != (n, h, w_) or not mask_hi.is_contiguous():
            raise ValueError("mask_hi must be a contiguous bf16 (N,H,W,Cpad) tensor at the conv's resolution")
        mask_cpad = mask_hi.shape[3]
    call("l2i_conv2d_fwd", n, h, w_, cin_pad, cout, taps, x.hi, x.lo, w_hi, w_lo, bias, residual, int(res_up2),
         float(res_scale), float(out_scale), mask_hi, mask_cpad, int(pool), out, pair.hi if pair else None,
         pair.lo if pair else None, pad8(cout), int(relu_pair))
    return out, pair


def conv2d_wgrad(dy: Pair, x: Pair, taps: int) -> torch.Tensor:
    """dW (Cout, taps, Cin) fp32 from the output-gradient pair and the saved input pair."""
    n, h, w_, cout_pad = dy.hi.shape
    if x.hi.shape[:3] != dy.hi.shape[:3]:
        raise ValueError("wgrad: dy and x spatial shapes differ")
    dw = torch.empty((dy.C, taps, x.C), dtype=torch.float32, device=dy.hi.device)
    call("l2i_conv2d_wgrad", n, h, w_, x.C, x.cpad, dy.C, cout_pad, taps, dy.hi, dy.lo, x.hi, x.lo, dw)
    return dw


# --------------------------------------------------------------------------------------------
# batch-norm statistics / ISLA
# --------------------------------------------------------------------------------------------
# Cross-rank batch statistics (the reference's multi-GPU SynchronizedBatchNorm2d, sync_batchnorm/batchnorm.py:90-111:
# sum and sum-of-squares reduced over all replicas).  Off by default: every rank normalises with its own shard,
# which is what the reference does on one GPU with the per-GPU batch.  set_sync_bn(True) turns the batch-norm
# layers that the reference synchronises (the ISLA norms and the affine norms of the mask heads / RGB head) into
# global-batch norms: one all-reduce of 2*C doubles in the forward and one in the backward of each layer.
_SYNC_BN = {"group": None, "world": 1}


def set_sync_bn(enabled, group=None):
    import torch.distributed as dist
    if enabled and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        _SYNC_BN["group"], _SYNC_BN["world"] = (group if group is not None else dist.group.WORLD), dist.get_world_size(group)
    else:
        _SYNC_BN["group"], _SYNC_BN["world"] = None, 1


def _sync_sum(t: torch.Tensor):
    if _SYNC_BN["world"] > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_SYNC_BN["group"])


def bn_batch_stats(x: torch.Tensor, running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor],
                   eps: float, momentum: float, sync: bool = True) -> torch.Tensor:
    """x (..., C) fp32 -> mean_invstd (2, C); updates the running statistics in place (train mode).  With
    set_sync_bn (and sync=True) the sums are all-reduced first, so mean / variance are those of the global batch."""
    _chk(x)
    c = x.shape[-1]
    pixels = x.numel() // c
    sums = torch.empty((c, 2), dtype=torch.float64, device=x.device)
    call("l2i_bn_stats", x, pixels, c, sums)
    world = _SYNC_BN["world"] if sync else 1
    if world > 1:
        _sync_sum(sums)
    mi = torch.empty((2, c), dtype=torch.float32, device=x.device)
    call("l2i_bn_finalize", sums, float(pixels * world), c, float(eps), float(momentum), running_mean,
         running_var, mi)
    return mi


def bn_eval_stats(running_mean: torch.Tensor, running_var: torch.Tensor, eps: float) -> torch.Tensor:
    c = running_mean.numel()
    mi = torch.empty((2, c), dtype=torch.float32, device=running_mean.device)
    call("l2i_bn_eval_stats", running_mean, running_var, c, float(eps), mi)
    return mi


def isla_fwd(x, mean_invstd, mask_pm, gamma, beta, aff_w, aff_b, relu: bool, up2: bool,
             want_f32: bool = False, want_pair: bool = True, chan_scale=None, relu_f32: bool = False):
    """x (B,H,W,C); mask_pm (B,H,W,O) or None; gamma/beta (B,O,C) -> (fp32 out or None, Pair or None).  chan_scale (B,C):
    O == 0 only; relu_f32: the fp32 output is ReLU'd too."""
    _chk(x)
    b, h, w, c = x.shape
    o = 0 if mask_pm is None else mask_pm.shape[-1]
    out = torch.empty_like(x) if want_f32 else None
    pair = None
    if want_pair:
        s = 2 if up2 else 1
        buf = torch.empty((2, b, h * s, w * s, pad8(c)), dtype=torch.bfloat16, device=x.device)
        pair = Pair(buf[0], buf[1], c)
    call("l2i_isla_fwd", x, mean_invstd, mask_pm, gamma, beta, aff_w, aff_b, chan_scale, b, h, w, c, o, out,
         pair.hi if pair else None, pair.lo if pair else None, pad8(c), int(bool(relu)) | (2 if relu_f32 else 0), int(up2))
    return out, pair


def isla_bwd(x, mean_invstd, mask_pm, gamma, beta, aff_w, aff_b, dout, relu: bool, up2: bool, train: bool,
             chan_scale=None, sync: bool = True):
    """-> dx (B,H,W,C), dmask_pm (B,H,W,O) | None, dgamma, dbeta (B,O,C) | None, csum (C,2) fp64."""
    _chk(x); _chk(dout)
    b, h, w, c = x.shape
    o = 0 if mask_pm is None else mask_pm.shape[-1]
    dx = torch.empty_like(x)
    csum = torch.empty((c, 2), dtype=torch.float64, device=x.device)
    dmask = torch.empty_like(mask_pm) if o else None
    dgamma = torch.empty_like(gamma) if o else None
    dbeta = torch.empty_like(beta) if o else None
    args = (x, mean_invstd, mask_pm, gamma, beta, aff_w, aff_b, chan_scale, dout, b, h, w, c, o, int(relu), int(up2),
            int(train), None, dmask, dgamma, dbeta, csum, dx)
    if train and sync and _SYNC_BN["world"] > 1:
        # global-batch norm: reduce (sum d xhat, sum d xhat * xhat) over the ranks between the two phases.  For the
        # affine form (O == 0) csum also carries the LOCAL (d bias, d weight); keep a copy for the caller.
        call("l2i_isla_bwd", *args, 1, 0.0)
        local = csum.clone() if o == 0 else None
        _sync_sum(csum)
        call("l2i_isla_bwd", *args, 2, float(b * h * w * _SYNC_BN["world"]))
        if local is not None:
            csum = local
    else:
        call("l2i_isla_bwd", *args, 0, 0.0)
    return dx, dmask, dgamma, dbeta, csum


# --------------------------------------------------------------------------------------------
# layout maps
# --------------------------------------------------------------------------------------------
def bbox_mask(bbox: torch.Tensor, H: int, W: int) -> torch.Tensor:
    _chk(bbox)
    b, o, _ = bbox.shape
    out = torch.empty((b, o, H, W), dtype=torch.float32, device=bbox.device)
    call("l2i_bbox_mask", bbox, b * o, H, W, out)
    return out


def masks_to_layout_fwd(bbox, masks, size: int):
    _chk(bbox); _chk(masks)
    b, o, m, m2 = masks.shape
    assert masks.shape == (b, o, m, m)          # reference utils/bilinear.py:149
    out = torch.empty((b, o, size, size), dtype=torch.float32, device=masks.device)
    call("l2i_masks_to_layout_fwd", bbox, masks, b * o, m, size, out)
    return out


def masks_to_layout_bwd(bbox, dout, m: int):
    _chk(dout)
    b, o, s, _ = dout.shape
    dm = torch.empty((b, o, m, m), dtype=torch.float32, device=dout.device)
    call("l2i_masks_to_layout_bwd", bbox, dout, b * o, m, s, dm)
    return dm


def mask_resize_fwd(mask, h: int, w: int, pixel_major: bool):
    _chk(mask)
    b, o, hi, wi = mask.shape
    out = torch.empty((b, h, w, o) if pixel_major else (b, o, h, w), dtype=torch.float32, device=mask.device)
    call("l2i_mask_resize_fwd", mask, b, o, hi, wi, h, w, int(pixel_major), out)
    return out


def mask_resize_bwd(dout, hi: int, wi: int, pixel_major: bool):
    _chk(dout)
    if pixel_major:
        b, h, w, o = dout.shape
    else:
        b, o, h, w = dout.shape
    din = torch.empty((b, o, hi, wi), dtype=torch.float32, device=dout.device)
    call("l2i_mask_resize_bwd", dout, b, o, hi, wi, h, w, int(pixel_major), din)
    return din


def stage_mix_fwd(stage, y, alpha, bmask, hard):
    _chk(stage); _chk(bmask); _chk(hard); _chk(alpha); _chk(y, torch.int64)
    b, h, w, nc = stage.shape
    o, s = bmask.shape[1], bmask.shape[2]
    out = torch.empty((b, o, h, w), dtype=torch.float32, device=stage.device)
    call("l2i_stage_mix_fwd", stage, y, alpha, bmask, hard, b, o, h, w, nc, s, out)
    return out


def stage_mix_bwd(stage, y, alpha, bmask, hard, dout):
    _chk(dout)
    b, h, w, nc = stage.shape
    o, s = bmask.shape[1], bmask.shape[2]
    dstage = torch.zeros_like(stage)
    dalpha = torch.zeros_like(alpha)
    dsoft = torch.empty_like(dout)
    call("l2i_stage_mix_bwd", stage, y, alpha, bmask, hard, dout, b, o, h, w, nc, s, dstage, dalpha, dsoft)
    return dstage, dalpha, dsoft


def class_mix_fwd(t, wc, bc, y, alpha, bmask, hard):
    """t (B,h,w,C), wc (NC,C), bc (NC,) | None, y (B,O) int64, alpha (NC,), bmask / hard (B,O,S,S) -> (sel, out) (B,O,h,w)."""
    _chk(t); _chk(wc); _chk(bmask); _chk(hard); _chk(alpha); _chk(y, torch.int64)
    b, h, w, c = t.shape
    o, s = bmask.shape[1], bmask.shape[2]
    nc = wc.shape[0]
    sel = torch.empty((b, o, h, w), dtype=torch.float32, device=t.device)
    out = torch.empty((b, o, h, w), dtype=torch.float32, device=t.device)
    call("l2i_class_mix_fwd", t, wc, bc, y, alpha, bmask, hard, b, o, h, w, c, nc, s, sel, out)
    return sel, out


def class_mix_bwd(t, wc, y, alpha, bmask, hard, sel, dout, need_db: bool):
    _chk(dout)
    b, h, w, c = t.shape
    o, s = bmask.shape[1], bmask.shape[2]
    nc = wc.shape[0]
    dev = t.device
    dt = torch.empty_like(t)
    dw = torch.empty((nc, c), dtype=torch.float32, device=dev)
    db = torch.empty((nc,), dtype=torch.float32, device=dev) if need_db else None
    dalpha = torch.empty((nc,), dtype=torch.float32, device=dev)
    dsoft = torch.empty_like(dout)
    call("l2i_class_mix_bwd", t, wc, y, alpha, bmask, hard, sel, dout, b, o, h, w, c, nc, s, dt, dw, db, dalpha, dsoft)
    return dt, dw, db, dalpha, dsoft


def inorm_relu_fwd(x, up2: bool, eps: float = 1e-5):
    """InstanceNorm (no affine) -> ReLU -> [bilinear x2] of x (N,H,W,C) -> (Pair at the output resolution, stats)."""
    _chk(x)
    n, h, w, c = x.shape
    s = 2 if up2 else 1
    stats = torch.empty((n, c, 2), dtype=torch.float32, device=x.device)
    buf = torch.empty((2, n, h * s, w * s, pad8(c)), dtype=torch.bfloat16, device=x.device)
    call("l2i_inorm_relu_fwd", x, n, h, w, c, int(up2), float(eps), stats, buf[0], buf[1], pad8(c))
    return Pair(buf[0], buf[1], c), stats


def inorm_relu_bwd(x, stats, da, up2: bool):
    _chk(x); _chk(da)
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    call("l2i_inorm_relu_bwd", x, stats, da, n, h, w, c, int(up2), dx)
    return dx


# --------------------------------------------------------------------------------------------
# pyramid pooling (PSPModule)
# --------------------------------------------------------------------------------------------
PSP_CELLS = 50          # sizes (1, 2, 3, 6)


def psp_pool_fwd(x):
    _chk(x)
    b, h, w, c = x.shape
    pooled = torch.empty((b, PSP_CELLS, c), dtype=torch.float32, device=x.device)
    call("l2i_psp_pool_fwd", x, b, h, w, c, pooled)
    return pooled


def psp_pool_bwd(dpooled, shape):
    _chk(dpooled)
    b, h, w, c = shape
    dx = torch.empty(shape, dtype=torch.float32, device=dpooled.device)
    call("l2i_psp_pool_bwd", dpooled, None, 0, 0, b, h, w, c, dx)
    return dx


def psp_concat_fwd(feats, priors) -> Pair:
    _chk(feats); _chk(priors)
    b, h, w, c = feats.shape
    cp = priors.shape[2]
    if priors.shape != (b, PSP_CELLS, cp):
        raise ValueError(f"priors must be (B, {PSP_CELLS}, CP), got {tuple(priors.shape)}")
    ctot = 4 * cp + c
    buf = torch.empty((2, b, h, w, pad8(ctot)), dtype=torch.bfloat16, device=feats.device)
    call("l2i_psp_concat_fwd", feats, priors, b, h, w, c, cp, buf[0], buf[1], pad8(ctot))
    return Pair(buf[0], buf[1], ctot)


def psp_concat_bwd(dcat, cp: int):
    _chk(dcat)
    b, h, w, cs = dcat.shape
    dpriors = torch.empty((b, PSP_CELLS, cp), dtype=torch.float32, device=dcat.device)
    call("l2i_psp_concat_bwd", dcat, b, h, w, cp, cs, dpriors)
    return dpriors


# --------------------------------------------------------------------------------------------
# ROIAlign / pooling
# --------------------------------------------------------------------------------------------
def roi_align_fwd(feat, rois, scale: float, P: int = 8):
    _chk(feat)
    n, h, w, c = feat.shape
    k = rois.shape[0]
    out = torch.empty((k, P, P, c), dtype=torch.float32, device=feat.device)
    if k:
        _chk(rois)
        call("l2i_roi_align_fwd", feat, rois, k, n, h, w, c, P, float(scale), out)
    return out


def roi_align_bwd(dout, rois, scale: float, shape, P: int = 8):
    n, h, w, c = shape
    k = rois.shape[0]
    dfeat = torch.empty(shape, dtype=torch.float32, device=dout.device)
    call("l2i_roi_align_bwd", _chk(dout) if k else None, rois if k else None, k, n, h, w, c, P, float(scale), dfeat)
    return dfeat


def roi_prepare(bbox: torch.Tensor, label: torch.Tensor, img_size: float, small_thresh: float = 64.0):
    """bbox (B,O,4) xywh fp32, label (B*O,) int64 on the device -> (rois (B*O,5), y_sorted (B*O,), level (B*O,) int32,
    perm (B*O,) int32, counts (2,) int32 = (n_large, n_small)); rows ordered [large, small, dropped].  No host sync."""
    _chk(bbox); _chk(label, torch.int64)
    b, o, _ = bbox.shape
    k = b * o
    dev = bbox.device
    rois = torch.empty((k, 5), dtype=torch.float32, device=dev)
    y_sorted = torch.empty((k,), dtype=torch.int64, device=dev)
    meta = torch.empty((2 * k + 2,), dtype=torch.int32, device=dev)
    level, perm, counts = meta[:k], meta[k:2 * k], meta[2 * k:]
    call("l2i_roi_prepare", bbox, label, b, o, float(img_size), float(small_thresh), rois, y_sorted, level, perm, counts)
    return rois, y_sorted, level, perm, counts


def roi_align2_fwd(feat_l, scale_l: float, feat_s, scale_s: float, rois, level, P: int = 8):
    _chk(feat_l); _chk(feat_s)
    n, hl, wl, c = feat_l.shape
    _, hs, ws, _ = feat_s.shape
    k = rois.shape[0]
    out = torch.empty((k, P, P, c), dtype=torch.float32, device=feat_l.device)
    call("l2i_roi_align2_fwd", feat_l, hl, wl, float(scale_l), feat_s, hs, ws, float(scale_s), rois, level, k, n, c, P, out)
    return out


def roi_align2_bwd(dout, rois, level, shape_l, scale_l: float, shape_s, scale_s: float, P: int = 8):
    n, hl, wl, c = shape_l
    _, hs, ws, _ = shape_s
    k = rois.shape[0]
    dl = torch.empty(shape_l, dtype=torch.float32, device=dout.device)
    ds = torch.empty(shape_s, dtype=torch.float32, device=dout.device)
    call("l2i_roi_align2_bwd", _chk(dout) if k else None, rois if k else None, level if k else None, k, n, c, P, hl, wl,
         float(scale_l), dl, hs, ws, float(scale_s), ds)
    return dl, ds


def avgpool2_fwd(x):
    _chk(x)
    n, h, w, c = x.shape
    out = torch.empty((n, h // 2, w // 2, c), dtype=torch.float32, device=x.device)
    call("l2i_avgpool2_fwd", x, n, h, w, c, out)
    return out


def avgpool2_bwd(dout):
    _chk(dout)
    n, ho, wo, c = dout.shape
    dx = torch.empty((n, ho * 2, wo * 2, c), dtype=torch.float32, device=dout.device)
    call("l2i_avgpool2_bwd", dout, n, ho * 2, wo * 2, c, dx)
    return dx


def colsum(x2: torch.Tensor) -> torch.Tensor:
    """(M, N) fp32 -> (N,) column sums."""
    _chk(x2)
    m, n = x2.shape
    out = torch.empty((n,), dtype=torch.float32, device=x2.device)
    call("l2i_colsum", x2, m, n, out)
    return out


def add_layernorm_fwd(a, b, w, bias, eps: float):
    _chk(a)
    d = a.shape[-1]
    rows = a.numel() // d
    y = torch.empty_like(a)
    stats = torch.empty((rows, 2), dtype=torch.float32, device=a.device)
    call("l2i_add_layernorm_fwd", a, b, w, bias, rows, d, float(eps), y, stats)
    return y, stats


def add_layernorm_bwd(a, b, w, stats, dy):
    _chk(dy)
    d = a.shape[-1]
    rows = a.numel() // d
    ds = torch.empty_like(a)
    dw = torch.empty((d,), dtype=torch.float32, device=a.device)
    db = torch.empty((d,), dtype=torch.float32, device=a.device)
    call("l2i_add_layernorm_bwd", a, b, w, stats, dy, rows, d, ds, dw, db)
    return ds, dw, db


def maxpool2_fwd(x):
    _chk(x)
    n, h, w, c = x.shape
    out = torch.empty((n, h // 2, w // 2, c), dtype=torch.float32, device=x.device)
    call("l2i_maxpool2_fwd", x, n, h, w, c, out)
    return out


def maxpool2_bwd(x, dout):
    _chk(dout)
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    call("l2i_maxpool2_bwd", x, dout, n, h, w, c, dx)
    return dx


# --------------------------------------------------------------------------------------------
# discriminator heads
# --------------------------------------------------------------------------------------------
def head_fwd(feat, w, sigma_w, bias, emb, sigma_e, y):
    """feat (N,P,C) -> (s (N,C), out (N,1)); see include/l2i.h."""
    _chk(feat); _chk(w)
    n, p, c = feat.shape
    s = torch.empty((n, c), dtype=torch.float32, device=feat.device)
    out = torch.empty((n, 1), dtype=torch.float32, device=feat.device)
    call("l2i_head_fwd", feat, n, p, c, w, sigma_w, bias, emb, sigma_e, y, s, out)
    return s, out


def head_bwd(feat, s, dout, w, sigma_w, emb, sigma_e, y, need_dfeat=True, need_gw=True, need_db=True):
    n, p, c = feat.shape
    dev = feat.device
    dfeat = torch.empty_like(feat) if need_dfeat else None
    gw = torch.empty((c,), dtype=torch.float32, device=dev) if need_gw else None
    gemb = torch.empty((emb.shape[0], c), dtype=torch.float32, device=dev) if (emb is not None and need_gw) else None
    dbias = torch.empty((1,), dtype=torch.float32, device=dev) if need_db else None
    call("l2i_head_bwd", feat, s, _chk(dout), n, p, c, w, sigma_w, emb, sigma_e, y, emb.shape[0] if emb is not None else 0,
         dfeat, gw, gemb, dbias)
    return dfeat, gw, gemb, dbias


def gram_proj_fwd(x, w, sigma_w, bias, emb, sigma_e, y):
    """x (K,P,C) -> (colsum (K,P), proj (K,P), out (K,1))."""
    _chk(x); _chk(w); _chk(emb)
    k, p, c = x.shape
    colsum = torch.empty((k, p), dtype=torch.float32, device=x.device)
    proj = torch.empty((k, p), dtype=torch.float32, device=x.device)
    out = torch.empty((k, 1), dtype=torch.float32, device=x.device)
    call("l2i_gram_proj_fwd", x, k, p, c, w, sigma_w, bias, emb, sigma_e, y, colsum, proj, out)
    return colsum, proj, out


def gram_proj_bwd(x, colsum, proj, dout, w, sigma_w, emb, sigma_e, y):
    k, p, c = x.shape
    dev = x.device
    dx = torch.empty_like(x)
    gw = torch.empty((2 * c,), dtype=torch.float32, device=dev)
    gemb = torch.empty((emb.shape[0], c), dtype=torch.float32, device=dev)
    dbias = torch.empty((1,), dtype=torch.float32, device=dev)
    call("l2i_gram_proj_bwd", x, colsum, proj, _chk(dout), k, p, c, w, sigma_w, emb, sigma_e, y, emb.shape[0], dx, gw, gemb, dbias)
    return dx, gw, gemb, dbias


# --------------------------------------------------------------------------------------------
# object-context attention
# --------------------------------------------------------------------------------------------
def box_attention_fwd(q, k, v, bbox, y, wg, bg):
    for t in (q, k, v, bbox, wg, bg):
        _chk(t)
    _chk(y, torch.int64)
    b, o, d = q.shape
    out = torch.empty_like(q)
    p = torch.empty((b, o, o), dtype=torch.float32, device=q.device)
    glin = torch.empty((b, o, o), dtype=torch.float32, device=q.device)
    call("l2i_box_attention_fwd", q, k, v, bbox, y, wg, bg, b, o, d, out, p, glin)
    return out, p, glin


def box_attention_bwd(q, k, v, bbox, y, p, glin, dout):
    _chk(dout)
    b, o, d = q.shape
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    dwg = torch.empty((64,), dtype=torch.float32, device=q.device)
    dbg = torch.empty((1,), dtype=torch.float32, device=q.device)
    call("l2i_box_attention_bwd", q, k, v, bbox, y, p, glin, dout, b, o, d, dq, dk, dv, dwg, dbg)
    return dq, dk, dv, dwg, dbg
