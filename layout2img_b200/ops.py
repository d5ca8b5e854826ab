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
    return (c + 7) // 8 * 8


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


def act_split(x: torch.Tensor, relu: bool = False, up2: bool = False) -> Pair:
    """x (N,H,W,C) fp32 -> Pair at (N, H<<up2, W<<up2, pad8(C)), optional ReLU first."""
    _chk(x)
    n, h, w, c = x.shape
    s = 2 if up2 else 1
    out = torch.empty((2, n, h * s, w * s, pad8(c)), dtype=torch.bfloat16, device=x.device)
    call("l2i_act_split", x, n, h, w, c, int(relu), int(up2), out[0], out[1], pad8(c))
    return Pair(out[0], out[1], c)


def conv2d_fwd(x: Pair, w_hi: torch.Tensor, w_lo: torch.Tensor, cout: int, taps: int,
               bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
               res_up2: bool = False, out_scale: float = 1.0, want_f32: bool = True,
               want_pair: bool = False, relu_pair: bool = False):
    """(conv(x, w) + bias + residual) * out_scale -> (fp32 NHWC or None, Pair or None)."""
    n, h, w_, cin_pad = x.hi.shape
    if w_hi.shape != (cout, taps, cin_pad):
        raise ValueError(f"weight operand {tuple(w_hi.shape)} does not match ({cout},{taps},{cin_pad})")
    out = torch.empty((n, h, w_, cout), dtype=torch.float32, device=x.hi.device) if want_f32 else None
    pair = None
    if want_pair:
        buf = torch.empty((2, n, h, w_, pad8(cout)), dtype=torch.bfloat16, device=x.hi.device)
        if pad8(cout) != cout:
            buf.zero_()
        pair = Pair(buf[0], buf[1], cout)
    if residual is not None:
        _chk(residual)
        exp = (n, h // 2, w_ // 2, cout) if res_up2 else (n, h, w_, cout)
        if tuple(residual.shape) != exp:
            raise ValueError(f"residual shape {tuple(residual.shape)} != {exp}")
    call("l2i_conv2d_fwd", n, h, w_, cin_pad, cout, taps, x.hi, x.lo, w_hi, w_lo, bias, residual, int(res_up2),
         float(out_scale), out, pair.hi if pair else None, pair.lo if pair else None, pad8(cout), int(relu_pair))
    return out, pair


def conv2d_wgrad(dy: Pair, x: Pair, taps: int) -> torch.Tensor:
    """dW (Cout, taps, Cin) fp32 from the output-gradient pair and the saved input pair."""
    n, h, w_, cout_pad = dy.hi.shape
    if x.hi.shape[:3] != dy.hi.shape[:3]:
        raise ValueError("wgrad: dy and x spatial shapes differ")
    dw = torch.empty((dy.C, taps, x.C), dtype=torch.float32, device=dy.hi.device)
    call("l2i_conv2d_wgrad", n, h, w_, x.C, x.cpad, dy.C, cout_pad, taps, dy.hi, dy.lo, x.hi, x.lo, dw)
    return dw
