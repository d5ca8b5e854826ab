"""Network-level spectral normalisation + weight preparation: ONE C-ABI call (6 launches) per network forward does the
power iteration / sigma of every `nn.utils.spectral_norm` module and writes the tensor-core operand pairs of every
convolution weight, instead of ~4 launches per module (l2i_sn_prepare_group, csrc/specnorm.cu + csrc/prep.cu).

torch runs the power iteration in each module's pre-forward hook (call sites resnet_generator_app_v2.py:681-686,
rcnn_discriminator_app.py:10-15); it depends on (W, u, v) only, so doing all of them before the first layer is the same
arithmetic.  A module that is CALLED TWICE per forward (D.block_obj4, rcnn_discriminator_app.py:137,141) iterates twice
in the reference: its second call finds no prepared entry (entries are consumed once) and runs the per-module kernels.

Results live in two buffers allocated per call (so the autograd nodes of earlier calls keep theirs), handed to the
consumers through functional.take_prepared(weight).
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional

import numpy as np
import torch

from ._lib import call
from .ops import SNState, WeightPair, pad8

_ENTRY = np.dtype([("W", "<u8"), ("u", "<u8"), ("v", "<u8"), ("f32_off", "<i8"), ("bf_off", "<i8"), ("R", "<i4"),
                   ("Cc", "<i4"), ("eps", "<f4"), ("training", "<i4"), ("has_sn", "<i4"), ("cin", "<i4"), ("taps", "<i4"),
                   ("pad", "<i4")])
assert _ENTRY.itemsize == 72

# id(weight tensor) -> (weakref to the tensor, its version counter, SNState | None, WeightPair | None); filled by
# SNGroup.prepare, consumed once by the autograd nodes.  An entry is only honoured for the very tensor it was made
# from, unmodified since: ids are recycled once a network is freed, and an entry that nobody consumed (a module the
# forward did not reach) must not be picked up after an optimizer step.
PREPARED: Dict[int, tuple] = {}


def _valid(entry, weight: torch.Tensor) -> bool:
    return entry is not None and entry[0]() is weight and entry[1] == weight._version


def take_prepared(weight: torch.Tensor):
    entry = PREPARED.pop(id(weight), None)
    return (entry[2], entry[3]) if _valid(entry, weight) else None


def has_prepared(weight: torch.Tensor) -> bool:
    return _valid(PREPARED.get(id(weight)), weight)


def _pad4(n: int) -> int:
    return (n + 3) & ~3


def _round64(n: int) -> int:
    return (n + 63) & ~63


def _sn_hook(m):
    from torch.nn.utils.spectral_norm import SpectralNorm
    for hook in m._forward_pre_hooks.values():
        if isinstance(hook, SpectralNorm):
            if hook.n_power_iterations != 1 or hook.dim != 0:
                raise ValueError("layout2img_b200 implements spectral_norm(n_power_iterations=1, dim=0)")
            return hook
    return None


class SNGroup:
    def __init__(self, net: torch.nn.Module):
        from .model.layers import Conv2d
        self.mods: List[dict] = []
        for m in net.modules():
            hook = _sn_hook(m)
            is_conv = isinstance(m, Conv2d)
            is_lin = isinstance(m, torch.nn.Linear)
            if hook is None and not is_conv and not is_lin:
                continue
            w = m.weight_orig if hook is not None else m.weight
            cin, taps, pairs = 0, 0, is_conv
            if is_lin and w.shape[0] > 1:        # linear layers run as 1x1 convolutions over M one-pixel images (functional.LinearFn)
                cin, taps, pairs = w.shape[1], 1, True
            if is_conv:
                cin, taps = w.shape[1], w.shape[2] * w.shape[3]
                if taps == 9 and cin <= 4:
                    cin, taps = 9 * cin, 1       # small-input 3x3: prepared as the 1x1 weight of its im2col form (same memory)
                elif taps == 9 and w.shape[0] <= 4:
                    pairs = False                # small-output 3x3: its consumer prepares the permuted (Cout*9, Cin) weight
                elif getattr(m, "_l2i_gathered_head", False):
                    pairs = False                # the mask heads' 1x1 convolution is evaluated gathered (functional.class_mix)
            self.mods.append({"m": m, "hook": hook, "conv": pairs, "R": w.shape[0], "Cc": w.numel() // w.shape[0],
                              "cin": cin if pairs else 0, "taps": taps if pairs else 0})
        # static layout of the two per-call buffers
        f_off = b_off = 0
        self.f32_sizes, self.bf_sizes = [], []
        for d in self.mods:
            d["f32_off"] = f_off
            sizes = [1, 3, _pad4(d["R"]), _pad4(d["Cc"]), _pad4(d["Cc"]), _pad4(d["R"])]
            self.f32_sizes += sizes
            f_off += sum(sizes)
            if d["conv"]:
                d["bf_off"] = b_off
                nf = _round64(d["R"] * d["taps"] * pad8(d["cin"]))
                nd = _round64(d["cin"] * d["taps"] * pad8(d["R"]))
                self.bf_sizes += [nf, nf, nd, nd]
                b_off += 2 * nf + 2 * nd
            else:
                d["bf_off"] = -1
        self.f32_floats, self.bf_elems = f_off, b_off
        self._table = None
        self._sig = None
        self._keys: List[int] = []

    # ------------------------------------------------------------------------------------------
    def _weights(self, d):
        m = d["m"]
        if d["hook"] is not None:
            return m.weight_orig, m.weight_u, m.weight_v
        return m.weight, None, None

    def _signature(self):
        sig = []
        for d in self.mods:
            w, u, v = self._weights(d)
            sig.append((w.data_ptr(), u.data_ptr() if u is not None else 0, v.data_ptr() if v is not None else 0, d["m"].training))
        return sig

    def _build(self, dev):
        n = len(self.mods)
        tab = np.zeros(n, dtype=_ENTRY)
        wt, wv, p9, p1 = [], [], [], []
        wt_smem, max_cc = 1, 1
        for i, d in enumerate(self.mods):
            w, u, v = self._weights(d)
            for t in (w, u, v):
                if t is not None and (t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous()):
                    raise ValueError("layout2img_b200 needs contiguous fp32 CUDA parameters (no CPU fallback)")
            has_sn = d["hook"] is not None
            tab[i] = (w.data_ptr(), u.data_ptr() if has_sn else 0, v.data_ptr() if has_sn else 0, d["f32_off"], d["bf_off"],
                      d["R"], d["Cc"], float(d["hook"].eps) if has_sn else 0.0, int(d["m"].training), int(has_sn), d["cin"],
                      d["taps"], 0)
            R, Cc = d["R"], d["Cc"]
            if has_sn:
                if Cc > 40000:
                    raise ValueError(f"spectral norm: {Cc} columns exceed the kernel's shared-memory vector")
                max_cc = max(max_cc, Cc)
                col_blocks = (Cc + 255) // 256
                splits = max(1, min((296 + col_blocks - 1) // col_blocks, (R + 15) // 16))
                rps = (R + splits - 1) // splits
                wt_smem = max(wt_smem, rps)
                for r0 in range(0, R, rps):
                    wt += [(i, cb, r0, min(R, r0 + rps)) for cb in range(col_blocks)]
                wv += [(i, rb) for rb in range((R + 7) // 8)]
            if d["conv"]:
                # (linear weights (R, K) have the memory layout of a (R, K, 1, 1) convolution weight)
                gx = (pad8(d["cin"]) + 31) // 32
                gy = (pad8(R) + 31) // 32
                (p9 if d["taps"] == 9 else p1).extend((i, tx, ty, 0) for ty in range(gy) for tx in range(gx))
        mk = lambda rows, width: torch.tensor(rows if rows else [[0] * width], dtype=torch.int32, device=dev).contiguous()
        self._table = {
            "tab": torch.from_numpy(tab.view(np.uint8)).to(dev), "n": n, "wt": mk(wt, 4), "n_wt": len(wt), "wt_smem": wt_smem,
            "wv": mk(wv, 2), "n_wv": len(wv), "max_cc": max_cc, "p9": mk(p9, 4), "n9": len(p9), "p1": mk(p1, 4), "n1": len(p1),
        }

    # ------------------------------------------------------------------------------------------
    def prepare(self, need_dgrad: bool):
        """Run the grouped launch and publish every module's (SNState, WeightPair) for its consumer."""
        if not self.mods:
            return
        w0 = self._weights(self.mods[0])[0]
        dev = w0.device
        if not w0.is_cuda:
            raise RuntimeError("layout2img_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        sig = self._signature()
        if sig != self._sig:
            self._build(dev)
            self._sig = sig
        t = self._table
        f32 = torch.empty(self.f32_floats, dtype=torch.float32, device=dev)
        bf = torch.empty(max(self.bf_elems, 64), dtype=torch.bfloat16, device=dev)
        call("l2i_sn_prepare_group", t["tab"], t["n"], t["wt"], t["n_wt"], t["wt_smem"], t["wv"], t["n_wv"], t["max_cc"],
             t["p9"], t["n9"], t["p1"], t["n1"], f32, self.f32_floats, bf, int(need_dgrad))
        fv = torch.split_with_sizes(f32, self.f32_sizes)
        bv = torch.split_with_sizes(bf[:self.bf_elems], self.bf_sizes) if self.bf_elems else ()
        for k in self._keys:                   # entries of the previous call that nobody consumed
            PREPARED.pop(k, None)
        for k in [k for k, e in PREPARED.items() if e[0]() is None]:      # ... and of networks that no longer exist
            del PREPARED[k]
        self._keys = []
        bi = 0
        for i, d in enumerate(self.mods):
            st = SNState(fv[6 * i], fv[6 * i + 2], fv[6 * i + 3]) if d["hook"] is not None else None
            wp = None
            if d["conv"]:
                wp = WeightPair(bv[bi], bv[bi + 1], bv[bi + 2] if need_dgrad else None, bv[bi + 3] if need_dgrad else None,
                                d["R"], d["cin"], d["taps"])
                bi += 4
            w = self._weights(d)[0]
            PREPARED[id(w)] = (weakref.ref(w), w._version, st, wp)
            self._keys.append(id(w))


def prepare_network(net: torch.nn.Module):
    """Called at the top of the generator's / discriminator's forward."""
    grp = net.__dict__.get("_l2i_sn_group")
    if grp is None:
        grp = SNGroup(net)
        net.__dict__["_l2i_sn_group"] = grp      # not a submodule / parameter / buffer: invisible to state_dict
    grp.prepare(torch.is_grad_enabled())
