"""Synthetic layouts and deterministic network state for tests, smoke and bench.

Host-side, torch-CPU only.  Two generators:

* ``synthetic_layout`` -- the input contract of the data loaders the hot path sits behind
  (reference data/cocostuff_loader.py:222-380: image [3,H,W] in [-1,1], objs [O], boxes [O,4]
  xywh in [0,1], padded with label 0 / box [-0.6,-0.6,0.5,0.5] at :301-303), drawn in the
  fixed order SURVEY.md section 8(d) prescribes so every arm of a comparison sees the same
  numbers.
* ``make_state`` -- fills a ``state_dict`` schema (name -> shape) with seeded values whose
  spectral-norm ``_u/_v`` buffers are already converged, so that activations are sane from
  the first forward without shipping 400 MB of weights.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Sequence, Tuple

import torch

PAD_BOX = (-0.6, -0.6, 0.5, 0.5)  # reference data/cocostuff_loader.py:301-303


def synthetic_layout(batch: int, num_obj: int, num_classes: int = 184, img_size: int = 128,
                     seed: int = 0, n_pad: int = 0, z_dim: int = 128):
    """Returns dict(real, label, bbox, z, z_im) of CPU tensors.

    ``n_pad`` trailing objects of every image are replaced with the loaders' padding entry
    (label 0, PAD_BOX) to exercise key masking in the attention and ROI filtering in D.
    """
    g = torch.Generator().manual_seed(seed)
    wh = torch.rand(batch, num_obj, 2, generator=g) * 0.8 + 0.1
    xy = torch.rand(batch, num_obj, 2, generator=g) * (1 - wh)
    label = torch.randint(1, num_classes, (batch, num_obj), generator=g)
    z = torch.randn(batch, num_obj, z_dim, generator=g)
    z_im = torch.randn(batch, z_dim, generator=g)
    real = torch.rand(batch, 3, img_size, img_size, generator=g) * 2 - 1
    bbox = torch.cat([xy, wh], dim=-1).float()
    if n_pad > 0:
        label[:, num_obj - n_pad:] = 0
        bbox[:, num_obj - n_pad:] = torch.tensor(PAD_BOX)
    return {"real": real, "label": label, "bbox": bbox, "z": z, "z_im": z_im}


def _power_iterate(w_mat: torch.Tensor, u: torch.Tensor, iters: int, eps: float):
    v = None
    for _ in range(iters):
        v = torch.mv(w_mat.t(), u)
        v = v / v.norm().clamp_min(eps)
        u = torch.mv(w_mat, v)
        u = u / u.norm().clamp_min(eps)
    return u, v


def make_state(schema: Mapping[str, Sequence[int]], seed: int = 0,
               sn_iters: int = 30) -> Dict[str, torch.Tensor]:
    """Deterministic values for every entry of a G or D ``state_dict`` schema.

    Keys are visited in sorted order with one CPU generator, so the result depends only on
    (schema, seed, torch version).  Weights are N(0, 1/fan_in)-scaled, biases small but
    non-zero (so bias paths are exercised), BN running stats away from (0, 1), and the
    spectral-norm vectors are the result of ``sn_iters`` power iterations.
    """
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    names = sorted(schema.keys())
    for name in names:
        shape = tuple(schema[name])
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros((), dtype=torch.long)
        elif leaf == "running_mean":
            out[name] = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            out[name] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif leaf in ("weight_u", "weight_v"):
            out[name] = torch.zeros(shape)  # filled below
        elif leaf == "bias":
            out[name] = torch.randn(shape, generator=g) * 0.05
        elif name.startswith("alpha"):
            out[name] = torch.randn(shape, generator=g) * 0.5
        elif len(shape) == 1:  # norm-layer scale
            out[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if "embedding" in name or leaf == "weight_orig" and name.split(".")[-2] in ("l_y", "l_y_app"):
                std = 1.0 if "label_embedding" in name else 1.0 / math.sqrt(shape[1])
            else:
                std = 1.0 / math.sqrt(fan_in)
            out[name] = torch.randn(shape, generator=g) * std
    for name in names:
        if name.endswith(".weight_orig"):
            base = name[: -len("_orig")]
            w = out[name]
            w_mat = w.reshape(w.shape[0], -1)
            u0 = torch.randn(w_mat.shape[0], generator=g)
            u0 = u0 / u0.norm()
            u, v = _power_iterate(w_mat, u0, sn_iters, 1e-12)
            out[base + "_u"] = u
            out[base + "_v"] = v
    return out


def schema_of(module: torch.nn.Module) -> Dict[str, Tuple[int, ...]]:
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}
