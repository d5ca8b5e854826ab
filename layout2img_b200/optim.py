"""Multi-tensor Adam over libl2i.so (csrc/optim.cu): torch.optim.Adam's update for every parameter of
a network in ONE kernel launch.

The reference builds Adam with one parameter group per tensor (train_context_app_v2.py:113-127), which
torch executes as ~8 small kernels per tensor (1 600 launches per G+D step).  Same constructor surface
(params or a list of {"params": [...], "lr": ...} groups, lr, betas, eps), same arithmetic, same state
names (`exp_avg`, `exp_avg_sq`, `step`).  weight_decay / amsgrad / maximize are not used by the
reference and are rejected.
"""
from __future__ import annotations

from typing import Dict, Iterable, List

import numpy as np
import torch

from ._lib import call

_ENTRY = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("lr", "<f4"), ("pad", "<i4")])
CHUNK = 16384            # elements per CTA


class FusedAdam:
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = False):
        if weight_decay != 0.0 or amsgrad:
            raise ValueError("FusedAdam implements the reference's configuration: weight_decay=0, amsgrad=False")
        params = list(params)
        if params and isinstance(params[0], dict):
            groups = [dict(g) for g in params]
        else:
            groups = [{"params": params}]
        self.param_groups: List[dict] = []
        for g in groups:
            ps = g["params"]
            ps = [ps] if isinstance(ps, torch.Tensor) else list(ps)
            self.param_groups.append({"params": ps, "lr": float(g.get("lr", lr)),
                                      "betas": tuple(float(b) for b in g.get("betas", betas)),
                                      "eps": float(g.get("eps", eps))})
        b0, e0 = self.param_groups[0]["betas"], self.param_groups[0]["eps"]
        if any(g["betas"] != b0 or g["eps"] != e0 for g in self.param_groups):
            raise ValueError("FusedAdam: betas/eps must be the same in every group (lr may differ)")
        self.betas, self.eps = b0, e0
        self.state: Dict[torch.Tensor, dict] = {}
        self.step_count = 0
        self._plan = None

    # ------------------------------------------------------------------------------------------
    def _tensors(self):
        return [(p, g["lr"]) for g in self.param_groups for p in g["params"] if p.requires_grad or p in self.state]

    def _build_plan(self, items):
        dev = items[0][0].device
        total = sum((p.numel() + 3) // 4 * 4 for p, _ in items)      # 16-byte aligned slices
        m_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        v_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off, chunks = 0, []
        for i, (p, _) in enumerate(items):
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError("FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
            n = p.numel()
            self.state[p] = {"step": 0, "exp_avg": m_flat[off:off + n].view_as(p), "exp_avg_sq": v_flat[off:off + n].view_as(p)}
            chunks += [(i, c) for c in range((n + CHUNK - 1) // CHUNK)]
            off += (n + 3) // 4 * 4
        hosts = [torch.empty(len(items) * _ENTRY.itemsize, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._plan = {
            "items": items, "dev": dev, "hosts": hosts, "table": torch.empty_like(hosts[0], device=dev),
            "chunks": torch.tensor(chunks, dtype=torch.int32, device=dev).contiguous(), "n_chunks": len(chunks),
            "np": [h.numpy().view(_ENTRY) for h in hosts], "events": [None, None],
        }

    @torch.no_grad()
    def step(self):
        items = self._tensors()
        if not items:
            return
        if self._plan is None or [id(p) for p, _ in self._plan["items"]] != [id(p) for p, _ in items]:
            self._build_plan(items)
        pl = self._plan
        self.step_count += 1
        t = self.step_count
        beta1, beta2 = self.betas
        slot = t & 1                            # two pinned staging tables: never rewrite one still being copied
        if pl["events"][slot] is not None:
            pl["events"][slot].synchronize()
        tab = pl["np"][slot]
        for i, (p, lr) in enumerate(items):
            st = self.state[p]
            st["step"] = t
            g = p.grad
            if g is None:                      # torch skips parameters without a gradient
                tab[i] = (p.data_ptr(), 0, st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), 0, lr, 0)
                continue
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = p.grad = g.float().contiguous()
            tab[i] = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), lr, 0)
        pl["table"].copy_(pl["hosts"][slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        pl["events"][slot] = ev
        call("l2i_adam_step", pl["table"], pl["chunks"], pl["n_chunks"], CHUNK, beta1, beta2, self.eps,
             1.0 - beta1 ** t, float(np.sqrt(1.0 - beta2 ** t)))

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()
