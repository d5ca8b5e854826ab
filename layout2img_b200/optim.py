"""Multi-tensor Adam over libl2i.so (csrc/optim.cu): torch.optim.Adam's update for every parameter of
a network in ONE kernel launch.

The reference builds Adam with one parameter group per tensor (train_context_app_v2.py:113-127), which
torch executes as ~8 small kernels per tensor (1 600 launches per G+D step).  Same constructor surface
(params or a list of {"params": [...], "lr": ...} groups, lr, betas, eps), same arithmetic, same state
names (`exp_avg`, `exp_avg_sq`, `step` -- one step count PER PARAMETER, advanced only when the parameter has
a gradient, as torch does), `state_dict()` / `load_state_dict()` in torch.optim's format (so optimizer state
saved by torch.optim.Adam resumes here and vice versa) and `add_param_group`.  weight_decay / amsgrad /
maximize are not used by the reference and are rejected.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch

from ._lib import call

_ENTRY = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("step_size", "<f4"),
                   ("bc2_sqrt", "<f4")])
CHUNK = 16384            # elements per CTA


class FusedAdam:
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = False, capturable: bool = False):
        if weight_decay != 0.0 or amsgrad:
            raise ValueError("FusedAdam implements the reference's configuration: weight_decay=0, amsgrad=False")
        # capturable: every step issues the same device work (one counter increment + one kernel over a STATIC table), so
        # it can be recorded in a CUDA graph.  Needs every parameter's .grad at a fixed address (train.GradBuckets) and
        # gives all tensors one common step count kept on the device.
        self.capturable = bool(capturable)
        self._step_dev = None
        self.defaults = {"lr": float(lr), "betas": tuple(float(b) for b in betas), "eps": float(eps)}
        self.param_groups: List[dict] = []
        self.state: Dict[torch.Tensor, dict] = {}
        self._plan = None
        params = list(params)
        groups = [dict(g) for g in params] if params and isinstance(params[0], dict) else [{"params": params}]
        for g in groups:
            self.add_param_group(g)

    def add_param_group(self, group: dict):
        ps = group["params"]
        ps = [ps] if isinstance(ps, torch.Tensor) else list(ps)
        seen = {id(p) for g in self.param_groups for p in g["params"]}
        if any(id(p) in seen for p in ps):
            raise ValueError("some parameters appear in more than one parameter group")
        g = {"params": ps, "lr": float(group.get("lr", self.defaults["lr"])),
             "betas": tuple(float(b) for b in group.get("betas", self.defaults["betas"])),
             "eps": float(group.get("eps", self.defaults["eps"]))}
        if self.param_groups and (g["betas"] != self.param_groups[0]["betas"] or g["eps"] != self.param_groups[0]["eps"]):
            raise ValueError("FusedAdam: betas/eps must be the same in every group (lr may differ)")
        self.param_groups.append(g)
        self._plan = None                     # rebuilt on the next step; existing moments are carried over

    @property
    def betas(self):
        return self.param_groups[0]["betas"]

    @property
    def eps(self):
        return self.param_groups[0]["eps"]

    # ------------------------------------------------------------------------------------------
    def _tensors(self):
        return [(p, g) for g in self.param_groups for p in g["params"] if p.requires_grad or p in self.state]

    def _build_plan(self, items):
        """Flat moment buffers + chunk list for `items`.  Moments and step counts of tensors that already have state
        are COPIED into the new buffers; only tensors seen for the first time start from zero."""
        dev = items[0][0].device
        total = sum((p.numel() + 3) // 4 * 4 for p, _ in items)      # 16-byte aligned slices
        m_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        v_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off, chunks = 0, []
        for i, (p, _) in enumerate(items):
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError("FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
            n = p.numel()
            m, v = m_flat[off:off + n].view_as(p), v_flat[off:off + n].view_as(p)
            old = self.state.get(p)
            if old is not None:
                m.copy_(old["exp_avg"]); v.copy_(old["exp_avg_sq"])
            self.state[p] = {"step": old["step"] if old is not None else 0, "exp_avg": m, "exp_avg_sq": v}
            chunks += [(i, c) for c in range((n + CHUNK - 1) // CHUNK)]
            off += (n + 3) // 4 * 4
        hosts = [torch.empty(len(items) * _ENTRY.itemsize, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._plan = {
            "ids": [id(p) for p, _ in items], "dev": dev, "hosts": hosts, "table": torch.empty_like(hosts[0], device=dev),
            "chunks": torch.tensor(chunks, dtype=torch.int32, device=dev).contiguous(), "n_chunks": len(chunks),
            "np": [h.numpy().view(_ENTRY) for h in hosts], "events": [None, None], "slot": 0,
        }

    def _step_capturable(self, items):
        pl = self._plan
        beta1, beta2 = self.betas
        if pl.get("static_ptrs") is not None and not torch.cuda.is_current_stream_capturing() and \
                pl["static_ptrs"] != [(p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0) for p, _ in items]:
            pl["static_ptrs"] = None          # a gradient moved (GradBuckets re-cut its buckets): rebuild the table, keep the count
        if pl.get("static_ptrs") is None:
            tab = pl["np"][0]
            ptrs = []
            for i, (p, grp) in enumerate(items):
                if p.grad is None or p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                    raise ValueError("capturable FusedAdam needs every .grad allocated at a fixed address (train.GradBuckets)")
                st = self.state[p]
                tab[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                          grp["lr"], 1.0)
                ptrs.append((p.data_ptr(), p.grad.data_ptr()))
            pl["table"].copy_(pl["hosts"][0])
            torch.cuda.current_stream().synchronize()
            pl["static_ptrs"] = ptrs
            if self._step_dev is None:
                first = max(int(self.state[p]["step"]) for p, _ in items)
                self._step_dev = torch.full((1,), first, dtype=torch.int64, device=pl["dev"])
        self._step_dev.add_(1)
        call("l2i_adam_step", pl["table"], pl["chunks"], pl["n_chunks"], CHUNK, beta1, beta2, self.eps, self._step_dev)

    def sync_step_counts(self):
        """capturable mode keeps the step count on the device; copy it into state[p]['step'] (before state_dict())."""
        if self._step_dev is not None:
            t = int(self._step_dev.item())
            for st in self.state.values():
                st["step"] = t

    @torch.no_grad()
    def step(self):
        items = self._tensors()
        if not items:
            return
        if self._plan is None or self._plan["ids"] != [id(p) for p, _ in items]:
            self._build_plan(items)
        if self.capturable:
            return self._step_capturable(items)
        pl = self._plan
        beta1, beta2 = self.betas
        slot = pl["slot"] = pl["slot"] ^ 1      # two pinned staging tables: never rewrite one still being copied
        if pl["events"][slot] is not None:
            pl["events"][slot].synchronize()
        tab = pl["np"][slot]
        touched = []
        for i, (p, grp) in enumerate(items):
            st = self.state[p]
            g = p.grad
            if g is None:                      # torch skips parameters without a gradient (their step does not advance)
                tab[i] = (p.data_ptr(), 0, st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), 0, 0.0, 1.0)
                continue
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = p.grad = g.float().contiguous()
            t = st["step"] = int(st["step"]) + 1
            tab[i] = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                      grp["lr"] / (1.0 - beta1 ** t), math.sqrt(1.0 - beta2 ** t))
            touched.append(p)
        pl["table"].copy_(pl["hosts"][slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        pl["events"][slot] = ev
        call("l2i_adam_step", pl["table"], pl["chunks"], pl["n_chunks"], CHUNK, beta1, beta2, self.eps, None)
        # the raw-pointer update bypasses autograd's version counters: bump them so that a stale graph that saved one of
        # these parameters fails loudly in backward instead of silently using the new values
        if _HAS_SET_VERSION:
            torch._C._autograd._unsafe_set_version_counter(touched, [p._version + 1 for p in touched])
        else:
            for p in touched:
                p.add_(0)

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    # ------------------------------------------------------------------------------------------
    # torch.optim-compatible (de)serialisation: {"state": {index: {...}}, "param_groups": [{..., "params": [indices]}]}
    def state_dict(self) -> dict:
        self.sync_step_counts()
        index, groups = {}, []
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                index[id(p)] = len(index)
                ids.append(index[id(p)])
            groups.append({"lr": g["lr"], "betas": g["betas"], "eps": g["eps"], "weight_decay": 0, "amsgrad": False,
                           "maximize": False, "params": ids})
        state = {}
        for p, st in self.state.items():
            state[index[id(p)]] = {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"].clone(),
                                   "exp_avg_sq": st["exp_avg_sq"].clone()}
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: dict):
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"]) for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        by_index = {}
        for saved, mine in zip(groups, self.param_groups):
            mine["lr"] = float(saved["lr"])
            for idx, p in zip(saved["params"], mine["params"]):
                by_index[idx] = p
        self.state = {}
        for idx, st in sd["state"].items():
            p = by_index[int(idx)]
            self.state[p] = {"step": int(float(st["step"])),
                             "exp_avg": st["exp_avg"].to(p.device, torch.float32).reshape(p.shape).clone(),
                             "exp_avg_sq": st["exp_avg_sq"].to(p.device, torch.float32).reshape(p.shape).clone()}
        self._plan = None                     # the next step moves the loaded moments into flat buffers
        self._step_dev = None                 # (capturable) the device-side count restarts from the loaded one


_HAS_SET_VERSION = hasattr(torch._C._autograd, "_unsafe_set_version_counter")
