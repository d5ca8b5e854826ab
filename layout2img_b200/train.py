"""The G+D training iteration the hot path serves, restated from the reference's
train_context_app_v2.py:148-189 (VGG perceptual term excluded: its weights need network access),
plus the one-process-per-GPU data-parallel wrapper (batch shard + one gradient all-reduce per
network per step).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

LAMB_OBJ, LAMB_IMG, LAMB_APP = 1.0, 0.1, 1.0       # train_context_app_v2.py:40-46


def make_optimizers(netG, netD, g_lr: float = 1e-4, d_lr: float = 1e-4):
    """Adam(betas=(0, 0.999)) with one param group per tensor (train_context_app_v2.py:113-127), executed as
    one multi-tensor kernel per network (optim.FusedAdam -> csrc/optim.cu)."""
    from .optim import FusedAdam
    g_opt = FusedAdam([{"params": [p], "lr": g_lr} for p in netG.parameters()], betas=(0.0, 0.999))
    d_opt = FusedAdam([{"params": [p], "lr": d_lr} for p in netD.parameters()], betas=(0.0, 0.999))
    return g_opt, d_opt


class _frozen:
    """Temporarily clear requires_grad on a module's parameters: in the G step the reference lets autograd
    compute D's weight gradients and then discards them at the next netD.zero_grad()
    (train_context_app_v2.py:156,178-188); skipping them changes nothing observable."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]

    def __enter__(self):
        for p in self.params:
            p.requires_grad_(False)

    def __exit__(self, *exc):
        for p in self.params:
            p.requires_grad_(True)


def d_loss_fn(real_out, fake_out):
    r_im, r_obj, r_app = real_out
    f_im, f_obj, f_app = fake_out
    return (LAMB_OBJ * (F.relu(1.0 - r_obj).mean() + F.relu(1.0 + f_obj).mean())
            + LAMB_IMG * (F.relu(1.0 - r_im).mean() + F.relu(1.0 + f_im).mean())
            + LAMB_APP * (F.relu(1.0 - r_app).mean() + F.relu(1.0 + f_app).mean()))


def g_loss_fn(g_out, fake, real):
    g_im, g_obj, g_app = g_out
    return (-g_obj.mean() * LAMB_OBJ - g_im.mean() * LAMB_IMG + (fake - real).abs().mean()
            - LAMB_APP * g_app.mean())


class GradAllReducer:
    """Flat-bucket gradient all-reduce (mean) over the default process group: one NCCL collective per
    network per step.  No-op on a single process."""

    def __init__(self, module: torch.nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.flat: Optional[torch.Tensor] = None

    def __call__(self):
        if self.world == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        n = sum(g.numel() for g in grads)
        if self.flat is None or self.flat.numel() != n or self.flat.device != grads[0].device:
            self.flat = torch.empty(n, dtype=torch.float32, device=grads[0].device)
        off = 0
        views = []
        for g in grads:
            v = self.flat[off:off + g.numel()]
            v.copy_(g.reshape(-1))
            views.append(v)
            off += g.numel()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.div_(self.world)
        for p, v in zip(self.params, views):
            if p.grad is None:
                p.grad = v.view_as(p).clone()
            else:
                p.grad.copy_(v.view_as(p))


def train_step(netG, netD, g_opt, d_opt, real, label, bbox, z, z_im=None, sync_g=None, sync_d=None,
               record=None):
    """One D step then one G step; returns (d_loss, g_loss, fake) as detached tensors (no host sync).
    `record(tag)` is an optional callback used by the parity tests to snapshot gradients."""
    lab3 = label.unsqueeze(-1) if label.dim() == 2 else label
    # ---- D step (:155-174)
    netD.zero_grad()
    real_out = netD(real, bbox, lab3)
    fake = netG(z, bbox, z_im, y=label.view(label.shape[0], -1))
    fake_out = netD(fake.detach(), bbox, lab3)
    d_loss = d_loss_fn(real_out, fake_out)
    d_loss.backward()
    if sync_d is not None:
        sync_d()
    if record is not None:
        record("d")
    d_opt.step()
    # ---- G step (:177-189); D's gradients produced here are discarded by the next zero_grad
    netG.zero_grad()
    with _frozen(netD):
        g_out = netD(fake, bbox, lab3)
        g_loss = g_loss_fn(g_out, fake, real)
        g_loss.backward()
    if sync_g is not None:
        sync_g()
    if record is not None:
        record("g")
    g_opt.step()
    return d_loss.detach(), g_loss.detach(), fake.detach()
