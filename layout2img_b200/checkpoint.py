"""Checkpoints in the reference authors' on-disk format (train_context_app_v2.py:77-103,215-217).

The reference trains under nn.DataParallel / DataParallelWithCallback and saves `netG.state_dict()` of the WRAPPED
module, so every key of `G_<epoch>.pth` / `D_<epoch>.pth` carries a `module.` prefix; its loaders strip the first seven
characters of every key (`k[7:]`), keep the keys the model knows, and `load_state_dict` the merged dict.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

PREFIX = "module."


def save_checkpoint(module: torch.nn.Module, path: str, data_parallel_prefix: bool = True) -> None:
    """torch.save of the state_dict as the reference writes it (train_context_app_v2.py:215-217): CPU tensors, keys
    prefixed with `module.` (what DataParallel's state_dict produces) unless data_parallel_prefix=False."""
    sd = OrderedDict(((PREFIX + k) if data_parallel_prefix else k, v.detach().cpu()) for k, v in module.state_dict().items())
    torch.save(sd, path)


def load_checkpoint(module: torch.nn.Module, path: str, map_location="cpu") -> dict:
    """The reference's loading sequence (train_context_app_v2.py:78-89, test_context_app_v2.py:48-59): strip the
    7-character `module.` prefix, intersect with the model's keys, update, load.  Files written without the prefix
    (a single-GPU run of the reference) are accepted too.  Returns {"loaded": [...], "ignored": [...], "missing": [...]}."""
    state = torch.load(path, map_location=map_location)
    stripped = OrderedDict()
    has_prefix = all(k.startswith(PREFIX) for k in state)
    for k, v in state.items():
        stripped[k[7:] if has_prefix else k] = v
    model_dict = module.state_dict()
    pretrained = {k: v for k, v in stripped.items() if k in model_dict}
    report = {"loaded": sorted(pretrained), "ignored": sorted(set(stripped) - set(model_dict)),
              "missing": sorted(set(model_dict) - set(stripped))}
    model_dict.update(pretrained)
    module.load_state_dict(model_dict)
    return report


def save_training_state(path: str, netG, netD, g_opt, d_opt, epoch: int, extra: Optional[dict] = None) -> None:
    """Everything needed to resume a run exactly (the reference only keeps the two network files and restarts Adam
    from zero, train_context_app_v2.py:71-103): both networks, both optimizers (torch.optim format), the epoch."""
    torch.save({"G": {k: v.detach().cpu() for k, v in netG.state_dict().items()},
                "D": {k: v.detach().cpu() for k, v in netD.state_dict().items()},
                "g_opt": g_opt.state_dict(), "d_opt": d_opt.state_dict(), "epoch": int(epoch), "extra": extra or {}}, path)


def load_training_state(path: str, netG, netD, g_opt, d_opt, map_location="cpu") -> int:
    st = torch.load(path, map_location=map_location)
    netG.load_state_dict(st["G"]); netD.load_state_dict(st["D"])
    g_opt.load_state_dict(st["g_opt"]); d_opt.load_state_dict(st["d_opt"])
    return int(st["epoch"])
