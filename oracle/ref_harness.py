"""Drive the UNMODIFIED reference modules staged in baseline/_ref/ (oracle/make_ref.py).  TEST / BENCH INFRASTRUCTURE ONLY.

Restates the loop body of the reference's train_context_app_v2.py:148-189 around its own nn.Modules (VGG term
excluded: the weights need network access; betas=(0.0, 0.999) because torch 2.11 rejects the script's int 0,
SURVEY.md section 0.6).  On CPU the reference's hard-coded `.cuda()` calls are neutralised with the one shim the
golden generator uses (tests/golden/make_golden.py); on a GPU nothing is patched.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
LAMB_OBJ, LAMB_IMG, LAMB_APP = 1.0, 0.1, 1.0       # train_context_app_v2.py:40-46


def available() -> bool:
    return os.path.exists(os.path.join(REF, "model", "resnet_generator_app_v2.py"))


def import_reference(cpu: bool):
    """-> (ResnetGenerator128_context, CombineDiscriminator128_app) classes of the reference.  Must run in a process
    that has not imported this repository's own top-level `model` shim package."""
    if not available():
        raise RuntimeError("baseline/_ref is not staged (run `python oracle/make_ref.py` in the build container)")
    if "model" in sys.modules and not os.path.abspath(getattr(sys.modules["model"], "__file__", "") or "").startswith(REF):
        raise RuntimeError("a different top-level `model` package is already imported in this process")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self       # the one shim
    import warnings
    warnings.filterwarnings("ignore")
    from model.rcnn_discriminator_app import CombineDiscriminator128_app
    from model.resnet_generator_app_v2 import ResnetGenerator128_context
    return ResnetGenerator128_context, CombineDiscriminator128_app


def make_optimizers(G, D, lr: float = 1e-4):
    """train_context_app_v2.py:113-127: one param group per tensor."""
    g_opt = torch.optim.Adam([{"params": [p], "lr": lr} for p in G.parameters()], betas=(0.0, 0.999))
    d_opt = torch.optim.Adam([{"params": [p], "lr": lr} for p in D.parameters()], betas=(0.0, 0.999))
    return g_opt, d_opt


def train_step(G, D, g_opt, d_opt, real, label, bbox, z, z_im=None, record=None):
    """train_context_app_v2.py:155-189 (bbox stays where the caller keeps it: the reference keeps it on the host,
    :153, and each D call clones it to the device at rcnn_discriminator_app.py:407)."""
    lab3 = label.unsqueeze(-1)
    D.zero_grad()
    r_im, r_obj, r_app = D(real, bbox.clone(), lab3)
    fake = G(z, bbox, z_im, y=label)
    f_im, f_obj, f_app = D(fake.detach(), bbox.clone(), lab3)
    d_loss = (LAMB_OBJ * (F.relu(1.0 - r_obj).mean() + F.relu(1.0 + f_obj).mean())
              + LAMB_IMG * (F.relu(1.0 - r_im).mean() + F.relu(1.0 + f_im).mean())
              + LAMB_APP * (F.relu(1.0 - r_app).mean() + F.relu(1.0 + f_app).mean()))
    d_loss.backward()
    if record is not None:
        record("d")
    d_opt.step()
    G.zero_grad()
    g_im, g_obj, g_app = D(fake, bbox.clone(), lab3)
    g_loss = (-g_obj.mean() * LAMB_OBJ - g_im.mean() * LAMB_IMG + (fake - real).abs().mean() - LAMB_APP * g_app.mean())
    g_loss.backward()
    if record is not None:
        record("g")
    g_opt.step()
    return d_loss.detach(), g_loss.detach(), fake.detach()
