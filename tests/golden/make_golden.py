"""Generate the golden fixtures in this directory by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is imported read-only with one shim (``torch.Tensor.cuda`` -> identity, because
the reference hard-codes ``.cuda()``).  Weights come from
``layout2img_b200.synth.make_state`` (seeded, converged spectral-norm vectors) and inputs from
``synthetic_layout``, so every arm can regenerate them; only the reference's *outputs* are
stored: small tensors in full, large ones as (sum, abs-sum, l2, strided sample) summaries.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)       # the reference's `model` / `utils` packages must win over this repository's `model` shim
sys.path.insert(1, ROOT)

torch.Tensor.cuda = lambda self, *a, **k: self      # the one shim
import warnings
warnings.filterwarnings("ignore")

from layout2img_b200.synth import make_state, schema_of, synthetic_layout  # noqa: E402
from model.resnet_generator_app_v2 import (ResnetGenerator128_context, ResnetGenerator128,  # noqa: E402
                                           BoxRelationalEmbedding, bbox_mask)
from model.rcnn_discriminator_app import CombineDiscriminator128_app  # noqa: E402
from utils.bilinear import masks_to_layout  # noqa: E402
import torchvision  # noqa: E402

N_SAMPLE = 64
DROPOUT_SEED = 777


def summarize(t: torch.Tensor, n: int = N_SAMPLE) -> np.ndarray:
    """[sum, abs-sum, l2, sample...] in float64; sample = n evenly spaced flat entries."""
    f = t.detach().double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, steps=min(n, f.numel())).long()
    head = torch.stack([f.sum(), f.abs().sum(), f.norm()])
    samp = torch.zeros(n, dtype=torch.float64)
    samp[: idx.numel()] = f[idx]
    return torch.cat([head, samp]).numpy()


def build(seed_g=11, seed_d=12, num_classes=184, context=True):
    G = (ResnetGenerator128_context if context else ResnetGenerator128)(num_classes=num_classes, output_dim=3)
    D = CombineDiscriminator128_app(num_classes=num_classes)
    G.load_state_dict(make_state(schema_of(G), seed_g))
    D.load_state_dict(make_state(schema_of(D), seed_d))
    return G, D


def run_case(name, batch, num_obj, num_classes, n_pad, seed):
    out = {}
    data = synthetic_layout(batch, num_obj, num_classes, seed=seed, n_pad=n_pad)
    real, label, bbox, z, z_im = (data[k] for k in ("real", "label", "bbox", "z", "z_im"))
    G, D = build(num_classes=num_classes)

    # --- leaf ops, stored in full -----------------------------------------------------
    out["op.bbox_mask64"] = np.packbits(bbox_mask(z, bbox, 64, 64).numpy().astype(np.uint8))
    out["op.bbox_mask128"] = np.packbits(bbox_mask(z, bbox, 128, 128).numpy().astype(np.uint8))
    out["op.box_rel_emb"] = BoxRelationalEmbedding(bbox).numpy()
    gm = torch.Generator().manual_seed(seed + 100)
    masks = torch.rand(batch, num_obj, 16, 16, generator=gm)
    out["op.masks_to_layout"] = masks_to_layout(bbox, masks, 64).numpy()

    # --- eval forward -----------------------------------------------------------------
    G.eval(); D.eval()
    taps = {}
    hooks = [G.context.register_forward_hook(lambda m, i, o: taps.__setitem__("context", o)),
             G.mask_regress.register_forward_hook(lambda m, i, o: taps.__setitem__("bmask", o))]
    for k in range(1, 6):
        hooks.append(getattr(G, f"res{k}").register_forward_hook(
            lambda m, i, o, k=k: taps.__setitem__(f"x{k}", o[0])))
    with torch.no_grad():
        fake = G(z, bbox, z_im, label)
        d_out = D(fake, bbox.clone(), label.unsqueeze(-1))
    for h in hooks:
        h.remove()
    out["eval.fake"] = fake.numpy() if name == "C" else summarize(fake)
    out["eval.fake.sum"] = summarize(fake)
    out["eval.context"] = taps["context"].numpy()
    out["eval.bmask"] = summarize(taps["bmask"], 512)
    for k in range(1, 6):
        out[f"eval.x{k}"] = summarize(taps[f"x{k}"], 256)
    for i, nm in enumerate(("d_im", "d_obj", "d_app")):
        out[f"eval.{nm}"] = d_out[i].numpy()

    # --- one full training iteration (train_context_app_v2.py:155-189, no VGG term) -----
    G.train(); D.train()
    g_opt = torch.optim.Adam([{"params": [p], "lr": 1e-4} for p in G.parameters()], betas=(0.0, 0.999))
    d_opt = torch.optim.Adam([{"params": [p], "lr": 1e-4} for p in D.parameters()], betas=(0.0, 0.999))
    lab3 = label.unsqueeze(-1)
    torch.manual_seed(DROPOUT_SEED)        # the only RNG draw in the step is PSP's Dropout2d
    D.zero_grad()
    r_im, r_obj, r_app = D(real, bbox.clone(), lab3)
    fake = G(z, bbox, z_im, y=label)
    f_im, f_obj, f_app = D(fake.detach(), bbox.clone(), lab3)
    relu = torch.nn.functional.relu
    d_loss = (1.0 * (relu(1.0 - r_obj).mean() + relu(1.0 + f_obj).mean())
              + 0.1 * (relu(1.0 - r_im).mean() + relu(1.0 + f_im).mean())
              + 1.0 * (relu(1.0 - r_app).mean() + relu(1.0 + f_app).mean()))
    d_loss.backward()
    for k, p in D.named_parameters():
        out[f"train.dgrad.{k}"] = summarize(p.grad)
    d_opt.step()
    G.zero_grad()
    g_im, g_obj, g_app = D(fake, bbox.clone(), lab3)
    g_loss = (-g_obj.mean() * 1.0 - g_im.mean() * 0.1 + (fake - real).abs().mean() - 1.0 * g_app.mean())
    g_loss.backward()
    for k, p in G.named_parameters():
        out[f"train.ggrad.{k}"] = summarize(p.grad)
    g_opt.step()
    out["train.fake"] = summarize(fake, 1024)
    out["train.real_out"] = np.concatenate([t.detach().numpy().reshape(-1) for t in (r_im, r_obj, r_app)])
    out["train.fake_out"] = np.concatenate([t.detach().numpy().reshape(-1) for t in (f_im, f_obj, f_app)])
    out["train.g_out"] = np.concatenate([t.detach().numpy().reshape(-1) for t in (g_im, g_obj, g_app)])
    out["train.d_loss"] = np.array(d_loss.item())
    out["train.g_loss"] = np.array(g_loss.item())
    for k, v in G.state_dict().items():
        out[f"train.gstate.{k}"] = summarize(v.float())
    for k, v in D.state_dict().items():
        out[f"train.dstate.{k}"] = summarize(v.float())

    meta = dict(case=name, batch=batch, num_obj=num_obj, num_classes=num_classes, n_pad=n_pad, seed=seed,
                seed_g=11, seed_d=12, dropout_seed=DROPOUT_SEED, torch=torch.__version__,
                torchvision=torchvision.__version__, n_sample=N_SAMPLE,
                reference="wtliao/layout2img (unmodified modules, .cuda() shimmed)")
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)
    print(name, "d_loss", d_loss.item(), "g_loss", g_loss.item(), "keys", len(out))
    return G, D


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    G, D = build()
    json.dump({k: list(v) for k, v in schema_of(G).items()}, open(os.path.join(HERE, "schema_G.json"), "w"), indent=0)
    json.dump({k: list(v) for k, v in schema_of(D).items()}, open(os.path.join(HERE, "schema_D.json"), "w"), indent=0)
    G0 = ResnetGenerator128(num_classes=184, output_dim=3)
    json.dump({k: list(v) for k, v in schema_of(G0).items()}, open(os.path.join(HERE, "schema_G_nocontext.json"), "w"), indent=0)
    run_case("C", 2, 4, 184, 0, 0)          # BASELINE.json configs[0]
    run_case("Cpad", 2, 4, 184, 1, 1)       # padded objects: key masking + ROI filtering
    run_case("V", 2, 16, 179, 3, 2)         # VG-shape object count / class count (b=1 raises in PSP BatchNorm, as in the reference)


if __name__ == "__main__":
    main()
