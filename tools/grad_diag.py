"""Per-parameter gradient error of the CUDA path vs the oracle in fp64, next to the fp32 oracle's own
error vs fp64 (the reference's noise floor).  python tools/grad_diag.py [C|Cpad|V]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import dropout_keep_mask, load_case, load_schema
from layout2img_b200.synth import make_state, synthetic_layout
from layout2img_b200.train import make_optimizers, train_step
from oracle import l2i_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "C"
dev = torch.device("cuda:0")
z, meta = load_case(name)
data = synthetic_layout(meta["batch"], meta["num_obj"], meta["num_classes"], seed=meta["seed"], n_pad=meta["n_pad"])
keep = dropout_keep_mask(meta["dropout_seed"], meta["batch"])

def oracle(dtype):
    PG = make_state(load_schema("G", meta["num_classes"]), meta["seed_g"])
    PD = make_state(load_schema("D", meta["num_classes"]), meta["seed_d"])
    cv = lambda P: {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in P.items()}
    PG, PD = cv(PG), cv(PD)
    O.set_requires_grad(PG); O.set_requires_grad(PD)
    og, od = O.make_adam(PG, 1e-4), O.make_adam(PD, 1e-4)
    ref = {}
    s1 = od.step
    def d_step(*a, **k):
        for n in O.param_names(PD): ref["d." + n] = PD[n].grad.detach().clone()
        return s1(*a, **k)
    od.step = d_step
    s2 = og.step
    def g_step(*a, **k):
        for n in O.param_names(PG): ref["g." + n] = PG[n].grad.detach().clone()
        return s2(*a, **k)
    og.step = g_step
    c = lambda t: t.to(dtype) if t.is_floating_point() else t
    rd, rg, rfake = O.train_step(PG, PD, og, od, c(data["real"]), data["label"], c(data["bbox"]), c(data["z"]), c(data["z_im"]),
                                 dropout_mask=keep.to(dtype))
    ref["fake"] = rfake; ref["d_loss"] = rd; ref["g_loss"] = rg
    return ref

from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
from layout2img_b200.model.resnet_generator_app_v2 import ResnetGenerator128_context
G = ResnetGenerator128_context(num_classes=meta["num_classes"], output_dim=3)
D = CombineDiscriminator128_app(num_classes=meta["num_classes"])
G.load_state_dict(make_state(load_schema("G", meta["num_classes"]), meta["seed_g"]))
D.load_state_dict(make_state(load_schema("D", meta["num_classes"]), meta["seed_d"]))
G.to(dev).train(); D.to(dev).train()
G.res4.conv_mask[0].dropout_mask = keep.view(meta["batch"], 100)
g_opt, d_opt = make_optimizers(G, D)
got = {}
def record(tag):
    net = D if tag == "d" else G
    for n, p in net.named_parameters():
        got[tag + "." + n] = p.grad.detach().clone().cpu()
dl, gl, fake = train_step(G, D, g_opt, d_opt, data["real"].to(dev), data["label"].to(dev), data["bbox"].to(dev),
                          data["z"].to(dev), data["z_im"].to(dev), record=record)
got["fake"] = fake.cpu(); got["d_loss"] = dl.cpu(); got["g_loss"] = gl.cpu()
r32, r64 = oracle(torch.float32), oracle(torch.float64)
rows = []
for k in r64:
    w = r64[k].double()
    m = max(w.abs().max().item(), 1e-30)
    e_ours = (got[k].double() - w).abs().max().item() / m
    e_o32 = (r32[k].double() - w).abs().max().item() / m
    nbad = int(((got[k].double() - w).abs() > 1e-3 * m).sum())
    rows.append((e_ours, e_o32, m, f"{k}  [{nbad}/{w.numel()} elements off by >1e-3 max]"))
rows.sort(reverse=True)
print(f"case {name}: max|err|/max|ref| -- ours vs fp64 oracle | fp32 oracle vs fp64 oracle | max|ref|")
rows = [r for r in rows if r[2] > 1e-12]
for e1, e2, m, k in rows[:25]:
    print(f"  {e1:.3e} | {e2:.3e} | {m:.3e} | {k}")
print("  -- D only")
for e1, e2, m, k in [r for r in rows if r[3].startswith("d.")][:25]:
    print(f"  {e1:.3e} | {e2:.3e} | {m:.3e} | {k}")
print("  ... median ours", sorted(r[0] for r in rows)[len(rows)//2], "median o32", sorted(r[1] for r in rows)[len(rows)//2])
