"""Helpers shared by the end-to-end parity tests (small golden cases and BASELINE's full sizes)."""
from __future__ import annotations

import os

import torch

from layout2img_b200.synth import make_state
from oracle import l2i_oracle as O

RTOL, ATOL = 1e-3, 1e-4          # north_star tolerance (fp32)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(got, want, rtol=RTOL, atol=ATOL, what=""):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    err = (got - want).abs()
    bad = err > atol + rtol * want.abs()
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} outside tol, max err {err.max().item():.3e} "
                           f"(ref max {want.abs().max().item():.3e})")


def close_modulo_kinks(got, want, rtol, atol, what=""):
    """`close` for two runs of OUR path whose summation order differs (atomics, different tile compositions): a
    pre-activation within rounding distance of zero can land on either side of a ReLU, which moves the few gradient
    elements it feeds by a visible amount.  At most max(2, 0.1 %) of the elements may miss (rtol, atol), none by more than
    5 % of the tensor's maximum."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    err = (got - want).abs()
    bad = err > atol + rtol * want.abs()
    m = want.abs().max().item()
    assert int(bad.sum()) <= max(2, bad.numel() // 1000) and err.max().item() <= max(atol, 5e-2 * m), \
        (f"{what}: {int(bad.sum())}/{bad.numel()} outside tol, max err {err.max().item():.3e} (ref max {m:.3e})")


def oracle_step(schema_g, schema_d, seed_g, seed_d, data, keep, dtype, device="cpu"):
    """One oracle iteration in `dtype` on `device`; returns grads ("d.<name>" after d_loss.backward, "g.<name>" after
    g_loss.backward), losses, fake and the post-step states.  On a CUDA device this is the same restatement running
    on library kernels (TF32 off) -- test infrastructure, used so that BASELINE's full sizes finish in seconds."""
    PG, PD = make_state(schema_g, seed_g), make_state(schema_d, seed_d)
    cv = lambda t: (t.to(dtype) if t.is_floating_point() else t).to(device)
    PG, PD = {k: cv(v) for k, v in PG.items()}, {k: cv(v) for k, v in PD.items()}
    O.set_requires_grad(PG); O.set_requires_grad(PD)
    og, od = O.make_adam(PG, 1e-4), O.make_adam(PD, 1e-4)
    out = {}
    od_step, og_step = od.step, og.step

    def d_step(*a, **k):
        for n in O.param_names(PD):
            out["d." + n] = PD[n].grad.detach().clone()
        return od_step(*a, **k)

    def g_step(*a, **k):
        for n in O.param_names(PG):
            out["g." + n] = PG[n].grad.detach().clone()
        return og_step(*a, **k)

    od.step, og.step = d_step, g_step
    rd, rg, rfake = O.train_step(PG, PD, og, od, cv(data["real"]), data["label"].to(device), cv(data["bbox"]), cv(data["z"]),
                                 cv(data["z_im"]), dropout_mask=cv(keep))
    out.update(d_loss=rd, g_loss=rg, fake=rfake, PG={k: v.detach() for k, v in PG.items()},
               PD={k: v.detach() for k, v in PD.items()})
    return out


def grad_close(got, want32, want64, what):
    """End-to-end gradient parity against the fp32 oracle.

    tol1 = 2e-3 * |ref| + 1e-5 + 2e-4 * max|ref| + 4 * (the fp32 oracle's own max deviation from the fp64
    oracle on this tensor -- gradients that pass through 1/(sum_o m + 1e-6) are only good to ~3e-3 of their
    max in the reference's own fp32 arithmetic).

    A ReLU whose input is within rounding distance of 0 may land on the other side in any implementation
    that is not bit-identical to the reference (the fp32 and fp64 oracles disagree with each other the same
    way, tools/grad_diag.py); every such kink crossing adds or removes one pixel's contribution to the
    gradients upstream of it.  So on top of tol1: all but 1 % of a tensor's elements must be within
    5e-3 * max|ref|, all but max(2, 1e-3 * numel) within 5e-2 * max|ref| (an element that is the sum of a handful of
    terms -- fc.bias: one term per image; a whole row of an ISLA projection's gradient: one d beta term per object --
    moves by a large fraction of itself when a single kink flips), and the relative L2 error below 2e-2.  A wrong formula or index moves most elements by O(max|ref|) and fails all three.
    (The tight, kink-free comparisons of every kernel's backward are the per-operator tests in test_gpu_ops.py.)"""
    got, w32, w64 = got.detach().double().cpu(), want32.detach().double().cpu(), want64.detach().double().cpu()
    m = w32.abs().max().item()
    if m == 0.0:
        assert got.abs().max().item() <= 1e-12, f"{what}: reference gradient is exactly zero"
        return
    noise = (w32 - w64).abs().max().item()
    err = (got - w32).abs()
    tol1 = 2e-3 * w32.abs() + 1e-5 + 2e-4 * m + 4 * noise
    n_loose = int((err > tol1 + 5e-3 * m).sum())
    msg = (f"{what}: max err {err.max().item():.3e}, ref max {m:.3e}, fp32-ref noise {noise:.3e}, "
           f"{int((err > tol1).sum())}/{err.numel()} outside tol1, {n_loose} outside tol1 + 5e-3 max")
    assert n_loose <= max(1, int(0.01 * err.numel())), msg
    assert int((err > tol1 + 5e-2 * m).sum()) <= max(2, int(1e-3 * err.numel())), msg
    assert err.norm().item() <= 2e-2 * w32.norm().item() + 8 * noise * err.numel() ** 0.5, msg


def tensor_class(name: str) -> str:
    """Coarse class of a parameter for the per-class tolerance report."""
    leaf = name.rsplit(".", 1)[-1]
    if "weight_proj" in name or "bias_proj" in name:
        return "ISLA gamma/beta projections"
    if name.startswith(("g.context", "g.label_embedding")):
        return "G attention / embedding"
    if name.startswith("g.mask_regress"):
        return "G mask regression"
    if "conv_mask" in name or name.startswith("g.alpha"):
        return "G mask heads / alpha"
    if leaf == "bias":
        return ("D" if name.startswith("d.") else "G") + " biases"
    if name.startswith("d."):
        return "D conv / linear weights"
    return "G conv / linear weights"


def parity_report(path, title, got, r32, r64, keys):
    """Per-tensor table: ours vs fp64 oracle next to fp32 oracle vs fp64 oracle (max-norm and relative L2), plus a
    per-class summary.  Returns the text."""
    rows, classes, zero = [], {}, []
    for k in keys:
        w = r64[k].detach().double().cpu()
        if w.abs().max().item() < 1e-12:
            # biases in front of a normalisation layer: the true gradient is exactly 0, every arm returns rounding residue
            zero.append((got[k].detach().double().abs().max().item(), r32[k].detach().double().abs().max().item(), k))
            continue
        m = max(w.abs().max().item(), 1e-30)
        n2 = max(w.norm().item(), 1e-30)
        eo = (got[k].detach().double().cpu() - w)
        er = (r32[k].detach().double().cpu() - w)
        row = (eo.abs().max().item() / m, er.abs().max().item() / m, eo.norm().item() / n2, er.norm().item() / n2, m, k)
        rows.append(row)
        c = classes.setdefault(tensor_class(k), [0.0, 0.0, 0.0, 0.0, 0])
        for i in range(4):
            c[i] = max(c[i], row[i])
        c[4] += 1
    lines = [title, "max|err|/max|ref| and relative L2 error against the fp64 oracle: CUDA path (ours) | fp32 oracle (the "
             "reference arithmetic's own rounding noise)", "",
             f"{'tensor class':34s} {'n':>4s} {'ours max':>10s} {'fp32 max':>10s} {'ours L2':>10s} {'fp32 L2':>10s}"]
    for c, v in sorted(classes.items()):
        lines.append(f"{c:34s} {v[4]:4d} {v[0]:10.2e} {v[1]:10.2e} {v[2]:10.2e} {v[3]:10.2e}")
    lines += ["", f"{'ours max':>10s} {'fp32 max':>10s} {'ours L2':>10s} {'fp32 L2':>10s} {'max|ref|':>10s}  tensor"]
    for r in sorted(rows, reverse=True):
        lines.append(f"{r[0]:10.2e} {r[1]:10.2e} {r[2]:10.2e} {r[3]:10.2e} {r[4]:10.2e}  {r[5]}")
    if zero:
        lines += ["", "tensors whose true gradient is exactly zero (conv biases in front of a normalisation layer): max|value| ours | fp32 oracle"]
        lines += [f"{a:10.2e} {b:10.2e}  {k}" for a, b, k in zero]
    text = "\n".join(lines) + "\n"
    if path:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(text)
    return text
