"""Parity at BASELINE.json's full sizes: one complete G+D training iteration at
    A  batch 64, 8 objects, 184 classes, 128x128   (configs[2], the benchmarked workload)
    V  batch 32, 16 objects, 179 classes, one padded object per image   (configs[4])
through the C ABI, against
  * the oracle (oracle/l2i_oracle.py) evaluated ON THE GPU in fp64 -- the same restatement on library kernels, so the
    full sizes finish in seconds; and in fp32 (TF32 off) to measure the reference arithmetic's own rounding noise;
  * the UNMODIFIED reference modules (baseline/_ref, staged by oracle/make_ref.py) run on the same GPU in fp32, TF32 off,
    in a subprocess (tests/ref_gpu_runner.py), when they are staged;
  * the CPU fp32 oracle on 8 images of the batch (eval mode, where nothing couples the images of a batch).
Losses, the generated batch, all 204 gradients and the post-step state are compared; the per-tensor error table
is written to gpurun_out/parity/ (committed copies: profiles/r02_fullsize_parity_*.txt).
Also here: every Appendix-A convolution shape at its full size, forward / data gradient / weight gradient vs fp64.
"""
import os
import subprocess
import sys
import tempfile

import pytest
import torch
import torch.nn.functional as F

from conftest import load_schema
from layout2img_b200.synth import make_state, synthetic_layout
from oracle import l2i_oracle as O
from parity_utils import ATOL, ROOT, RTOL, close, grad_close, oracle_step, parity_report

pytestmark = pytest.mark.gpu

CONFIGS = {"A": dict(batch=64, num_obj=8, num_classes=184, n_pad=0, seed=31),
           "V": dict(batch=32, num_obj=16, num_classes=179, n_pad=1, seed=32)}
SEED_G, SEED_D, DROPOUT_SEED = 41, 42, 777


def _reference_step(cfg):
    """Results of the unmodified reference on this GPU (None when baseline/_ref is not staged)."""
    if not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "model", "resnet_generator_app_v2.py")):
        return None
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "ref.pt")
        cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_gpu_runner.py"), out] + [
            str(v) for v in (cfg["batch"], cfg["num_obj"], cfg["num_classes"], cfg["seed"], cfg["n_pad"], SEED_G, SEED_D,
                             DROPOUT_SEED)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, "reference runner failed:\n" + r.stdout[-2000:] + r.stderr[-4000:]
        return torch.load(out, map_location="cpu")


@pytest.mark.parametrize("name", ["A", "V"])
def test_full_size_train_step(name):
    from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
    from layout2img_b200.model.resnet_generator_app_v2 import ResnetGenerator128_context
    from layout2img_b200.train import make_optimizers, train_step
    cfg = CONFIGS[name]
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ncls, b = cfg["num_classes"], cfg["batch"]
    sg, sd = load_schema("G", ncls), load_schema("D", ncls)
    data = synthetic_layout(b, cfg["num_obj"], ncls, seed=cfg["seed"], n_pad=cfg["n_pad"])

    ref = _reference_step(cfg)
    if ref is not None:
        keep = ref["keep"].view(b, 100, 1, 1)
    else:
        torch.manual_seed(DROPOUT_SEED)
        keep = (torch.rand(b, 100, 1, 1) >= 0.1).float() / 0.9

    # ---- the CUDA path
    G = ResnetGenerator128_context(num_classes=ncls, output_dim=3)
    D = CombineDiscriminator128_app(num_classes=ncls)
    G.load_state_dict(make_state(sg, SEED_G)); D.load_state_dict(make_state(sd, SEED_D))
    G.to(dev).train(); D.to(dev).train()
    G.res4.conv_mask[0].dropout_mask = keep.view(b, 100)
    g_opt, d_opt = make_optimizers(G, D)
    got = {}

    def record(tag):
        net = D if tag == "d" else G
        for n, p in net.named_parameters():
            got[tag + "." + n] = p.grad.detach().clone().cpu()

    dl, gl, fake = train_step(G, D, g_opt, d_opt, data["real"].to(dev), data["label"].to(dev), data["bbox"].to(dev),
                              data["z"].to(dev), data["z_im"].to(dev), record=record)
    torch.cuda.synchronize()
    got.update(d_loss=dl.cpu(), g_loss=gl.cpu(), fake=fake.cpu())
    sdG = {k: v.detach().cpu() for k, v in G.state_dict().items()}
    sdD = {k: v.detach().cpu() for k, v in D.state_dict().items()}
    del G, D, g_opt, d_opt, fake
    torch.cuda.empty_cache()

    # ---- oracle on the GPU: fp64 (truth) and fp32 (the reference arithmetic's rounding noise)
    def oracle(dtype):
        r = oracle_step(sg, sd, SEED_G, SEED_D, data, keep, dtype, device=dev)
        out = {k: (v.cpu() if torch.is_tensor(v) else {kk: vv.cpu() for kk, vv in v.items()}) for k, v in r.items()}
        del r
        torch.cuda.empty_cache()
        return out

    r64 = oracle(torch.float64)
    r32 = oracle(torch.float32)
    gkeys = [k for k in r64 if k.startswith(("d.", "g."))]
    report = parity_report(os.path.join(ROOT, "gpurun_out", "parity", f"fullsize_{name}.txt"),
                           f"config {name}: batch {b}, {cfg['num_obj']} objects, {ncls} classes -- one G+D training iteration",
                           got, r32, r64, gkeys + ["fake"])
    print(report[:3000])

    # ---- losses and the generated batch
    for k in ("d_loss", "g_loss"):
        for want in (r32[k].item(), r64[k].item()) + ((ref[k].item(),) if ref is not None else ()):
            assert abs(got[k].item() - want) <= ATOL + RTOL * abs(want), (k, got[k].item(), want)
    noise_fake = (r32["fake"].double() - r64["fake"]).abs().max().item()
    close(got["fake"], r64["fake"], RTOL, ATOL + 4 * noise_fake, what=f"train fake vs fp64 oracle (fp32 noise {noise_fake:.2e})")
    # ---- every gradient: against the fp32 oracle with the fp32-vs-fp64 noise allowance, and against the reference itself
    for k in gkeys:
        grad_close(got[k], r32[k], r64[k], k + " vs oracle")
    if ref is not None:
        close(got["fake"], ref["fake"], RTOL, ATOL + 4 * noise_fake, what="train fake vs reference modules on GPU")
        for k in gkeys:
            grad_close(got[k], ref[k], r64[k], k + " vs reference modules on GPU")
    # ---- post-step state (Adam-updated parameters, spectral-norm vectors, batch-norm running statistics); an element
    # whose gradient is rounding noise moves by +-lr in either direction (first Adam step with betas=(0, .999))
    for mine, theirs, tag in ((sdG, r32["PG"], "G"), (sdD, r32["PD"], "D")):
        for n, v in theirs.items():
            close(mine[n].float(), v.float(), 1e-3, 1e-3 if n.endswith(("_u", "_v")) else 2e-4, f"{tag} state {n}")


@pytest.mark.parametrize("name", ["A", "V"])
def test_full_size_eval_forward_cpu_oracle_spot_check(name):
    """Eval-mode forward of the full batch through the CUDA path; 8 of its images against the CPU fp32 oracle (in eval
    mode batch norm uses running statistics and D has no norm layer, so images are independent)."""
    from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
    from layout2img_b200.model.resnet_generator_app_v2 import ResnetGenerator128_context
    cfg = CONFIGS[name]
    dev = torch.device("cuda:0")
    ncls, b = cfg["num_classes"], cfg["batch"]
    PG, PD = make_state(load_schema("G", ncls), SEED_G), make_state(load_schema("D", ncls), SEED_D)
    G = ResnetGenerator128_context(num_classes=ncls, output_dim=3)
    D = CombineDiscriminator128_app(num_classes=ncls)
    G.load_state_dict(PG); D.load_state_dict(PD)
    G.to(dev).eval(); D.to(dev).eval()
    data = synthetic_layout(b, cfg["num_obj"], ncls, seed=cfg["seed"], n_pad=cfg["n_pad"])
    d = {k: v.to(dev) for k, v in data.items()}
    with torch.no_grad():
        fake = G(d["z"], d["bbox"], d["z_im"], d["label"])
        d_im, d_obj, d_app = D(fake, d["bbox"], d["label"].unsqueeze(-1))
    idx = torch.linspace(0, b - 1, 8).long()
    with torch.no_grad():
        rf = O.g_forward(PG, data["z"][idx], data["bbox"][idx], data["z_im"][idx], data["label"][idx], False)
        # D is checked on the CUDA path's own images, so that its comparison does not inherit G's rounding differences
        r_im, r_obj, r_app = O.d_forward(PD, fake[idx.to(dev)].cpu(), data["bbox"][idx], data["label"][idx], False)
    close(fake[idx.to(dev)], rf, what="full-batch eval fake (8 images) vs CPU oracle")
    close(d_im[idx.to(dev)], r_im, what="d_im (8 images) vs CPU oracle")
    # per-object outputs: [all large ROIs, all small ROIs] of the call -- compare the multiset belonging to the 8 images
    rois, lab = O.d_rois(data["bbox"], data["label"], 128)
    small = ((rois[:, 3] - rois[:, 1]) < 64) & ((rois[:, 4] - rois[:, 2]) < 64)
    img_of = torch.cat([rois[~small, 0], rois[small, 0]]).long()
    sel = torch.isin(img_of, idx)
    for mine, want, nm in ((d_obj, r_obj, "d_obj"), (d_app, r_app, "d_app")):
        a = mine.detach().cpu().flatten()[sel].sort().values
        w = want.flatten().sort().values
        assert a.shape == w.shape
        close(a, w, RTOL, ATOL * max(1.0, w.abs().max().item()), nm + " (8 images) vs CPU oracle")


# every 3x3 shape of SURVEY.md Appendix A plus the 1x1 shapes with peculiar channel counts: (N, Cin, Cout, H, k)
APPENDIX_A = [
    (512, 1024, 1024, 8, 3), (64, 512, 512, 32, 3), (512, 512, 1024, 8, 3), (64, 528, 100, 64, 3), (512, 256, 256, 16, 3),
    (64, 1024, 512, 16, 3), (64, 512, 256, 32, 3), (64, 256, 128, 64, 3), (64, 128, 64, 128, 3), (64, 256, 512, 32, 3),
    (512, 512, 512, 8, 3), (64, 1024, 1024, 8, 3), (64, 512, 512, 16, 3), (64, 256, 256, 32, 3), (64, 128, 128, 64, 3),
    (64, 64, 64, 128, 3), (512, 256, 256, 8, 3), (64, 64, 128, 64, 3), (64, 128, 256, 32, 3), (64, 256, 512, 16, 3),
    (64, 512, 1024, 8, 3), (64, 256, 100, 32, 3), (64, 1024, 1024, 4, 3), (64, 512, 100, 16, 3), (512, 256, 256, 4, 3),
    (64, 1024, 100, 8, 3), (64, 64, 3, 128, 3), (64, 3, 64, 128, 3),
    (64, 1024, 1024, 8, 1), (64, 128, 64, 128, 1), (64, 3, 64, 64, 1), (512, 512, 1024, 4, 1), (64, 100, 184, 64, 1),
    (512, 256, 1, 16, 1),
]


@pytest.mark.parametrize("shape", APPENDIX_A, ids=lambda s: "N%d_%dto%d_H%d_k%d" % s)
def test_conv_full_size_shapes_vs_fp64(shape):
    """Forward, data gradient and weight gradient (split-K atomics, multi-pass K loops) of the tensor-core convolution
    at the benchmarked sizes against fp64 library convolutions: max error <= 1e-4 of the result's max (the fp32-class
    bound; a single-pass bf16/TF32 product is 1e-2/1e-3)."""
    from layout2img_b200 import ops
    N, Cin, Cout, H, k = shape
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(N + Cin + Cout + H + k)
    x = torch.randn(N, Cin, H, H, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    dy = torch.randn(N, Cout, H, H, generator=g).to(dev)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()

    def rel(got, want):
        want = want.double()
        return ((got.double() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()

    wp = ops.conv_weight_prep(w)
    xp, dyp = ops.act_split(nhwc(x)), ops.act_split(nhwc(dy))
    y, _ = ops.conv2d_fwd(xp, wp.f_hi, wp.f_lo, Cout, k * k, bias=bias)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2)
    e_fwd = rel(y.permute(0, 3, 1, 2), ref)
    del ref, y
    dx, _ = ops.conv2d_fwd(dyp, wp.d_hi, wp.d_lo, Cin, k * k)
    ref = torch.nn.grad.conv2d_input(x.shape, w.double(), dy.double(), padding=k // 2)
    e_dg = rel(dx.permute(0, 3, 1, 2), ref)
    del ref, dx
    dw = ops.conv2d_wgrad(dyp, xp, k * k).view(Cout, k, k, Cin).permute(0, 3, 1, 2)
    ref = torch.nn.grad.conv2d_weight(x.double(), w.shape, dy.double(), padding=k // 2)
    e_wg = rel(dw, ref)
    assert max(e_fwd, e_dg, e_wg) <= 1e-4, f"fwd {e_fwd:.2e} dgrad {e_dg:.2e} wgrad {e_wg:.2e}"
